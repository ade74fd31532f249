"""Packs the reference's 44 stored beliefs (PackedManifoldKernelDensity JSON, N=100, RoME.Point2) into one
small .npz so that the GPU box -- which has no /root/reference -- can run the golden tests.

Source: /root/reference/test/ICRA2022_tests/test_data/tut3/*.json (written by upstream `storeAllBeliefs`,
test/ICRA2022_tests/insufficient_data.jl:59-64).  Each file holds `pts` (100 x 2) and the `bw` that
upstream `manikde!` selected for exactly those points, i.e. 88 (input -> output) pairs of the
leave-one-out bandwidth search, plus deterministic inputs for mmd / evaluate / product tests.

Run in the build container:  python tests/golden/make_tut3_fixture.py
"""
import glob
import json
import os

import numpy as np

SRC = "/root/reference/test/ICRA2022_tests/test_data/tut3"
names, pts, bws = [], [], []
for f in sorted(glob.glob(os.path.join(SRC, "*.json"))):
    d = json.load(open(f))
    assert d["_type"] == "IncrementalInference.PackedManifoldKernelDensity" and d["varType"] == "RoME.Point2"
    assert d["partial"] == [1, 2] and d["infoPerCoord"] == [1.0, 1.0]
    names.append(os.path.basename(f)[:-5])
    pts.append(np.array(d["pts"], dtype=np.float64))
    bws.append(np.array(d["bw"], dtype=np.float64))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tut3_beliefs.npz")
np.savez_compressed(out, names=np.array(names), pts=np.stack(pts), bw=np.stack(bws))
print(out, len(names), np.stack(pts).shape)
