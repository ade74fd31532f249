"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the header declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from __graft_entry__ import ROOT, build, load_package


@pytest.fixture(scope="module")
def cb():
    build()
    return load_package()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "caesar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(cb):
    L = cb._capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/caesar_b200.h but not exported"
    assert sorted(cb._capi.EXPORTS) == syms
    out = subprocess.check_output(["nm", "-D", "--defined-only", cb._capi.SO_PATH], text=True)
    exported = set(re.findall(r" T (cb2_[a-z0-9_]+)", out))
    assert set(syms) <= exported


def test_library_carries_sm100a_code_only(cb):
    out = subprocess.run(["cuobjdump", "-lelf", cb._capi.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(cb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cb.CB2Error) as ei:
        cb.Context(0)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)
    # and the product package does not import the oracle
    for fn in os.listdir(os.path.join(ROOT, "caesar.jl_b200")):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "caesar.jl_b200", fn)).read().replace("the oracle", "")


def test_serialization_roundtrip_matches_fixture_schema(cb, tut3):
    pts, bw = tut3["l1_1"]
    m = cb.ManifoldKernelDensity("Point2", pts, bw)
    packed = cb.packDistribution(m)
    assert set(packed) == {"_type", "varType", "pts", "bw", "partial", "infoPerCoord"}
    assert packed["_type"] == "IncrementalInference.PackedManifoldKernelDensity" and packed["varType"] == "RoME.Point2"
    assert packed["partial"] == [1, 2] and packed["infoPerCoord"] == [1.0, 1.0]
    import json
    m2 = cb.unpackDistribution(json.dumps(packed))
    assert np.array_equal(m2.pts, pts) and np.array_equal(m2.bw, bw) and m2.circular == (0, 0)


def test_host_api_validation(cb):
    with pytest.raises(KeyError):
        cb.ManifoldKernelDensity("NoSuchVariable", np.zeros((3, 2)), [1, 1])
    with pytest.raises(ValueError):
        cb.ManifoldKernelDensity("Pose2", np.zeros((3, 2)), [1, 1])
    m = cb.ManifoldKernelDensity("Pose2", np.zeros((3, 3)), 0.5)
    assert m.circular == (0, 0, 1) and np.allclose(m.bw, 0.5)
    p2 = cb.ManifoldKernelDensity("Point2", np.zeros((4, 2)), 1.0)
    with pytest.raises(ValueError):
        cb.ScatterAlignPose3(cloud1=p2, cloud2=p2)
    sap = cb.ScatterAlignPose2(cloud1=cb.ManifoldKernelDensity("Point2", np.zeros((4, 2)), 1.0),
                               cloud2=cb.ManifoldKernelDensity("Point2", np.zeros((4, 2)), 1.0))
    assert sap.align.sample_count == 75 and sap.align.bw == 5e-5  # reference defaults, ScatterAlignPose2.jl:12-13
    cache = cb.preambleCache(None, [], sap)
    assert set(cache) == {"M", "e0", "smps1", "smps2", "bw", "score"}
