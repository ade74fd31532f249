"""Host-side pieces of the ScatterAlign factor plugin that need no GPU: the packed factor format, the parched (stashed) form and
its reconstitution from a blob store in preambleCache.  Mirrors R:test/testScatterAlignPose2.jl:72-82 (pack / unpack keeps the
fields) and R:test/testScatterAlignParched.jl:62-130 (stashed clouds: small packed factor, clouds reloaded by preambleCache)."""
import json

import numpy as np
import pytest

from __graft_entry__ import load_package


@pytest.fixture(scope="module")
def cb():
    return load_package()


def _clouds(cb, dim, n=400, seed=0):
    rng = np.random.default_rng(seed)
    M = "Point2" if dim == 2 else "Point3"
    a = cb.ManifoldKernelDensity(M, rng.normal(size=(n, dim)), np.full(dim, 0.1))
    b = cb.ManifoldKernelDensity(M, rng.normal(size=(n, dim)) + 1.0, np.full(dim, 0.2))
    return a, b


@pytest.mark.parametrize("dim", [2, 3])
def test_pack_unpack_keeps_fields_and_points(cb, dim):
    c1, c2 = _clouds(cb, dim)
    cls = cb.ScatterAlignPose2 if dim == 2 else cb.ScatterAlignPose3
    sap = cls(cloud1=c1, cloud2=c2, sample_count=100, bw=1.0, rescale=0.5)
    psap = cb.packScatterAlign(sap)
    assert psap["_type"] == f"Caesar.PackedScatterAlignPose{dim}"
    assert psap["cloud1"]["_type"] == "IncrementalInference.PackedManifoldKernelDensity"
    sap_ = cb.unpackScatterAlign(json.dumps(psap))  # through the JSON text, as DFG.packFactor / unpackFactor do
    assert type(sap_) is cls
    assert sap.align.gridscale == sap_.align.gridscale
    assert sap.align.sample_count == sap_.align.sample_count
    assert sap.align.bw == sap_.align.bw
    assert np.array_equal(sap.align.cloud1.pts, sap_.align.cloud1.pts) and np.array_equal(sap.align.cloud1.bw, sap_.align.cloud1.bw)
    assert np.array_equal(sap.align.cloud2.pts, sap_.align.cloud2.pts) and np.array_equal(sap.align.cloud2.bw, sap_.align.cloud2.bw)


def test_stashed_factor_is_small_and_reloads_its_clouds(cb):
    c1, c2 = _clouds(cb, 3, n=5000)
    fg = cb.BlobStoreGraph()
    deb0 = fg.addData("x0", "pointcloud", json.dumps(cb.packDistribution(c1)).encode())
    deb1 = fg.addData("x1", "pointcloud", json.dumps(cb.packDistribution(c2)).encode())
    sap = cb.ScatterAlignPose3(cloud1=c1, cloud2=c2, sample_count=100, useStashing=True, dataEntry_cloud1=deb0, dataEntry_cloud2=deb1)
    jpf = json.dumps(cb.packScatterAlign(sap))
    assert len(jpf) < 1500  # the massive point clouds are not stored in the packed factor (testScatterAlignParched.jl:107)
    sap_ = cb.unpackScatterAlign(jpf)
    assert sap_.align.cloud1.N == 1 and sap_.align.useStashing
    cache = cb.preambleCache(fg, ["x0", "x1"], sap_)  # pulls both clouds from the blob store (ScatterAlignPose2.jl:124-137)
    assert cache["M"] == "SpecialEuclidean(3)" and cache["bw"][0] == sap.align.bw
    assert sap_.align.cloud1.N == 5000 and np.allclose(sap_.align.cloud1.pts, c1.pts) and np.allclose(sap_.align.cloud2.bw, c2.bw)


def test_stashing_needs_entries_and_a_store(cb):
    c1, c2 = _clouds(cb, 2)
    with pytest.raises(AssertionError):
        cb.packScatterAlign(cb.ScatterAlignPose2(cloud1=c1, cloud2=c2, useStashing=True))
    sap = cb.ScatterAlignPose2(cloud1=c1, cloud2=c2, useStashing=True, dataEntry_cloud1="a", dataEntry_cloud2="b")
    with pytest.raises(ValueError):
        cb.preambleCache(None, ["x0", "x1"], sap)


def test_get_manifold(cb):
    c1, c2 = _clouds(cb, 2)
    assert cb.getManifold(cb.ScatterAlignPose2(cloud1=c1, cloud2=c2)) == "SpecialEuclidean(2)"
