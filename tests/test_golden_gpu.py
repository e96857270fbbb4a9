"""GPU vs the committed golden fixtures (tests/golden/golden_v1.npz) -- no oracle involved at run time."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


@pytest.mark.parametrize("name", ["S64", "S128-odd", "S128-clutter"])
def test_gpu_reproduces_golden(dmf, ctx, name):
    g = np.load(GOLDEN)
    K = g["K"]; H, W = (int(v) for v in g["HW"])
    sc = dmf.scenes.scene(name)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    eng = dmf.RayTracingEngine(dmf.Camera(K, H, W), ctx)
    poses = g[f"{name}/poses"]; zd = int(g[f"{name}/zdelta"]); n = len(poses)
    gv._commit(ctx)
    assert np.array_equal(gv.occupied_cells_, g[f"{name}/occupied"])
    for sparse in (0, 1):
        ctx.reset_counters()
        r = eng.forward_views(gv, poses, dmf.MODE_POINTS, zd, bool(sparse))
        c = ctx.counters()
        want = sum(g[f"{name}/{i}/counters/s{sparse}"] for i in range(n))
        assert [c["samples"], c["inbounds"], c["hits"]] == want.tolist()
        rg = eng.forward_views(gv, poses, dmf.MODE_GOOD_POINTS, zd, bool(sparse), want=("ids",))
        for i in range(n):
            assert np.array_equal(r["depth"][i], g[f"{name}/{i}/depth/s{sparse}"].astype(np.int32))
            assert np.array_equal(r["ids"][i], g[f"{name}/{i}/points/s{sparse}/ids"])
            assert np.array_equal(rg["ids"][i], g[f"{name}/{i}/good/s{sparse}/ids"])
    mins = eng.forward_views(gv, poses, dmf.MODE_MINIMUM, 1, True, want=())["min_depth"]
    assert [int(m) for m in mins] == [int(g[f"{name}/{i}/min"]) for i in range(n)]
    rv = eng.reverse_views(gv, poses, fast=True)
    for i in range(n):
        assert np.array_equal(rv["ids"][i], g[f"{name}/{i}/reverse_fast/ids"])
        flags = np.zeros(len(gv.occupied_cells_), np.uint8)
        flags[dmf.bits_to_indices(rv["unoccluded"][i])] |= 1
        flags[dmf.bits_to_indices(rv["visibility"][i])] |= 2
        assert np.array_equal(flags, g[f"{name}/{i}/reverse_fast/flags"])
        assert np.array_equal(eng.reverseRayTrace(gv, poses[i], False)[1], g[f"{name}/{i}/reverse_slow/ids"])
        zb, cnt = eng.rayTraceVolume(gv, poses[i], return_depth=True)
        assert np.array_equal(zb, g[f"{name}/{i}/zbuffer"]) and cnt == int(g[f"{name}/{i}/zbuffer_n"])
    gv.clear_marks()
    eng.forward_views(gv, poses, dmf.MODE_CLASSIFY, zd, False, view_id0=1, want=())
    view, good = gv.marks()
    assert np.array_equal(view, g[f"{name}/classify/view"]) and np.array_equal(good, g[f"{name}/classify/good"])
    assert np.array_equal(dmf.greedySetCover(rv["visibility"], ctx), g[f"{name}/setcover"])


@pytest.mark.parametrize("name", ["S64", "S128-odd", "S128-clutter"])
def test_gpu_reproduces_carve_golden(dmf, ctx, name):
    """carve mode (DMF_FWD_CARVE) against tests/golden/golden_carve_v1.npz: the observed-voxel bit grid, both device paths"""
    g = np.load(GOLDEN)
    gc = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_carve_v1.npz"))
    K = g["K"]; H, W = (int(v) for v in g["HW"])
    sc = dmf.scenes.scene(name)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    poses = g[f"{name}/poses"]; zd = int(g[f"{name}/zdelta"])
    for fmt, skip in ((dmf.GRID_BYTE, True), (dmf.GRID_BYTE, False), (dmf.GRID_BIT, True)):
        eng = dmf.RayTracingEngine(dmf.Camera(K, H, W), ctx, fmt, skip_empty=skip)
        for sparse in (0, 1):
            gv._commit(ctx)
            ctx.clear_observed(); ctx.reset_counters()
            eng.forward_views(gv, poses[:1], dmf.MODE_POINTS, zd, bool(sparse), want=(), carve=True)
            assert np.array_equal(ctx.observed_words(), gc[f"{name}/s{sparse}/view0"])
            eng.forward_views(gv, poses[1:], dmf.MODE_POINTS, zd, bool(sparse), want=(), carve=True)
            assert np.array_equal(ctx.observed_words(), gc[f"{name}/s{sparse}/all_views"])
            assert ctx.counters()["inbounds"] == int(gc[f"{name}/s{sparse}/inbounds"])
