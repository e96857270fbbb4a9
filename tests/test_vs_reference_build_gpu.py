"""GPU tier: the CUDA path against the reference's OWN compiled headers (oracle/_ref/libref_dmf.so, built in the authoring
container from /root/reference and shipped to the GPU box as a binary) -- no restatement in between."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
H, W = 480, 640


@pytest.fixture(scope="module")
def ref():
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libref_dmf.so was not shipped")
    return ref_py


@pytest.mark.parametrize("name", ["S128", "S128-clutter", "S128d"])
def test_all_eight_methods_against_reference_source(dmf, ref, ctx, name):
    sc = dmf.scenes.scene(name)
    K = dmf.scenes.REFERENCE_K
    rv = ref.volume_from_scene(sc)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BYTE)
    L = float(sc.bounds[1])
    poses = np.stack([dmf.scenes.pose_p1(L)[0]] + list(dmf.scenes.poses_sphere_lookat(L, 200)[::70]) + [dmf.scenes.poses_position_camera(L, 40)[23]])
    gv._commit(ctx)
    assert np.array_equal(gv.occupied_cells_, rv.occupied())
    g_pts = eng.forward_views(gv, poses, dmf.MODE_POINTS, sc.zdelta, False, want=("ids",))
    g_good = eng.forward_views(gv, poses, dmf.MODE_GOOD_POINTS, 10, True, want=("ids",))
    g_min = eng.forward_views(gv, poses, dmf.MODE_MINIMUM, 1, True, want=())["min_depth"]
    g_rev = eng.reverse_views(gv, poses, fast=True, want=("ids",))
    total = 0
    for i, p in enumerate(poses):
        r = ref.forward(rv, K, H, W, p, 0, sc.zdelta, False)
        assert bool(g_pts["found_any"][i]) == r["found_any"] and np.array_equal(g_pts["ids"][i], r["ids"])
        r = ref.forward(rv, K, H, W, p, 1, 10, True)
        assert np.array_equal(g_good["ids"][i], r["ids"])
        assert int(g_min[i]) == ref.forward(rv, K, H, W, p, 4, 1, True)["min_depth"]
        r = ref.reverse(rv, K, H, W, p, fast=True)
        assert bool(g_rev["found_any"][i]) == r["found_any"] and np.array_equal(g_rev["ids"][i], r["ids"])
        assert np.array_equal(eng.reverseRayTrace(gv, p, False)[1], ref.reverse(rv, K, H, W, p, fast=False)["ids"])
        total += len(g_pts["ids"][i]) + len(g_rev["ids"][i])
    assert total > 2000
    # mutating routines: Voxel::view / Voxel::good
    gv.clear_marks(); rv.clear_marks()
    eng.forward_views(gv, poses, dmf.MODE_CLASSIFY, sc.zdelta, False, view_id0=1, want=())
    for i, p in enumerate(poses):
        ref.forward(rv, K, H, W, p, 2, sc.zdelta, False, view=1 + i)
    assert all(np.array_equal(a, b) for a, b in zip(gv.marks(), rv.marks()))
    gv.clear_marks(); rv.clear_marks()
    eng.rayTrace(gv, poses[1], 10, True); ref.forward(rv, K, H, W, poses[1], 3, 10, True)
    eng.reverseRayTraceFast(gv, poses[2], True); ref.reverse(rv, K, H, W, poses[2], fast=True, viz=True)
    eng.rayTraceVolume(gv, poses[0]); ref.zbuffer(rv, K, H, W, poses[0])
    assert all(np.array_equal(a, b) for a, b in zip(gv.marks(), rv.marks())) and gv.marks()[0].sum() > 0
