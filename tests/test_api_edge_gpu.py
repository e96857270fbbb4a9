"""GPU tier: error behaviour and edge cases of the C ABI (no reference equivalent: the reference prints and continues or
hits UB; the library must fail cleanly with a message and never crash)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _vol(dmf, ctx, name="S64"):
    sc = dmf.scenes.scene(name)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    return sc, gv


def test_calls_before_setup_fail_cleanly(dmf):
    from dmf_b200._lib import ForwardOut, ForwardParams
    c = dmf.Context(0)
    try:
        p = ForwardParams(0, 8, 0, 1, 0, 0)
        o = ForwardOut()
        pose = dmf.scenes.pose_p1(1.0)[0]
        rc = c.lib.dmf_forward(c.h, C.byref(p), pose.ctypes.data_as(C.POINTER(C.c_float)), 1, C.byref(o))
        assert rc != 0 and b"dmf_set_camera" in c.lib.dmf_last_error()
        K = dmf.scenes.REFERENCE_K
        assert c.lib.dmf_set_camera(c.h, K.ctypes.data_as(C.POINTER(C.c_float)), 480, 640) == 0
        rc = c.lib.dmf_forward(c.h, C.byref(p), pose.ctypes.data_as(C.POINTER(C.c_float)), 1, C.byref(o))
        assert rc != 0 and b"no volume" in c.lib.dmf_last_error()
        assert c.lib.dmf_set_camera(c.h, K.ctypes.data_as(C.POINTER(C.c_float)), 0, 640) != 0
    finally:
        c.close()


def test_bad_parameters_and_capacity(dmf, ctx):
    from dmf_b200._lib import ForwardOut, ForwardParams
    sc, gv = _vol(dmf, ctx)
    eng = dmf.RayTracingEngine(dmf.Camera(dmf.scenes.REFERENCE_K), ctx)
    eng._prepare(gv)
    pose = np.ascontiguousarray(dmf.scenes.pose_p1(1.024)[0])
    fp = pose.ctypes.data_as(C.POINTER(C.c_float))
    o = ForwardOut()
    for bad in (ForwardParams(7, 8, 0, 1, 0, 0), ForwardParams(0, 0, 0, 1, 0, 0), ForwardParams(0, 8, 0, 1, 5, 0)):
        assert ctx.lib.dmf_forward(ctx.h, C.byref(bad), fp, 1, C.byref(o)) != 0
    # id list larger than the caller's buffer: clean failure, message says how many are needed
    ids = np.zeros(4, np.uint64); offs = np.zeros(2, np.int64)
    o.ids, o.ids_offsets, o.ids_capacity = ids.ctypes.data, offs.ctypes.data, 4
    p = ForwardParams(0, 8, 0, 1, 0, 0)
    assert ctx.lib.dmf_forward(ctx.h, C.byref(p), fp, 1, C.byref(o)) != 0
    assert b"ids_capacity" in ctx.lib.dmf_last_error()
    # MINIMUM without an output buffer: computed and discarded on the host-buffer path, refused on the device-buffer path
    o2 = ForwardOut()
    assert ctx.lib.dmf_forward(ctx.h, C.byref(ForwardParams(4, 1, 1, 1, 0, 0)), fp, 1, C.byref(o2)) == 0
    # zero views is a no-op
    assert ctx.lib.dmf_forward(ctx.h, C.byref(p), fp, 0, C.byref(ForwardOut())) == 0
    # the context still works afterwards
    found, got = eng.rayTraceAndGetPoints(gv, pose, 8, False)
    assert found and len(got) > 100


def test_upload_volume_validation(dmf, ctx):
    lib = ctx.lib
    b = np.array([0, 1, 0, 1, 0, 1], np.float64); d = np.full(3, 1 / 8, np.float64); dim = np.array([8, 8, 8], np.int32)
    dp, ip, u64p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint64)
    dup = np.array([(1 << 40) | (2 << 20) | 3] * 2, np.uint64)
    assert lib.dmf_upload_volume(ctx.h, b.ctypes.data_as(dp), d.ctypes.data_as(dp), dim.ctypes.data_as(ip), dup.ctypes.data_as(u64p), 2, None, None) != 0
    assert b"duplicate" in lib.dmf_last_error()
    out = np.array([(9 << 40)], np.uint64)
    assert lib.dmf_upload_volume(ctx.h, b.ctypes.data_as(dp), d.ctypes.data_as(dp), dim.ctypes.data_as(ip), out.ctypes.data_as(u64p), 1, None, None) != 0
    big = np.array([4096, 8, 8], np.int32)
    assert lib.dmf_upload_volume(ctx.h, b.ctypes.data_as(dp), d.ctypes.data_as(dp), big.ctypes.data_as(ip), None, 0, None, None) != 0
    ctx._volume_token = None   # the context's volume is undefined after failed uploads: force the next test to re-upload


def test_nan_and_degenerate_poses_do_not_crash(dmf, ctx):
    sc, gv = _vol(dmf, ctx)
    eng = dmf.RayTracingEngine(dmf.Camera(dmf.scenes.REFERENCE_K), ctx)
    nan_pose = np.full(12, np.nan, np.float32)
    zero_pose = np.zeros(12, np.float32)
    huge = dmf.scenes.pose_p1(1.024)[0].copy(); huge[[3, 7, 11]] = 1e30
    for fmt in (dmf.GRID_BYTE, dmf.GRID_BIT):
        eng.grid_format = fmt
        r = eng.forward_views(gv, np.stack([nan_pose, zero_pose, huge]), 0, 8, False, want=("depth",))
        assert (r["depth"] == -1).all() and not r["found_any"].any()
    rv = eng.reverse_views(gv, np.stack([nan_pose, zero_pose]), fast=True, want=("visibility",))
    assert not rv["found_any"][0]
    ctx.synchronize()


def test_partially_non_finite_poses_match_oracle(dmf, oracle, ctx):
    """one NaN / inf entry in an otherwise sane pose: fmaxf-style bounds drop NaNs, so the line-first march must be
    told explicitly not to follow such a view (regression: garbage line addresses); results still equal the oracle's"""
    sc, gv = _vol(dmf, ctx)
    ov = oracle.volume_from_scene(sc, flat=True)
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    base = dmf.scenes.pose_p1(float(sc.bounds[1]))[0]
    poses = [base.copy()]
    for idx, val in ((3, np.nan), (0, np.nan), (10, np.inf), (7, -np.inf), (5, np.nan)):
        p = base.copy(); p[idx] = val; poses.append(p)
    poses = np.stack(poses)
    for fmt in (dmf.GRID_BYTE, dmf.GRID_BIT):
        eng.grid_format = fmt
        r = eng.forward_views(gv, poses, 0, sc.zdelta, False, want=("depth", "visibility"))
        for i, p in enumerate(poses):
            o = oracle.forward(ov, K, 480, 640, p, oracle.MODE_POINTS, sc.zdelta, False)
            assert np.array_equal(r["depth"][i], o["depth"]), (fmt, i)
            assert bool(r["found_any"][i]) == o["found_any"]
    ctx.synchronize()


def test_two_contexts_and_reupload(dmf, oracle, ctx):
    sc_a, gv_a = _vol(dmf, ctx, "S64")
    c2 = dmf.Context(0)
    try:
        sc_b, gv_b = _vol(dmf, c2, "S128")
        K = dmf.scenes.REFERENCE_K
        e1 = dmf.RayTracingEngine(dmf.Camera(K), ctx); e2 = dmf.RayTracingEngine(dmf.Camera(K), c2)
        pose = dmf.scenes.pose_p1(1.024)[0]
        ids_a = e1.rayTraceAndGetPoints(gv_a, pose, 8, False)[1]
        ids_b = e2.rayTraceAndGetPoints(gv_b, pose, 8, False)[1]
        assert np.array_equal(ids_a, oracle.forward(oracle.volume_from_scene(sc_a), K, 480, 640, pose, 0, 8, False, want_pixels=False)["ids"])
        assert np.array_equal(ids_b, oracle.forward(oracle.volume_from_scene(sc_b), K, 480, 640, pose, 0, 8, False, want_pixels=False)["ids"])
        # re-using context 1 for another volume re-uploads transparently
        ids_b1 = e1.rayTraceAndGetPoints(gv_b, pose, 8, False)[1]
        assert np.array_equal(ids_b1, ids_b)
        assert np.array_equal(e1.rayTraceAndGetPoints(gv_a, pose, 8, False)[1], ids_a)
    finally:
        c2.close()


def test_many_small_views_are_chunked(dmf, oracle, ctx):
    """5000 views of a 16x12 camera: more than one 4096-view chunk through the double-buffered host path"""
    sc, gv = _vol(dmf, ctx)
    K = dmf.scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.025
    eng = dmf.RayTracingEngine(dmf.Camera(K, 12, 16), ctx)
    base = dmf.scenes.poses_sphere_lookat(1.024, 1000)
    poses = np.tile(base, (5, 1))
    r = eng.forward_views(gv, poses, 0, 8, False, want=("depth", "visibility"))
    assert np.array_equal(r["depth"][:1000], r["depth"][4000:]) and np.array_equal(r["visibility"][:1000], r["visibility"][4000:])
    ov = oracle.volume_from_scene(sc)
    for i in (0, 777, 4999):
        assert np.array_equal(r["depth"][i], oracle.forward(ov, K, 12, 16, poses[i], 0, 8, False)["depth"])


@pytest.mark.parametrize("dims,L", [((32, 32, 32), 1.0), ((93, 40, 57), 0.937), ((128, 128, 128), 1.024)])
def test_integrate_point_cloud_on_gpu(dmf, oracle, ctx, dims, L):
    """K0: integratePointCloud on the device == the oracle's (Volume.hpp:199-228): first-insertion order of
    occupied_cells_, per-voxel normal lists in point order; ragged clouds with repeats, out-of-bounds and boundary points"""
    rng = np.random.default_rng(5)
    n = 60000
    pts = rng.uniform(-0.05 * L, 1.05 * L, size=(n, 3)).astype(np.float32)
    pts[:2000] = pts[rng.integers(2000, n, 2000)]                        # exact repeats -> several normals per voxel
    pts[100] = [0.0, 0.5 * L, 0.5 * L]; pts[101] = [L, 0.5 * L, 0.5 * L]  # on the faces: validPoints is strict
    pts[102] = [np.float32(L) - np.float32(1e-7), 0.5 * L, 0.5 * L]
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    bounds = [0, L, 0, L, 0, L]
    ov = oracle.Volume(bounds, dims)
    ov.integrate(pts, nrm)
    off_o, nrm_o = ov.normals_csr()
    for on_gpu in (True, False):
        gv = dmf.VoxelVolume(ctx, integrate_on_gpu=on_gpu)
        gv.setDimensions(*bounds); gv.setVolumeSize(*dims); gv.constructVolume(); gv.integratePointCloud(pts, nrm)
        gv._commit(ctx)
        assert (gv.xdim_, gv.ydim_, gv.zdim_) == tuple(int(d) for d in ov.dims)
        assert np.array_equal(gv.occupied_cells_, ov.occupied()), on_gpu
        off_g, nrm_g = gv.normals_csr()
        assert np.array_equal(off_g, off_o) and np.array_equal(nrm_g, nrm_o), on_gpu
    # empty cloud and a cloud entirely outside
    for cloud in (np.zeros((0, 3), np.float32), np.full((10, 3), 5.0, np.float32)):
        gv = dmf.VoxelVolume(ctx, integrate_on_gpu=True)
        gv.setDimensions(*bounds); gv.setVolumeSize(*dims); gv.constructVolume(); gv.integratePointCloud(cloud, np.zeros_like(cloud))
        gv._commit(ctx)
        assert len(gv.occupied_cells_) == 0


def test_gpu_integrated_volume_marches_identically(dmf, ctx):
    sc = dmf.scenes.scene("S128-clutter")
    K = dmf.scenes.REFERENCE_K
    poses = dmf.scenes.poses_sphere_lookat(1.024, 60)[::20]
    res = []
    for on_gpu in (False, True):
        gv = dmf.VoxelVolume(ctx, integrate_on_gpu=on_gpu)
        gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
        eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
        res.append((eng.forward_views(gv, poses, dmf.MODE_GOOD_POINTS, sc.zdelta, False, want=("depth", "visibility")), eng.reverse_views(gv, poses, want=("visibility",))))
    assert np.array_equal(res[0][0]["depth"], res[1][0]["depth"]) and np.array_equal(res[0][0]["visibility"], res[1][0]["visibility"])
    assert np.array_equal(res[0][1]["visibility"], res[1][1]["visibility"])


def test_round2_api_additions(dmf, ctx):
    """DMF_GRID_AUTO, DMF_FWD_NO_COUNTERS, the n argument of the marks calls, upload validation, dmf_volume_prepare_ms"""
    from dmf_b200._lib import ForwardOut, ForwardParams
    sc, gv = _vol(dmf, ctx, "S128")
    K = dmf.scenes.REFERENCE_K
    poses = dmf.scenes.poses_sphere_lookat(1.024, 40)[3::9]
    res = {}
    for fmt in (dmf.GRID_BIT, dmf.GRID_BYTE):
        res[fmt] = dmf.RayTracingEngine(dmf.Camera(K), ctx, fmt).forward_views(gv, poses, dmf.MODE_POINTS, sc.zdelta, False)
    # AUTO: bit grid on the first call after an upload, distance bytes once they exist / from the second call on -- same results
    gv._dirty = True                                              # force a fresh upload: no distance bytes yet
    auto = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_AUTO)
    for _ in range(3):
        r = auto.forward_views(gv, poses, dmf.MODE_POINTS, sc.zdelta, False)
        for key in ("depth", "voxel", "visibility", "found_any"):
            assert np.array_equal(r[key], res[dmf.GRID_BIT][key]) and np.array_equal(r[key], res[dmf.GRID_BYTE][key]), key
        assert all(np.array_equal(a, b) for a, b in zip(r["ids"], res[dmf.GRID_BYTE]["ids"]))
    build_ms, bytes_ms = C.c_float(), C.c_float()
    assert ctx.lib.dmf_volume_prepare_ms(ctx.h, C.byref(build_ms), C.byref(bytes_ms)) == 0
    assert 0 < build_ms.value < 50 and 0 < bytes_ms.value < 200, (build_ms.value, bytes_ms.value)   # AUTO has built the bytes by now
    # NO_COUNTERS: same outputs, counters untouched
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BYTE)
    eng._prepare(gv)
    ctx.reset_counters()
    p = ForwardParams(dmf.MODE_POINTS, sc.zdelta, 0, 1, dmf.GRID_BYTE, dmf.FWD_NO_COUNTERS)
    depth = np.zeros((len(poses), 480, 640), np.int32)
    o = ForwardOut(); o.depth_mm = depth.ctypes.data
    pp = np.ascontiguousarray(poses)
    assert ctx.lib.dmf_forward(ctx.h, C.byref(p), pp.ctypes.data_as(C.POINTER(C.c_float)), len(pp), C.byref(o)) == 0
    assert np.array_equal(depth, res[dmf.GRID_BYTE]["depth"])
    c = ctx.counters()
    assert c["samples"] == 0 and c["inbounds"] == 0 and c["hits"] == 0
    # marks: the caller's n must be the uploaded volume's
    n = len(gv.occupied_cells_)
    view, good = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.uint8)
    assert ctx.lib.dmf_download_marks(ctx.h, view.ctypes.data_as(C.POINTER(C.c_int32)), good.ctypes.data_as(C.POINTER(C.c_uint8)), n + 1) != 0
    assert b"voxels" in ctx.lib.dmf_last_error()
    assert ctx.lib.dmf_download_marks(ctx.h, view.ctypes.data_as(C.POINTER(C.c_int32)), good.ctypes.data_as(C.POINTER(C.c_uint8)), n) == 0
    # upload validation: a dim too small for bounds/delta, and a CSR that is not monotone
    c2 = dmf.Context(0)
    try:
        bounds = np.array([0, 1, 0, 1, 0, 1], np.float64); delta = np.full(3, 1 / 64, np.float64)
        ids = np.array([dmf.VoxelVolume.getHashId(1, 2, 3), dmf.VoxelVolume.getHashId(4, 5, 6)], np.uint64)
        def up(dim, noff=None, nrm=None):
            d = np.array(dim, np.int32)
            return c2.lib.dmf_upload_volume(c2.h, bounds.ctypes.data_as(C.POINTER(C.c_double)), delta.ctypes.data_as(C.POINTER(C.c_double)), d.ctypes.data_as(C.POINTER(C.c_int)),
                                            ids.ctypes.data_as(C.POINTER(C.c_uint64)), len(ids), None if noff is None else noff.ctypes.data_as(C.POINTER(C.c_uint32)),
                                            None if nrm is None else nrm.ctypes.data_as(C.POINTER(C.c_float)))
        assert up([64, 64, 64]) == 0
        assert up([60, 64, 64]) != 0 and b"inconsistent" in c2.lib.dmf_last_error()
        assert up([64, 64, 64], np.array([0, 2, 1], np.uint32), np.zeros(6, np.float32)) != 0 and b"monotone" in c2.lib.dmf_last_error()
        assert up([64, 64, 64], np.array([1, 1, 2], np.uint32), np.zeros(6, np.float32)) != 0
        assert up([64, 64, 64], np.array([0, 1, 2], np.uint32), np.zeros(6, np.float32)) == 0
    finally:
        c2.close()


def test_device_integration_equals_host_integration(dmf, ctx):
    """dmf_volume_from_points_gpu (K0 -> structures built where the ids lie, no host round trip) == dmf_volume_from_points"""
    sc = dmf.scenes.scene("S128-clutter")
    K = dmf.scenes.REFERENCE_K
    poses = dmf.scenes.poses_sphere_lookat(1.024, 40)[5::11]
    out = []
    for on_gpu in (False, True):
        gv = dmf.VoxelVolume(ctx, integrate_on_gpu=on_gpu)
        gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
        eng = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BYTE)
        r = eng.forward_views(gv, poses, dmf.MODE_GOOD_POINTS, sc.zdelta, False, want=("depth", "visibility", "ids"))
        rv = eng.reverse_views(gv, poses, fast=True)
        off, nrm = gv.normals_csr()
        out.append((gv.occupied_cells_.copy(), off, nrm, r, rv))
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for key in ("depth", "visibility", "found_any"):
        assert np.array_equal(a[3][key], b[3][key])
    assert all(np.array_equal(x, y) for x, y in zip(a[3]["ids"], b[3]["ids"])) and all(np.array_equal(x, y) for x, y in zip(a[4]["ids"], b[4]["ids"]))


def test_graph_replay_survives_intervening_calls(dmf):
    """The single-view id-list call is replayed as a captured CUDA graph from its second identical call on.  A call with other
    parameters in between (other z tables: MINIMUM starts at 5 mm and strides 10 px) must neither break the capture -- a table
    rebuild synchronises the device, which cannot be captured -- nor change a result."""
    c = dmf.Context(0)
    try:
        sc, gv = _vol(dmf, c)
        pose = np.ascontiguousarray(dmf.scenes.pose_p1(1.024)[0])
        for grid in (dmf.GRID_BYTE, dmf.GRID_BIT):
            eng = dmf.RayTracingEngine(dmf.Camera(dmf.scenes.REFERENCE_K), c, grid)
            found0, ids0 = eng.rayTraceAndGetPoints(gv, pose, 8, False)
            for _ in range(3):
                zmin = eng.rayTraceAndGetMinimum(gv, pose, 1, True)          # replaces the tables between two identical id-list calls
                found, ids = eng.rayTraceAndGetPoints(gv, pose, 8, False)
                assert found == found0 and np.array_equal(ids, ids0)
                found, ids = eng.rayTraceAndGetPoints(gv, pose, 8, False)    # (and two in a row: candidate -> capture -> replay)
                assert found == found0 and np.array_equal(ids, ids0)
            assert zmin > 0
            v0, r0 = eng.reverseRayTraceFast(gv, pose, False)
            for _ in range(3):
                eng.rayTraceAndGetMinimum(gv, pose, 1, True)
                v, r = eng.reverseRayTraceFast(gv, pose, False)
                assert v == v0 and np.array_equal(r, r0)
    finally:
        c.close()
