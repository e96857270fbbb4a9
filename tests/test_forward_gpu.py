"""GPU parity: forward per-pixel march (K1) through the C ABI vs the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): voxel hit sets, occupancy counts and visibility bitsets bit-exact; simulated
depths / points within 1e-5 relative (they are in fact compared bit-exact; the tolerance is stated for the record).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K_H, K_W = 480, 640
DEPTH_RTOL = 1e-5


def _scene_pair(dmf, oracle, ctx, name):
    sc = dmf.scenes.scene(name)
    ov = oracle.volume_from_scene(sc, flat=True)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds)
    gv.setVolumeSize(*sc.dims)
    gv.constructVolume()
    gv.integratePointCloud(sc.points, sc.normals)
    return sc, ov, gv


def _poses(dmf, sc, n_sphere=3):
    L = float(sc.bounds[1])
    ps = [dmf.scenes.pose_p1(L)[0]]
    ps += list(dmf.scenes.poses_sphere_lookat(L, 200)[:: max(1, 200 // n_sphere)][:n_sphere])
    ps += list(dmf.scenes.poses_position_camera(L, 40)[[7, 23]])
    return np.stack(ps)


def _check_forward(dmf, oracle, ctx, sc, ov, gv, poses, mode, zdelta, sparse, grid_format=0, H=K_H, W=K_W, K=None):
    """both kernels -- empty-space skipping (default) and brute force (DMF_FWD_NO_SKIP) -- against the oracle"""
    cache = {}
    a = _check_forward_one(dmf, oracle, ctx, sc, ov, gv, poses, mode, zdelta, sparse, grid_format, H, W, K, True, cache)
    b = _check_forward_one(dmf, oracle, ctx, sc, ov, gv, poses, mode, zdelta, sparse, grid_format, H, W, K, False, cache)
    assert b["skipped"] == 0
    return a


def _check_forward_one(dmf, oracle, ctx, sc, ov, gv, poses, mode, zdelta, sparse, grid_format, H, W, K, skip_empty, cache):
    K = dmf.scenes.REFERENCE_K if K is None else K
    eng = dmf.RayTracingEngine(dmf.Camera(K, H, W), ctx, grid_format, skip_empty=skip_empty)
    ctx.reset_counters()
    g = eng.forward_views(gv, poses, mode, zdelta, sparse)
    cnt = ctx.counters()
    occ = gv.occupied_cells_
    assert np.array_equal(occ, ov.occupied()), "occupied_cells_ differ"
    tot = dict(samples=0, inbounds=0, hits=0)
    for i, pose in enumerate(poses):
        if i not in cache:
            cache[i] = oracle.forward(ov, K, H, W, pose, mode, zdelta, sparse)
        o = cache[i]
        assert bool(g["found_any"][i]) == o["found_any"], f"view {i}: found_any"
        assert np.array_equal(g["depth"][i], o["depth"]), f"view {i}: first-hit depth image differs in {(g['depth'][i] != o['depth']).sum()} px"
        hit = o["depth"] >= 0
        assert np.array_equal(g["voxel"][i][hit], o["voxel"][hit]), f"view {i}: hit voxel ids differ"
        assert np.all(g["voxel"][i][~hit] == dmf.NO_VOXEL)
        # simulated depth cloud: stated tolerance 1e-5 relative; observed: bit-exact
        np.testing.assert_allclose(g["points"][i][hit], o["points"][hit], rtol=DEPTH_RTOL, atol=0)
        assert np.array_equal(g["points"][i][hit], o["points"][hit]), f"view {i}: points not bit-exact"
        # returned id list: exact, in the reference's discovery order
        assert np.array_equal(g["ids"][i], o["ids"]), f"view {i}: id list differs (gpu {len(g['ids'][i])}, oracle {len(o['ids'])})"
        # visibility bitset == set of returned ids
        vis_idx = dmf.bits_to_indices(g["visibility"][i])
        assert np.array_equal(np.sort(occ[vis_idx]), np.sort(o["ids"])), f"view {i}: visibility bitset != id set"
        for k in tot:
            tot[k] += o["counters"][k]
    assert cnt["samples"] == tot["samples"] and cnt["inbounds"] == tot["inbounds"] and cnt["hits"] == tot["hits"], (cnt, tot)
    return cnt


@pytest.mark.parametrize("name", ["S64", "S128", "S128-odd", "S128-clutter"])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("fmt", [0, 1])
def test_points_and_good_points_dense(dmf, oracle, ctx, name, mode, fmt):
    """fmt 0: bit grid + macro-cell clearance; fmt 1: per-voxel Chebyshev distance bytes"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, name)
    _check_forward(dmf, oracle, ctx, sc, ov, gv, _poses(dmf, sc), mode, sc.zdelta, False, grid_format=fmt)


def test_cameras_outside_the_volume(dmf, oracle, ctx):
    """rays that start outside the AABB, graze it, or cross it and leave: the out-of-bounds jump must not change a probe"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S128")
    L = float(sc.bounds[1])
    poses = np.stack([
        dmf.scenes.look_at([-0.4 * L, 0.5 * L, 0.5 * L], [0.5 * L, 0.5 * L, 0.5 * L]),      # outside, looking in
        dmf.scenes.look_at([0.5 * L, 0.5 * L, 1.3 * L], [0.5 * L, 0.45 * L, 0.5 * L]),
        dmf.scenes.look_at([-0.2 * L, -0.2 * L, 0.001], [L, L, 0.004]),                      # grazing the z = 0 face
        dmf.scenes.look_at([0.5 * L, 0.5 * L, 0.99 * L], [0.5 * L, 0.5 * L, 2 * L]),         # inside, leaving immediately
        dmf.scenes.look_at([1.2 * L, 0.3 * L, 0.5 * L], [1.2 * L, 0.9 * L, 0.5 * L]),        # parallel to a face, never inside
    ] + list(dmf.scenes.poses_position_camera(L, 60)[::12]))
    for fmt in (0, 1):
        for sparse, zd in ((False, sc.zdelta), (True, 3)):
            _check_forward(dmf, oracle, ctx, sc, ov, gv, poses, 0, zd, sparse, grid_format=fmt)


@pytest.mark.parametrize("mode", [0, 1])
def test_sparse_default_arguments(dmf, oracle, ctx, mode):
    """reference defaults: zdelta=10, sparse=true (pixel stride 5)"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S128")
    _check_forward(dmf, oracle, ctx, sc, ov, gv, _poses(dmf, sc), mode, 10, True)


def test_ragged_image_and_odd_zdelta(dmf, oracle, ctx):
    """image size not a multiple of the tile, z stride that does not divide the range"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S64")
    K = dmf.scenes.REFERENCE_K.copy()
    K[[0, 2, 4, 5]] *= 0.25
    for sparse in (False, True):
        _check_forward(dmf, oracle, ctx, sc, ov, gv, _poses(dmf, sc, 2), 0, 7, sparse, H=123, W=157, K=K)


def test_depth_u16_output(dmf, ctx):
    sc = dmf.scenes.scene("S64")
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    eng = dmf.RayTracingEngine(dmf.Camera(dmf.scenes.REFERENCE_K), ctx)
    poses = dmf.scenes.poses_sphere_lookat(1.024, 40)[::10]
    for sparse in (False, True):
        r = eng.forward_views(gv, poses, 0, 8, sparse, want=("depth", "depth16"))
        want = r["depth"].copy(); want[want < 0] = 0xFFFF
        assert np.array_equal(r["depth16"].astype(np.int32), want) and (r["depth"] >= 0).sum() > 1000


def test_empty_volume_and_camera_outside(dmf, oracle, ctx):
    sc = dmf.scenes.scene("S64")
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume()
    gv.integratePointCloud(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    eng = dmf.RayTracingEngine(dmf.Camera(dmf.scenes.REFERENCE_K), ctx)
    found, ids = eng.rayTraceAndGetPoints(gv, dmf.scenes.pose_p1(1.024)[0], 8, False)
    assert not found and len(ids) == 0
    # camera far outside the volume looking away: nothing in bounds
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S64")
    away = dmf.scenes.look_at([5.0, 5.0, 5.0], [9.0, 9.0, 9.0])
    _check_forward(dmf, oracle, ctx, sc, ov, gv, away[None], 0, 8, False)


def test_classify_and_mark(dmf, oracle, ctx):
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S128")
    poses = _poses(dmf, sc)
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    # sequential rayTraceAndClassify calls with increasing view ids == one batched call
    gv._commit(ctx); gv.clear_marks(); ov.clear_marks()
    eng.forward_views(gv, poses, dmf.MODE_CLASSIFY, sc.zdelta, False, view_id0=3, want=())
    for i, p in enumerate(poses):
        oracle.forward(ov, K, K_H, K_W, p, oracle.MODE_CLASSIFY, sc.zdelta, False, view=3 + i, want_pixels=False)
    gview, ggood = gv.marks()
    oview, ogood = ov.marks()
    assert np.array_equal(gview, oview), f"Voxel::view differs on {(gview != oview).sum()} voxels"
    assert np.array_equal(ggood, ogood), f"Voxel::good differs on {(ggood != ogood).sum()} voxels"
    assert gview.max() > 3 and ggood.sum() > 0
    # a second classify pass must not overwrite existing view ids
    eng.rayTraceAndClassify(gv, poses[1], sc.zdelta, 99, False)
    oracle.forward(ov, K, K_H, K_W, poses[1], oracle.MODE_CLASSIFY, sc.zdelta, False, view=99, want_pixels=False)
    assert np.array_equal(gv.marks()[0], ov.marks()[0])
    # rayTrace: view = 1 on every first-hit voxel
    gv.clear_marks(); ov.clear_marks()
    eng.rayTrace(gv, poses[2], sc.zdelta, True)
    oracle.forward(ov, K, K_H, K_W, poses[2], oracle.MODE_MARK, sc.zdelta, True, want_pixels=False)
    assert np.array_equal(gv.marks()[0], ov.marks()[0]) and gv.marks()[0].sum() > 0


@pytest.mark.parametrize("sparse,zdelta", [(True, 1), (False, 1), (True, 3)])
def test_minimum(dmf, oracle, ctx, sparse, zdelta):
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S128")
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    poses = _poses(dmf, sc)
    g = eng.forward_views(gv, poses, dmf.MODE_MINIMUM, zdelta, sparse, want=())["min_depth"]
    for i, p in enumerate(poses):
        o = oracle.forward(ov, K, K_H, K_W, p, oracle.MODE_MINIMUM, zdelta, sparse, want_pixels=False)
        assert int(g[i]) == o["min_depth"], (i, int(g[i]), o["min_depth"])
    assert eng.rayTraceAndGetMinimum(gv, dmf.scenes.look_at([5, 5, 5], [9, 9, 9]), 1, True) == -1


def test_config1_512_dyadic_single_view(dmf, oracle, ctx):
    """BASELINE.json configs[1]: one 640x480 view into the 512^3 grid (dyadic bounds => op-order independent)."""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S512")
    poses = np.stack([dmf.scenes.pose_p1(1.0)[0], dmf.scenes.poses_sphere_lookat(1.0, 64)[37]])
    _check_forward(dmf, oracle, ctx, sc, ov, gv, poses, 0, sc.zdelta, False, grid_format=1)
    cnt = _check_forward(dmf, oracle, ctx, sc, ov, gv, poses, 0, sc.zdelta, False)
    assert cnt["exact_div"] == 0 and cnt["f64_path"] == 0   # power-of-two voxel size, vmin = 0: the float quotient is exact
    assert cnt["skipped"] > 0.5 * cnt["inbounds"]            # most of the box scene is empty space


def test_full_size_properties_512(dmf, ctx):
    """Size-independent properties at full size (no oracle): byte grid == bit grid, batched == one-by-one,
    sparse lattice is a sub-sampling of dense, visibility popcount == unique hit voxels, idempotence."""
    sc = dmf.scenes.scene("S512")
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    K = dmf.scenes.REFERENCE_K
    poses = dmf.scenes.poses_sphere_lookat(1.0, 256)[::32]
    e_bit = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BIT)
    e_byte = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BYTE)
    e_brute = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BIT, skip_empty=False)
    ctx.reset_counters()
    a = e_bit.forward_views(gv, poses, 0, sc.zdelta, False)
    ca = ctx.counters(); ctx.reset_counters()
    b = e_byte.forward_views(gv, poses, 0, sc.zdelta, False)
    ctx.reset_counters()
    c = e_brute.forward_views(gv, poses, 0, sc.zdelta, False)
    cc = ctx.counters()
    for k in ("depth", "voxel", "points", "visibility"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
    # skipping changes neither the probes that count nor the in-bounds tally
    assert (ca["samples"], ca["inbounds"], ca["hits"]) == (cc["samples"], cc["inbounds"], cc["hits"])
    assert ca["skipped"] > 0 and cc["skipped"] == 0
    one = e_bit.forward_views(gv, poses[3:4], 0, sc.zdelta, False)
    assert np.array_equal(one["depth"][0], a["depth"][3]) and np.array_equal(one["ids"][0], a["ids"][3])
    again = e_bit.forward_views(gv, poses, 0, sc.zdelta, False)
    assert np.array_equal(again["depth"], a["depth"])
    sp = e_bit.forward_views(gv, poses, 0, sc.zdelta, True)
    assert np.array_equal(sp["depth"][:, ::5, ::5], a["depth"][:, ::5, ::5])
    mask = np.ones(sp["depth"].shape[1:], bool); mask[::5, ::5] = False
    assert np.all(sp["depth"][:, mask] == -1)
    for i in range(len(poses)):
        hit = a["depth"][i] >= 0
        uniq = np.unique(a["voxel"][i][hit])
        assert len(uniq) == len(a["ids"][i]) == int(np.unpackbits(a["visibility"][i].view(np.uint8)).sum())
        assert np.array_equal(np.sort(a["ids"][i]), uniq)
        assert np.isin(uniq, gv.occupied_cells_).all()


def test_config3_1024_grid_against_oracle(dmf, oracle, ctx):
    """BASELINE.json configs[3]: 1024^3 grid (0.977 mm voxels, zdelta = 1 mm).  Oracle on a 320x240 camera to keep it to seconds."""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S1024")
    assert sc.zdelta == 1 and len(ov.occupied()) == 565496
    K = dmf.scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.5
    poses = np.stack([dmf.scenes.pose_p1(1.0)[0], dmf.scenes.poses_helix(1.0, 40)[17]])
    for fmt in (1, 0):
        cnt = _check_forward(dmf, oracle, ctx, sc, ov, gv, poses, 0, 1, False, grid_format=fmt, H=240, W=320, K=K)
        assert cnt["f64_path"] == 0 and cnt["skipped"] > 0.8 * cnt["inbounds"]


def test_config4_1080p_camera_against_oracle(dmf, oracle, ctx):
    """BASELINE.json configs[4] camera: 1080x1920 with 3x intrinsics (2 073 600 rays: close to the 2^21 limit of the id-list keys)."""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S128")
    K = dmf.scenes.scaled_K(3.0)
    poses = dmf.scenes.poses_fibonacci(float(sc.bounds[1]), 50)[[21]]
    _check_forward(dmf, oracle, ctx, sc, ov, gv, poses, 0, sc.zdelta, False, grid_format=1, H=1080, W=1920, K=K)


def test_full_size_properties_1024_1080p(dmf, ctx):
    """configs[4] at full size without the oracle: 1080x1920 into 1024^3 -- distance bytes == bit grid == brute force,
    counters identical, id list == unique hit voxels == visibility popcount."""
    sc = dmf.scenes.scene("S1024")
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    K = dmf.scenes.scaled_K(3.0)
    poses = dmf.scenes.poses_fibonacci(1.0, 64)[[5, 40]]
    res, cnts = [], []
    for fmt, skip in ((dmf.GRID_BYTE, True), (dmf.GRID_BIT, True), (dmf.GRID_BIT, False)):
        eng = dmf.RayTracingEngine(dmf.Camera(K, 1080, 1920), ctx, fmt, skip_empty=skip)
        ctx.reset_counters()
        res.append(eng.forward_views(gv, poses, 0, 1, False, want=("depth", "voxel", "visibility", "ids")))
        cnts.append(ctx.counters())
    for r in res[1:]:
        for k in ("depth", "voxel", "visibility"):
            assert np.array_equal(res[0][k], r[k]), k
        assert all(np.array_equal(a, b) for a, b in zip(res[0]["ids"], r["ids"]))
    for c in cnts[1:]:
        assert (c["samples"], c["inbounds"], c["hits"]) == (cnts[0]["samples"], cnts[0]["inbounds"], cnts[0]["hits"])
    a = res[0]
    for i in range(len(poses)):
        hit = a["depth"][i] >= 0
        uniq = np.unique(a["voxel"][i][hit])
        assert hit.sum() > 100000
        assert len(uniq) == len(a["ids"][i]) == int(np.unpackbits(a["visibility"][i].view(np.uint8)).sum())
        assert np.array_equal(np.sort(a["ids"][i]), uniq)


def _aniso_scene(dmf, seed=3):
    """anisotropic, off-origin, non-dyadic volume (negative vmin, different voxel size per axis) with random blobs"""
    from dmf_b200.scenes import Scene
    rng = np.random.default_rng(seed)
    bounds = np.array([-0.31, 0.47, -0.22, 0.63, 0.11, 0.93], np.float64)
    dims = np.array([100, 120, 90], np.int32)
    cen = rng.uniform([-0.2, -0.1, 0.25], [0.35, 0.5, 0.8], size=(40, 3))
    pts, nrm = [], []
    for c in cen:
        r = rng.uniform(0.01, 0.05)
        d = rng.normal(size=(400, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts.append(c + r * d); nrm.append(d)
    pts = np.concatenate(pts).astype(np.float32); nrm = np.concatenate(nrm).astype(np.float32)
    return Scene("aniso", bounds, dims, pts, nrm)


def test_anisotropic_offset_volume(dmf, oracle, ctx):
    sc = _aniso_scene(dmf)
    ov = oracle.volume_from_scene(sc, flat=True)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    centre = np.array([0.08, 0.2, 0.52])
    poses = np.stack([dmf.scenes.look_at(centre + 0.45 * v, centre) for v in dmf.scenes.sphere_directions(6.0)[3::9][:5]]
                     + [dmf.scenes.look_at([0.9, 0.9, 1.4], centre), dmf.scenes.look_at([-0.30, -0.21, 0.12], centre)])
    for fmt in (1, 0):
        for mode, zd, sparse in ((0, 5, False), (1, 3, True)):
            cnt = _check_forward(dmf, oracle, ctx, sc, ov, gv, poses, mode, zd, sparse, grid_format=fmt)
    assert cnt["hits"] > 1000
