"""GPU parity: carve mode (DMF_FWD_CARVE, "occupied/free voxel marking") vs the CPU oracle.

Carve mode is an extension over the reference (which never records free space): oracle/dmf_oracle.hpp
(PixelOut::observed) is its definition -- every sample that passes validPoints and is visited by its ray up to and
including the first hit sets the bit of the voxel getVoxel puts it in.  Bar: the observed bit grid is bit-exact, for both
device implementations (carve_on_line behind k_forward_line, and the brute-force k_forward<.., CARVE>), and everything
else the call returns is unchanged by the flag.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K_H, K_W = 480, 640


def _scene_pair(dmf, oracle, ctx, name):
    sc = dmf.scenes.scene(name)
    ov = oracle.volume_from_scene(sc, flat=True)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    return sc, ov, gv


def _poses(dmf, sc):
    L = float(sc.bounds[1])
    ps = [dmf.scenes.pose_p1(L)[0]]
    ps += list(dmf.scenes.poses_sphere_lookat(L, 200)[::67][:3])
    ps += list(dmf.scenes.poses_position_camera(L, 40)[[7, 23]])
    ps += [dmf.scenes.look_at([-0.4 * L, 0.5 * L, 0.5 * L], [0.5 * L, 0.5 * L, 0.5 * L]),     # camera outside, looking in
           dmf.scenes.look_at([-0.2 * L, -0.2 * L, 0.001], [L, L, 0.004]),                     # grazing the z = 0 face
           dmf.scenes.look_at([0.5 * L, 0.5 * L, 0.5 * L], [0.9 * L, 0.7 * L, 0.8 * L])]       # camera inside the box
    return np.stack(ps)


def _popcount(words):
    return int(np.unpackbits(words.view(np.uint8)).sum())


def _oracle_observed(oracle, ov, K, H, W, poses, mode, zdelta, sparse):
    obs, tot = None, dict(samples=0, inbounds=0, hits=0)
    for p in poses:
        obs, c = oracle.forward_observed(ov, K, H, W, p, mode, zdelta, sparse, observed=obs)
        for k in tot:
            tot[k] += c[k]
    return obs, tot


def _gpu_observed(dmf, ctx, gv, K, H, W, poses, mode, zdelta, sparse, fmt, skip_empty):
    eng = dmf.RayTracingEngine(dmf.Camera(K, H, W), ctx, fmt, skip_empty=skip_empty)
    gv._commit(ctx)
    ctx.clear_observed()
    ctx.reset_counters()
    res = eng.forward_views(gv, poses, mode, zdelta, sparse, want=("depth", "visibility"), carve=True)
    return ctx.observed_words(), ctx.counters(), res


@pytest.mark.parametrize("name", ["S64", "S128", "S128-odd", "S128-clutter"])
def test_observed_grid_matches_oracle(dmf, oracle, ctx, name):
    """all three device paths: line-first + carve_on_line (byte grid), brute force on the byte grid, brute force on the bit grid"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, name)
    K = dmf.scenes.REFERENCE_K
    poses = _poses(dmf, sc)
    want, tot = _oracle_observed(oracle, ov, K, K_H, K_W, poses, 0, sc.zdelta, False)
    assert _popcount(want) > 1000
    plain = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BYTE).forward_views(gv, poses, 0, sc.zdelta, False, want=("depth", "visibility"))
    for fmt, skip in ((dmf.GRID_BYTE, True), (dmf.GRID_BYTE, False), (dmf.GRID_BIT, True)):
        got, cnt, res = _gpu_observed(dmf, ctx, gv, K, K_H, K_W, poses, 0, sc.zdelta, False, fmt, skip)
        diff = got ^ want
        assert not diff.any(), f"{name} fmt={fmt} skip={skip}: observed grid differs in {_popcount(diff)} voxels (gpu {_popcount(got)}, oracle {_popcount(want)})"
        assert cnt["samples"] == tot["samples"] and cnt["inbounds"] == tot["inbounds"] and cnt["hits"] == tot["hits"], (cnt, tot)
        # the flag changes nothing else
        assert np.array_equal(res["depth"], plain["depth"]) and np.array_equal(res["visibility"], plain["visibility"])
        c = ctx.observed_counts()
        assert c["observed"] == _popcount(want) and c["hit"] + c["free"] == c["observed"]


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_other_modes_sparse_and_odd_strides(dmf, oracle, ctx, mode):
    """GOOD_POINTS / CLASSIFY / MARK visit the same samples as POINTS; sparse lattice; z stride that does not divide the range"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S64")
    K = dmf.scenes.REFERENCE_K
    poses = _poses(dmf, sc)[:5]
    for sparse, zd in ((True, 10), (False, 7)):
        ov.clear_marks()
        want, tot = _oracle_observed(oracle, ov, K, K_H, K_W, poses, mode, zd, sparse)
        for fmt, skip in ((dmf.GRID_BYTE, True), (dmf.GRID_BYTE, False)):
            gv._commit(ctx); gv.clear_marks()
            got, cnt, _ = _gpu_observed(dmf, ctx, gv, K, K_H, K_W, poses, mode, zd, sparse, fmt, skip)
            assert np.array_equal(got, want), f"mode {mode} sparse={sparse} fmt={fmt} skip={skip}: {_popcount(got ^ want)} voxels differ"
            assert cnt["inbounds"] == tot["inbounds"]


def test_ragged_image_and_accumulation(dmf, oracle, ctx):
    """image not a multiple of the tile; the grid accumulates over calls until cleared; MINIMUM mode is refused"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S64")
    K = dmf.scenes.REFERENCE_K.copy()
    K[[0, 2, 4, 5]] *= 0.25
    H, W = 123, 157
    poses = _poses(dmf, sc)[:4]
    want, _ = _oracle_observed(oracle, ov, K, H, W, poses, 0, 5, False)
    eng = dmf.RayTracingEngine(dmf.Camera(K, H, W), ctx, dmf.GRID_BYTE)
    gv._commit(ctx)
    ctx.clear_observed()
    for p in poses:                                   # one view per call: same union
        eng.forward_views(gv, p, 0, 5, False, want=(), carve=True)
    assert np.array_equal(ctx.observed_words(), want)
    one, _ = _oracle_observed(oracle, ov, K, H, W, poses[:1], 0, 5, False)
    ctx.clear_observed()
    eng.forward_views(gv, poses[0], 0, 5, False, want=(), carve=True)
    assert np.array_equal(ctx.observed_words(), one)
    with pytest.raises(dmf.DmfError):
        eng.forward_views(gv, poses[0], dmf.MODE_MINIMUM, 1, True, want=(), carve=True)


def test_config1_512_observed(dmf, oracle, ctx):
    """BASELINE.json configs[1]: one 640x480 view into the 512^3 grid, carve mode, against the oracle"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S512")
    K = dmf.scenes.REFERENCE_K
    poses = np.stack([dmf.scenes.pose_p1(1.0)[0], dmf.scenes.poses_sphere_lookat(1.0, 64)[37]])
    want, tot = _oracle_observed(oracle, ov, K, K_H, K_W, poses, 0, sc.zdelta, False)
    for skip in (True, False):
        got, cnt, _ = _gpu_observed(dmf, ctx, gv, K, K_H, K_W, poses, 0, sc.zdelta, False, dmf.GRID_BYTE, skip)
        assert np.array_equal(got, want), f"skip={skip}: {_popcount(got ^ want)} voxels differ"
        assert cnt["inbounds"] == tot["inbounds"]


def test_full_size_properties(dmf, ctx):
    """No oracle, full size (S512, 640x480, 16 views): both device implementations agree bit for bit; observed & occupied
    is exactly the set of first-hit voxels; free + hit = observed <= in-bounds samples; a second pass changes nothing."""
    sc = dmf.scenes.scene("S512")
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    K = dmf.scenes.REFERENCE_K
    poses = dmf.scenes.poses_sphere_lookat(1.0, 256)[::16]
    line, cnt, res = _gpu_observed(dmf, ctx, gv, K, K_H, K_W, poses, 0, sc.zdelta, False, dmf.GRID_BYTE, True)
    counts = ctx.observed_counts()
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx, dmf.GRID_BYTE)
    eng.forward_views(gv, poses, 0, sc.zdelta, False, want=(), carve=True)          # idempotent
    assert np.array_equal(ctx.observed_words(), line)
    brute, cnt_b, _ = _gpu_observed(dmf, ctx, gv, K, K_H, K_W, poses, 0, sc.zdelta, False, dmf.GRID_BYTE, False)
    assert np.array_equal(line, brute), f"{_popcount(line ^ brute)} voxels differ between carve_on_line and the brute-force march"
    assert cnt["inbounds"] == cnt_b["inbounds"] and cnt["hits"] == cnt_b["hits"]
    assert counts["observed"] == _popcount(line) and counts["hit"] + counts["free"] == counts["observed"]
    assert counts["observed"] <= cnt["inbounds"]
    # observed & occupied == union over the views of the visibility bitsets (POINTS mode: every first-hit voxel is emitted)
    vis = np.bitwise_or.reduce(res["visibility"], axis=0)
    assert counts["hit"] == _popcount(vis)
    dims = np.asarray(sc.dims, np.int64) + 1
    ids = gv.occupied_cells_[dmf.bits_to_indices(vis)]
    x, y, z = (ids >> np.uint64(40)).astype(np.int64), ((ids >> np.uint64(20)) & np.uint64(0xFFFFF)).astype(np.int64), (ids & np.uint64(0xFFFFF)).astype(np.int64)
    idx = (x * dims[1] + y) * dims[2] + z
    assert np.all((line[idx >> 5] >> (idx & 31).astype(np.uint32)) & 1), "a first-hit voxel is not marked observed"


def test_anisotropic_offset_volume(dmf, oracle, ctx):
    """non-cubic voxels, bounds off the origin, non-dyadic sizes: the fixed-point line and the face test scale per axis"""
    from tests.test_forward_gpu import _aniso_scene
    sc = _aniso_scene(dmf)
    ov = oracle.volume_from_scene(sc, flat=True)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    K = dmf.scenes.REFERENCE_K
    centre = np.array([0.08, 0.2, 0.52])
    poses = np.stack([dmf.scenes.look_at(centre + 0.45 * v, centre) for v in dmf.scenes.sphere_directions(6.0)[3::9][:5]]
                     + [dmf.scenes.look_at([0.9, 0.9, 1.4], centre), dmf.scenes.look_at([-0.30, -0.21, 0.12], centre), dmf.scenes.look_at(centre, centre + [0.3, 0.1, 0.2])])
    for zd, sparse in ((5, False), (3, True)):
        want, tot = _oracle_observed(oracle, ov, K, K_H, K_W, poses, 0, zd, sparse)
        for fmt, skip in ((dmf.GRID_BYTE, True), (dmf.GRID_BYTE, False)):
            got, cnt, _ = _gpu_observed(dmf, ctx, gv, K, K_H, K_W, poses, 0, zd, sparse, fmt, skip)
            assert np.array_equal(got, want), f"zd={zd} sparse={sparse} fmt={fmt} skip={skip}: {_popcount(got ^ want)} voxels differ"
            assert cnt["samples"] == tot["samples"] and cnt["inbounds"] == tot["inbounds"] and cnt["hits"] == tot["hits"], (cnt, tot)
    assert _popcount(want) > 1000


def test_config3_1024_grid_observed(dmf, oracle, ctx):
    """BASELINE.json configs[3] grid: 1024^3 (0.977 mm voxels, zdelta = 1 mm, 990 samples per ray), 320x240 camera for the oracle"""
    sc, ov, gv = _scene_pair(dmf, oracle, ctx, "S1024")
    K = dmf.scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.5
    poses = np.stack([dmf.scenes.pose_p1(1.0)[0], dmf.scenes.poses_helix(1.0, 40)[17]])
    want, tot = _oracle_observed(oracle, ov, K, 240, 320, poses, 0, 1, False)
    got, cnt, _ = _gpu_observed(dmf, ctx, gv, K, 240, 320, poses, 0, 1, False, dmf.GRID_BYTE, True)
    assert np.array_equal(got, want), f"{_popcount(got ^ want)} voxels differ"
    assert cnt["inbounds"] == tot["inbounds"] and cnt["hits"] == tot["hits"]
