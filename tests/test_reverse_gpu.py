"""GPU parity: reverse per-voxel march (K2), z-buffer (K3), set cover (K6) and OR combine (K5) vs the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H, W = 480, 640


def _pair(dmf, oracle, ctx, name):
    sc = dmf.scenes.scene(name)
    ov = oracle.volume_from_scene(sc, flat=True)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    return sc, ov, gv


def _poses(dmf, sc):
    L = float(sc.bounds[1])
    return np.stack([dmf.scenes.pose_p1(L)[0]] + list(dmf.scenes.poses_sphere_lookat(L, 300)[::60]) + list(dmf.scenes.poses_position_camera(L, 40)[[7, 23]]))


@pytest.fixture(autouse=True)
def _default_reverse_format(dmf, ctx):
    yield
    ctx.set_reverse_format(dmf.GRID_BYTE)


def test_div1000_exhaustive(ctx):
    """every float bit pattern: the division-free x/1000 of the reverse march == IEEE division wherever the kernels use it"""
    long_bad, short_bad, hi_bits, lo_bits, short_bad_above = ctx.selftest_div1000()
    assert short_bad_above == 0                         # one-correction form exact for all finite |a| > 2^-101
    assert hi_bits < 0x0d000000                         # every finite failure of the two-step form lies below 2^-101 too
    assert long_bad < 100000 and short_bad < 100000     # (subnormal quotients, -0, +-inf only)


@pytest.mark.parametrize("name", ["S64", "S128", "S128-odd", "S128-clutter", "S256"])
@pytest.mark.parametrize("fmt", [1, 0])
def test_reverse_fast(dmf, oracle, ctx, name, fmt):
    sc, ov, gv = _pair(dmf, oracle, ctx, name)
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    ctx.set_reverse_format(fmt)
    poses = _poses(dmf, sc)
    ctx.reset_counters()
    g = eng.reverse_views(gv, poses, fast=True, viz=False)
    cnt = ctx.counters()
    assert (cnt["skipped"] > 0) == (fmt == 1)
    occ = gv.occupied_cells_
    tot = dict(samples=0, inbounds=0, hits=0)
    n_vis = 0
    for i, p in enumerate(poses):
        o = oracle.reverse(ov, K, H, W, p, fast=True)
        assert bool(g["found_any"][i]) == o["found_any"]
        assert np.array_equal(g["ids"][i], o["ids"]), f"view {i}: emitted ids differ ({len(g['ids'][i])} vs {len(o['ids'])})"
        un = np.zeros(len(occ), np.uint8); un[dmf.bits_to_indices(g["unoccluded"][i])] = 1
        em = np.zeros(len(occ), np.uint8); em[dmf.bits_to_indices(g["visibility"][i])] = 1
        assert np.array_equal(un, o["flags"] & 1), f"view {i}: unoccluded set differs on {(un != (o['flags'] & 1)).sum()} voxels"
        assert np.array_equal(em, (o["flags"] >> 1) & 1), f"view {i}: emitted set differs"
        n_vis += int(un.sum())
        for k in tot:
            tot[k] += o["counters"][k]
    assert n_vis > 0
    assert (cnt["samples"], cnt["inbounds"], cnt["hits"]) == (tot["samples"], tot["inbounds"], tot["hits"]), (cnt, tot)
    assert cnt["runaway"] == 0


def test_reverse_fast_viz_marks(dmf, oracle, ctx):
    sc, ov, gv = _pair(dmf, oracle, ctx, "S128")
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    gv._commit(ctx); gv.clear_marks(); ov.clear_marks()
    for p in _poses(dmf, sc)[:3]:
        found, ids = eng.reverseRayTraceFast(gv, p, True)
        o = oracle.reverse(ov, K, H, W, p, fast=True, viz=True)
        assert found == o["found_any"] and np.array_equal(ids, o["ids"])
    assert np.array_equal(gv.marks()[0], ov.marks()[0]) and np.array_equal(gv.marks()[1], ov.marks()[1])
    assert gv.marks()[0].sum() > 0


@pytest.mark.parametrize("name", ["S64", "S128", "S256"])
@pytest.mark.parametrize("fmt", [1, 0])
def test_reverse_whole_grid(dmf, oracle, ctx, name, fmt):
    """reverseRayTrace: float-accumulated whole-grid scan.  S256 is dyadic (scan visits each voxel once);
    S64/S128 (8/16 mm voxels on a 1.024 m cube) exercise the drifting float loop positions."""
    sc, ov, gv = _pair(dmf, oracle, ctx, name)
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    ctx.set_reverse_format(fmt)
    gv._commit(ctx); gv.clear_marks(); ov.clear_marks()
    any_found = False
    for p in _poses(dmf, sc)[:4]:
        found, ids = eng.reverseRayTrace(gv, p, True)
        o = oracle.reverse(ov, K, H, W, p, fast=False, viz=True)
        assert found == o["found_any"]
        assert np.array_equal(ids, o["ids"]), f"{len(ids)} vs {len(o['ids'])}"
        any_found |= found
    assert np.array_equal(gv.marks()[0], ov.marks()[0]) and np.array_equal(gv.marks()[1], ov.marks()[1])
    if name == "S256":
        assert any_found


@pytest.mark.parametrize("name", ["S64", "S128", "S256"])
def test_zbuffer(dmf, oracle, ctx, name):
    sc, ov, gv = _pair(dmf, oracle, ctx, name)
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    for p in _poses(dmf, sc)[:4]:
        gv._commit(ctx); gv.clear_marks(); ov.clear_marks()
        gd, gn = eng.rayTraceVolume(gv, p, return_depth=True)
        od, on = oracle.zbuffer(ov, K, H, W, p)
        assert gn == on
        assert np.array_equal(gd, od), f"z-buffer differs in {(gd != od).sum()} px"
        assert np.array_equal(gv.marks()[0], ov.marks()[0])


def test_affine_inverse_matches_restated_eigen(dmf, oracle, ctx):
    """k_invert_poses is exercised through deProjectPoint in every reverse test; here: non-rotation linear parts
    (Algorithms::positionCamera poses) give identical visible sets, which requires a bit-identical inverse."""
    sc, ov, gv = _pair(dmf, oracle, ctx, "S128")
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    poses = dmf.scenes.poses_position_camera(float(sc.bounds[1]), 64, standoff=0.3)[::8]
    g = eng.reverse_views(gv, poses, fast=True)
    for i, p in enumerate(poses):
        assert np.array_equal(g["ids"][i], oracle.reverse(ov, K, H, W, p, fast=True)["ids"])


def test_greedy_set_cover(dmf, oracle, ctx):
    sc, ov, gv = _pair(dmf, oracle, ctx, "S128")
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    poses = dmf.scenes.poses_sphere_lookat(float(sc.bounds[1]), 96)
    g = eng.reverse_views(gv, poses, fast=True, want=("visibility", "ids"))
    # the reference pipeline: per-view sorted id lists -> greedySetCover (tests/SetCover.cpp:218-240)
    sets = [np.sort(ids) for ids in g["ids"]]
    want = oracle.greedy_set_cover(sets)
    got = dmf.greedySetCover(g["visibility"], ctx)
    assert len(want) > 1
    assert np.array_equal(got, want), (got, want)
    # degenerate inputs: all empty, and all below the 5-point threshold
    assert len(dmf.greedySetCover(np.zeros((4, 8), np.uint64), ctx)) == 0
    small = np.zeros((3, 2), np.uint64); small[:, 0] = [0b1111, 0b0111, 0b1]
    assert len(dmf.greedySetCover(small, ctx)) == 0
    ties = np.zeros((3, 1), np.uint64); ties[:, 0] = [0b11111, 0b1111100000, 0b111110000000000]
    assert list(dmf.greedySetCover(ties, ctx)) == [0, 1, 2]   # equal gains: lowest index first


def test_or_reduce_dev(dmf, ctx):
    import torch
    rng = np.random.default_rng(7)
    src = rng.integers(0, 2**63, size=(5, 1000), dtype=np.int64)
    dst = rng.integers(0, 2**63, size=1000, dtype=np.int64)
    d_src = torch.from_numpy(src).cuda(); d_dst = torch.from_numpy(dst).cuda()
    torch.cuda.synchronize()
    from dmf_b200._lib import check
    check(ctx.lib.dmf_or_reduce_dev(ctx.h, C.c_void_p(d_dst.data_ptr()), C.c_void_p(d_src.data_ptr()), 5, 1000, None))
    ctx.synchronize()
    assert np.array_equal(d_dst.cpu().numpy(), np.bitwise_or.reduce(src, axis=0) | dst)


@pytest.mark.parametrize("fmt", [1, 0])
def test_will_collide_batch(dmf, oracle, ctx, fmt):
    """willCollide (tests/CameraPathGen.cpp:128-156 and the unguarded copies) on a batch of segments"""
    sc, ov, gv = _pair(dmf, oracle, ctx, "S128-clutter")
    ctx.set_reverse_format(fmt)
    L = float(sc.bounds[1])
    rng = np.random.default_rng(11)
    a = rng.uniform(-0.3 * L, 1.3 * L, size=(96, 3)).astype(np.float32)
    b = rng.uniform(-0.3 * L, 1.3 * L, size=(96, 3)).astype(np.float32)
    # a few structured cases: through the box, along an edge of the volume, zero length, fully outside
    a[:6] = [[0.1 * L, 0.5 * L, 0.5 * L], [0.0, 0.0, 0.0], [0.5 * L, 0.5 * L, 0.5 * L], [2 * L, 2 * L, 2 * L], [0.5 * L, 0.5 * L, 0.01 * L], [0.34 * L, 0.2 * L, 0.5 * L]]
    b[:6] = [[0.9 * L, 0.5 * L, 0.5 * L], [L, 0.0, 0.0], [0.5 * L, 0.5 * L, 0.5 * L], [3 * L, 2 * L, 2 * L], [0.5 * L, 0.5 * L, 0.30 * L], [0.34 * L, 0.8 * L, 0.5 * L]]
    for guard in (True, False):
        ctx.reset_counters()
        got = dmf.willCollide(ctx, gv, a, b, guard_coords=guard)
        cnt = ctx.counters()
        want, steps = zip(*[oracle.will_collide(ov, a[i], b[i], guard) for i in range(len(a))])
        assert np.array_equal(got, np.array(want)), np.nonzero(got != np.array(want))
        assert cnt["samples"] == sum(steps)
        assert 0 < got.sum() < len(a)


def test_optimize_standoff_batch(dmf, oracle, ctx):
    """Algorithms::optimizeCameraPosition (Algorithms.hpp:394-421): batched binary search == per-camera reference search"""
    sc, ov, gv = _pair(dmf, oracle, ctx, "S128d")
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    L = float(sc.bounds[1])
    cams = np.concatenate([dmf.scenes.poses_sphere_lookat(L, 200, radius=0.1)[::50], dmf.scenes.poses_position_camera(L, 40, standoff=0.0)[[5, 22]]])
    mid, poses = dmf.optimizeCameraPosition(gv, eng, cams)
    mids = []
    for i, c in enumerate(cams):
        m, p = oracle.optimize_standoff(ov, K, H, W, c)
        mids.append(m)
        assert int(mid[i]) == m, (i, int(mid[i]), m)
        assert np.array_equal(poses[i], p)
    assert len(set(mids)) > 1     # the search actually depends on the camera


def test_reverse_anisotropic_offset_volume(dmf, oracle, ctx):
    from tests.test_forward_gpu import _aniso_scene
    sc = _aniso_scene(dmf)
    ov = oracle.volume_from_scene(sc, flat=True)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    K = dmf.scenes.REFERENCE_K
    eng = dmf.RayTracingEngine(dmf.Camera(K), ctx)
    centre = np.array([0.08, 0.2, 0.52])
    poses = np.stack([dmf.scenes.look_at(centre + 0.45 * v, centre) for v in dmf.scenes.sphere_directions(6.0)[3::9][:5]])
    for fmt in (1, 0):
        ctx.set_reverse_format(fmt)
        ctx.reset_counters()
        g = eng.reverse_views(gv, poses, fast=True)
        cnt = ctx.counters()
        tot = dict(samples=0, inbounds=0, hits=0)
        for i, p in enumerate(poses):
            o = oracle.reverse(ov, K, H, W, p, fast=True)
            assert np.array_equal(g["ids"][i], o["ids"])
            for k in tot:
                tot[k] += o["counters"][k]
            assert np.array_equal(eng.reverseRayTrace(gv, p, False)[1], oracle.reverse(ov, K, H, W, p, fast=False)["ids"])
        assert (cnt["samples"], cnt["inbounds"], cnt["hits"]) == (tot["samples"], tot["inbounds"], tot["hits"])
    a = np.random.default_rng(2).uniform(-0.4, 1.0, size=(64, 3)).astype(np.float32)
    b = np.random.default_rng(3).uniform(-0.4, 1.0, size=(64, 3)).astype(np.float32)
    got = dmf.willCollide(ctx, gv, a, b, True)
    assert np.array_equal(got, np.array([oracle.will_collide(ov, a[i], b[i], True)[0] for i in range(64)]))


def test_reposition_cameras_sampled_batched(dmf, oracle, ctx):
    """repositionCamerasSampled (tests/CameraPathGen.cpp:94-126) as one batched minimum cast: equals the per-camera
    restatement (whose arithmetic tests/test_reference_build_cpu.py pins to the reference's compiled driver)."""
    sc = dmf.scenes.scene("S64")
    ov = oracle.volume_from_scene(sc, flat=True)
    gv = dmf.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    K = dmf.scenes.REFERENCE_K
    L = float(sc.bounds[1])
    poses = np.concatenate([dmf.scenes.poses_sphere_lookat(L, 90)[::9], dmf.scenes.look_at([0.5 * L, 0.5 * L, 0.9 * L], [0.5 * L, 0.5 * L, 2.0 * L])[None, :]])
    near = [oracle.forward(ov, K, 480, 640, p, oracle.MODE_MINIMUM, 1, True, want_pixels=False)["min_depth"] for p in poses]
    got = dmf.repositionCamerasSampled(poses, gv, dmf.RayTracingEngine(dmf.Camera(K), ctx))
    assert np.array_equal(got, dmf.reposition_from_minimum(poses, near))
    assert near[-1] == -1 and np.array_equal(got[-1], np.asarray(poses[-1], np.float32))
