"""CPU model check of the index-range claim carve_on_line_sign relies on (csrc/dmf_forward.cuh).

The kernel drops the clamp of the located voxel index because every located sample has k in [kin, kout], where the float
line Q(k) = fma(k, QB, QA) is at least 0.25 voxel inside the volume on every axis, so rint(Q - 0.5) lies in [0, dim] -- the
padded index space (pdim = dim + 1).  This restates the kernel's per-ray interval arithmetic in numpy float32 (the reciprocal
perturbed by +-2^-20 relative, more than MUFU.RCP's error) for cameras outside, inside, grazing and axis-parallel, on the bench
volume, an anisotropic off-origin one and a non-dyadic one, and checks the claim for every k of every ray.  It also checks
the chunk walk: every sample of a safe range is located, none outside it, always in full groups.
"""
import numpy as np
import pytest

f32 = np.float32


def _fma(a, b, c):
    # float32 fma: the product of two float32 is exact in float64; one rounding of the sum to float64, one to float32.
    # (Double rounding can differ from a true fma by one float32 ulp in rare cases: irrelevant against a 0.25-voxel margin.)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _line_and_interval(pose, K, H, W, bounds, dims, z0, zdelta, S, rcp_eps, rng, n_rays):
    """qa, qb, kin, kout per ray exactly as k_forward_line's prologue computes them (:960-1000)."""
    fx, cx, fy, cy = (float(K[i]) for i in (0, 2, 4, 5))
    c = rng.integers(0, W, n_rays); r = rng.integers(0, H, n_rays)
    dcx = ((c.astype(np.float64) - cx) / fx).astype(f32); dcy = ((r.astype(np.float64) - cy) / fy).astype(f32)
    R = pose.reshape(3, 4).astype(f32)
    vmin = np.asarray(bounds[0::2], np.float64); vmax = np.asarray(bounds[1::2], np.float64)
    delta = (vmax - vmin) / np.asarray(dims, np.float64)
    dim = np.floor((vmax - vmin) / delta).astype(np.int64)            # constructVolume recomputes (Volume.hpp:121-123)
    inv = 1.0 / delta
    inv32 = inv.astype(f32); c32 = (-vmin * inv).astype(f32)
    ext = np.nextafter(((vmax - vmin) * inv * (1.0 + 1e-7)).astype(f32), f32(np.inf))
    z0m = f32(z0) * f32(0.001); zdm = f32(zdelta) * f32(0.001)
    qa = np.empty((3, n_rays), f32); qb = np.empty((3, n_rays), f32)
    ti0 = np.full(n_rays, -1e30, f32); ti1 = np.full(n_rays, 1e30, f32)
    for ax in range(3):
        one = np.ones(n_rays, f32)
        g = _fma(R[ax, 0] * one, dcx, _fma(R[ax, 1] * one, dcy, R[ax, 2] * one))
        qa[ax] = _fma(_fma(z0m * one, g, R[ax, 3] * one), inv32[ax] * one, c32[ax] * one)
        qb[ax] = (zdm * g) * inv32[ax]
        with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
            rc = np.where(np.abs(qb[ax]) > f32(1e-12), (f32(1.0) / qb[ax]) * f32(1.0 + rcp_eps), f32(1e30)).astype(f32)
            ta = (f32(0.25) - qa[ax]) * rc; tb = (ext[ax] - f32(0.25) - qa[ax]) * rc
        ti0 = np.maximum(ti0, np.minimum(ta, tb)); ti1 = np.minimum(ti1, np.maximum(ta, tb))
    Sf = f32(S)
    ok = ti0 <= ti1
    kin = np.where(ok, np.minimum(np.maximum(np.ceil(ti0) + 1, 0), Sf), 1).astype(np.int64)
    kout = np.where(ok, np.minimum(np.maximum(np.floor(ti1) - 1, -1), Sf - 1), 0).astype(np.int64)
    return qa, qb, kin, kout, dim


def _look_at(eye, target, up=(0.0, 0.0, 1.0)):
    import dmf_b200
    return dmf_b200.scenes.look_at(eye, target, up)


VOLUMES = {
    "S512": ((0, 1, 0, 1, 0, 1), (512, 512, 512), 2),
    "aniso-offset": ((-0.31, 0.47, -0.21, 0.61, 0.12, 0.92), (200, 312, 160), 5),
    "non-dyadic": ((0, 0.937, 0, 0.937, 0, 0.937), (512, 512, 512), 2),
    "S1024": ((0, 1, 0, 1, 0, 1), (1024, 1024, 1024), 1),
}


@pytest.mark.parametrize("name", list(VOLUMES))
@pytest.mark.parametrize("rcp_eps", [0.0, 2.0 ** -20, -(2.0 ** -20)])
def test_located_index_stays_in_the_padded_grid(dmf, name, rcp_eps):
    bounds, dims, zdelta = VOLUMES[name]
    K = dmf.scenes.REFERENCE_K
    H, W = 480, 640
    z0, S = 10, -(-(1000 - 10) // zdelta)
    lo = np.asarray(bounds[0::2]); hi = np.asarray(bounds[1::2]); ctr = 0.5 * (lo + hi); L = float((hi - lo).max())
    poses = list(dmf.scenes.poses_sphere_lookat(L, 200)[::23] + 0)          # around the (unit) cube
    poses += [_look_at(ctr + [-0.9 * L, 0, 0], ctr), _look_at(ctr, ctr + [0.3, 0.1, 0.2]),                   # outside / inside
              _look_at(lo + [-0.2 * L, -0.2 * L, 0.001], hi * [1, 1, 0] + [0, 0, lo[2] + 0.004]),            # grazing a face
              dmf.scenes.pose_p1(L)[0],                                                                       # axis-parallel rays
              _look_at(lo - 0.05 * L, hi), _look_at(hi + 0.3 * L, lo)]
    rng = np.random.default_rng(7)
    checked = 0
    for pose in poses:
        qa, qb, kin, kout, dim = _line_and_interval(np.asarray(pose, f32), K, H, W, bounds, dims, z0, zdelta, S, rcp_eps, rng, 4096)
        k = np.arange(S, dtype=np.int64)[:, None]
        inside = (k >= kin[None, :]) & (k <= kout[None, :])
        kf = np.broadcast_to(k.astype(f32), inside.shape)
        for ax in range(3):
            t = _fma(kf, np.broadcast_to(qb[ax], inside.shape), np.broadcast_to(qa[ax] - f32(0.5), inside.shape))
            n = np.rint(t)[inside]
            assert n.size == 0 or (n.min() >= 0 and n.max() <= dim[ax]), (name, ax, n.min(), n.max(), dim[ax])
        checked += int(inside.sum())
    assert checked > 100000


def test_chunk_walk_locates_every_sample_once_or_twice_in_full_groups():
    """carve_on_line_sign's walk over a safe range [b0, b1]: head group at b0 when the length is not a multiple of the group,
    then chunks of up to 32 samples in full groups; bit b of a chunk's `unsafe` word is sample kb + (b ^ (MLP - 1))."""
    for MLP in (4, 8, 16):
        for b0 in (0, 3, 17):
            for length in range(MLP, 200):
                b1 = b0 + length - 1
                head = length % MLP
                kb, n, nxt = b0, (MLP if head else min(32, length)), (b0 + head if head else b0 + min(32, length))
                seen = {}
                while True:
                    assert n % MLP == 0 and 0 < n <= 32
                    for j in range(0, n, MLP):
                        for u in range(MLP):
                            k = kb + j + u
                            assert b0 <= k <= b1
                            seen[k] = seen.get(k, 0) + 1
                            assert kb + ((j + (MLP - 1 - u)) ^ (MLP - 1)) == k
                    kb = nxt
                    if kb > b1:
                        break
                    n = min(32, b1 - kb + 1); nxt = kb + n
                assert sorted(seen) == list(range(b0, b1 + 1)) and max(seen.values()) <= 2
