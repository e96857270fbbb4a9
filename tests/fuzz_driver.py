"""Differential fuzz of the skipping kernels against the brute-force march, run by tests/test_fuzz_gpu.py in a subprocess
(optionally against the bounds-checked build, DMF_B200_LIB=.../libdmf_b200_checked.so).

For every volume: N random poses (inside, outside, far away, grazing; rotations, the reference's non-rotation positionCamera
poses, axis-aligned rays, huge translations that disable skipping) -> for k_forward_line (GRID_BYTE), k_forward_skip (GRID_BIT)
and the brute-force k_forward: first-hit depth, hit voxel id, simulated points, visibility rows, found flags and the probe
counters (samples / inbounds / hits) must be identical; carve-on-line must leave the same observed grid as the brute-force
carve; the reverse march on distance bytes must equal the one on the bit grid.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
sys.path.insert(0, ROOT)
import dmf_b200 as D  # noqa: E402


def random_poses(rng, n, lo, hi):
    """n poses around the box [lo, hi]"""
    ext = hi - lo
    out = np.zeros((n, 3, 4), np.float32)
    for i in range(n):
        kind = i % 8
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        t = lo + ext * rng.uniform(-0.35, 1.35, 3)
        if kind == 1:                                   # looks at the middle of the volume from outside
            eye = lo + ext * (0.5 + rng.choice([-1, 1], 3) * rng.uniform(0.55, 1.2, 3))
            out[i] = D.scenes.look_at(eye, lo + ext * rng.uniform(0.3, 0.7, 3)).reshape(3, 4); continue
        if kind == 2:                                   # inside, looking anywhere
            t = lo + ext * rng.uniform(0.02, 0.98, 3)
        if kind == 3:                                   # axis-aligned rotation: rays parallel to voxel faces, samples on faces
            p = rng.permutation(3); q = np.zeros((3, 3)); q[p, np.arange(3)] = rng.choice([-1.0, 1.0], 3)
            t = lo + ext * np.round(rng.uniform(0, 1, 3) * 16) / 16
        if kind == 4:                                   # the reference's positionCamera shape: x = (0,-1,0), y = (1,0,0), z = -normal (not a rotation)
            nrm = rng.standard_normal(3); nrm /= np.linalg.norm(nrm)
            q = np.stack([[0, -1, 0], [1, 0, 0], -nrm], axis=1)
        if kind == 5:                                   # far away: large |t| makes the error bound exceed 0.1 voxel -> every sample evaluated exactly
            t = lo + ext * rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(1, 3)
        if kind == 6:                                   # just outside a face, looking along it (grazing)
            a = rng.integers(3); t = lo + ext * rng.uniform(0.1, 0.9, 3); t[a] = (lo if rng.random() < 0.5 else hi)[a] + ext[a] * rng.uniform(-0.01, 0.01)
        if kind == 7:                                   # scaled / sheared linear part
            q = q * rng.uniform(0.5, 1.5, 3)[None, :]
        out[i, :, :3] = q; out[i, :, 3] = t
    return out.reshape(n, 12)


def volumes():
    from tests.test_forward_gpu import _aniso_scene
    yield "S64", D.scenes.scene("S64"), None
    yield "S128-odd", D.scenes.scene("S128-odd"), None
    yield "S128-clutter", D.scenes.scene("S128-clutter"), None
    yield "aniso", _aniso_scene(D), None


def main():
    n_poses = int(os.environ.get("DMF_FUZZ_POSES", "2000"))
    H, W = 120, 160
    K = D.scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.25
    ctx = D.Context(0)
    lib_version = ctx.lib.dmf_version()
    rng = np.random.default_rng(20261019)
    report = {"lib_version": int(lib_version), "poses_per_volume": n_poses, "volumes": {}, "mismatches": 0, "bounds_violations": 0}
    for name, sc, _ in volumes():
        vol = D.VoxelVolume(ctx)
        vol.setDimensions(*sc.bounds); vol.setVolumeSize(*sc.dims); vol.constructVolume(); vol.integratePointCloud(sc.points, sc.normals)
        b = np.asarray(sc.bounds, np.float64)
        lo, hi = b[0::2], b[1::2]
        engines = {"line": D.RayTracingEngine(D.Camera(K, H, W), ctx, D.GRID_BYTE), "skip": D.RayTracingEngine(D.Camera(K, H, W), ctx, D.GRID_BIT),
                   "brute": D.RayTracingEngine(D.Camera(K, H, W), ctx, D.GRID_BIT, skip_empty=False)}
        bad = 0
        hits = 0
        for b0 in range(0, n_poses, 250):
            poses = random_poses(rng, min(250, n_poses - b0), lo, hi)
            zd, sparse = (sc.zdelta, False) if (b0 // 250) % 2 == 0 else (max(1, sc.zdelta // 2) + 1, True)
            res, cnts = {}, {}
            for mode in (D.MODE_POINTS, D.MODE_GOOD_POINTS):
                for k, eng in engines.items():
                    ctx.reset_counters()
                    res[k] = eng.forward_views(vol, poses, mode, zd, sparse, want=("depth", "points", "voxel", "visibility"))
                    c = ctx.counters(); cnts[k] = (c["samples"], c["inbounds"], c["hits"]); report["bounds_violations"] += c["bounds"]
                for k in ("line", "skip"):
                    for key in ("depth", "voxel", "visibility", "found_any"):
                        bad += int(not np.array_equal(res[k][key], res["brute"][key]))
                    bad += int(not np.array_equal(res[k]["points"].view(np.uint32), res["brute"]["points"].view(np.uint32)))
                    bad += int(cnts[k] != cnts["brute"])
                hits += int((res["brute"]["depth"] >= 0).sum())
            # carve: line-first vs brute force, same observed grid and counters
            obs = {}
            for k in ("line", "brute"):
                ctx.clear_observed(); ctx.reset_counters()
                engines[k if k == "line" else "brute"].forward_views(vol, poses, D.MODE_POINTS, zd, sparse, want=(), carve=True)
                c = ctx.counters(); cnts[k] = (c["samples"], c["inbounds"], c["hits"]); report["bounds_violations"] += c["bounds"]
                obs[k] = ctx.observed_words().copy()
            bad += int(not np.array_equal(obs["line"], obs["brute"])) + int(cnts["line"] != cnts["brute"])
            # reverse march: distance bytes (line-first) vs bit grid (every step)
            rp = poses[:40]
            rv = {}
            for fmt in (D.GRID_BYTE, D.GRID_BIT):
                ctx.set_reverse_format(fmt); ctx.reset_counters()
                rv[fmt] = engines["line"].reverse_views(vol, rp, fast=True, want=("visibility", "unoccluded"))
                c = ctx.counters(); rv[fmt]["cnt"] = (c["samples"], c["inbounds"], c["hits"]); report["bounds_violations"] += c["bounds"]
            ctx.set_reverse_format(D.GRID_BYTE)
            for key in ("visibility", "unoccluded", "found_any"):
                bad += int(not np.array_equal(rv[D.GRID_BYTE][key], rv[D.GRID_BIT][key]))
            bad += int(rv[D.GRID_BYTE]["cnt"] != rv[D.GRID_BIT]["cnt"])
        report["volumes"][name] = {"mismatching_comparisons": bad, "hit_pixels_seen": hits, "n_occupied": int(len(vol.occupied_cells_))}
        report["mismatches"] += bad
    print(json.dumps(report), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
