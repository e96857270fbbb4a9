"""Generates tests/golden/golden_v1.npz from the CPU oracle (oracle/, canonical Eigen order 0).

The reference ships no fixtures and cannot be built here, so these vectors pin OUR oracle
at the commit that made them: any later change of the oracle's arithmetic, of the scene/pose generators, or of
the CUDA path shows up as a diff against this file.  Regenerate only on purpose:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from dmf_b200 import scenes  # noqa: E402

H, W = 96, 128
K = scenes.REFERENCE_K.copy()
K[[0, 2, 4, 5]] *= 0.2
CASES = [("S64", 6), ("S128-odd", 5), ("S128-clutter", 5)]


def poses_for(sc):
    L = float(sc.bounds[1])
    return np.stack([scenes.pose_p1(L)[0]] + list(scenes.poses_sphere_lookat(L, 120)[::40]) + [scenes.poses_position_camera(L, 40)[23]])


def main():
    out = {"K": K, "HW": np.array([H, W], np.int32)}
    for name, zdelta in CASES:
        sc = scenes.scene(name)
        vol = O.volume_from_scene(sc, flat=True)
        poses = poses_for(sc)
        out[f"{name}/poses"] = poses
        out[f"{name}/zdelta"] = np.int32(zdelta)
        out[f"{name}/occupied"] = vol.occupied()
        for i, p in enumerate(poses):
            for mode, tag in ((O.MODE_POINTS, "points"), (O.MODE_GOOD_POINTS, "good")):
                for sparse in (0, 1):
                    r = O.forward(vol, K, H, W, p, mode, zdelta, bool(sparse))
                    out[f"{name}/{i}/{tag}/s{sparse}/ids"] = r["ids"]
                    if mode == O.MODE_POINTS:
                        out[f"{name}/{i}/depth/s{sparse}"] = r["depth"].astype(np.int16)
                        out[f"{name}/{i}/counters/s{sparse}"] = np.array([r["counters"][k] for k in ("samples", "inbounds", "hits")], np.int64)
            out[f"{name}/{i}/min"] = np.int32(O.forward(vol, K, H, W, p, O.MODE_MINIMUM, 1, True, want_pixels=False)["min_depth"])
            rv = O.reverse(vol, K, H, W, p, fast=True)
            out[f"{name}/{i}/reverse_fast/ids"] = rv["ids"]
            out[f"{name}/{i}/reverse_fast/flags"] = rv["flags"]
            out[f"{name}/{i}/reverse_slow/ids"] = O.reverse(vol, K, H, W, p, fast=False)["ids"]
            zb, n = O.zbuffer(vol, K, H, W, p)
            out[f"{name}/{i}/zbuffer"] = zb.astype(np.int32)
            out[f"{name}/{i}/zbuffer_n"] = np.int64(n)
        vol.clear_marks()
        for i, p in enumerate(poses):
            O.forward(vol, K, H, W, p, O.MODE_CLASSIFY, zdelta, False, view=1 + i, want_pixels=False)
        view, good = vol.marks()
        out[f"{name}/classify/view"] = view
        out[f"{name}/classify/good"] = good
        sets = [np.sort(out[f"{name}/{i}/reverse_fast/ids"]) for i in range(len(poses))]
        out[f"{name}/setcover"] = O.greedy_set_cover(sets)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", len(out), "arrays")


if __name__ == "__main__":
    main()
