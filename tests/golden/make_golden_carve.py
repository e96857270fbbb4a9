"""Generates tests/golden/golden_carve_v1.npz from the CPU oracle: the observed-voxel bit grids of carve mode.

Carve mode ("occupied/free voxel marking") is an extension over the reference, defined by oracle/dmf_oracle.hpp
(PixelOut::observed).  Same scenes, poses, camera and z strides as golden_v1.npz (read from it), so the two files
describe the same casts.  Regenerate only on purpose:

    python tests/golden/make_golden_carve.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from dmf_b200 import scenes  # noqa: E402

CASES = ["S64", "S128-odd", "S128-clutter"]


def main():
    g = np.load(os.path.join(HERE, "golden_v1.npz"))
    K = g["K"]; H, W = (int(v) for v in g["HW"])
    out = {}
    for name in CASES:
        sc = scenes.scene(name)
        vol = O.volume_from_scene(sc, flat=True)
        poses = g[f"{name}/poses"]; zd = int(g[f"{name}/zdelta"])
        for sparse in (0, 1):
            obs, inb = None, 0
            for i, p in enumerate(poses):
                obs, c = O.forward_observed(vol, K, H, W, p, O.MODE_POINTS, zd, bool(sparse), observed=obs)
                inb += c["inbounds"]
                if i == 0:
                    out[f"{name}/s{sparse}/view0"] = obs.copy()
            out[f"{name}/s{sparse}/all_views"] = obs
            out[f"{name}/s{sparse}/inbounds"] = np.int64(inb)
    path = os.path.join(HERE, "golden_carve_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", len(out), "arrays")


if __name__ == "__main__":
    main()
