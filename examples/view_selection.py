#!/usr/bin/env python
"""The reference's view-selection pipeline (tests/SetCover.cpp:255-318) end to end on the GPU, for BASELINE.json configs[2]:
1 024 sphere-sampled candidate views x 640x480 into a 512^3 grid -> per-view visibility bitsets -> greedy set cover ->
the chosen poses written in the reference's camera-file format (FileRoutines.hpp:98-112).

    python examples/view_selection.py [--views 1024] [--scene S512] [--forward] [--out cameras.txt]

Default uses reverseRayTraceFast like the shipped driver; --forward uses rayTraceAndGetGoodPoints (the commented
alternative at tests/SetCover.cpp:228).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import dmf_b200 as D  # noqa: E402
from dmf_b200.posefile import write_camera_locations  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=1024)
    ap.add_argument("--scene", default="S512")
    ap.add_argument("--forward", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sc = D.scenes.scene(a.scene)
    ctx = D.Context(0)
    vol = D.VoxelVolume(ctx)
    vol.setDimensions(*sc.bounds); vol.setVolumeSize(*sc.dims); vol.constructVolume(); vol.integratePointCloud(sc.points, sc.normals)
    eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE)
    poses = D.scenes.bench_poses(float(sc.bounds[1]), a.views)
    eng._prepare(vol)
    t0 = time.perf_counter()
    if a.forward:
        vis = eng.forward_views(vol, poses, D.MODE_GOOD_POINTS, sc.zdelta, False, want=("visibility",))["visibility"]
    else:
        vis = eng.reverse_views(vol, poses, fast=True, want=("visibility",))["visibility"]
    t1 = time.perf_counter()
    selected = D.greedySetCover(vis, ctx)
    t2 = time.perf_counter()
    covered = np.bitwise_or.reduce(vis[selected], axis=0) if len(selected) else np.zeros(vis.shape[1], np.uint64)
    n_cov = int(np.unpackbits(covered.view(np.uint8)).sum())
    n_any = int(np.unpackbits(np.bitwise_or.reduce(vis, axis=0).view(np.uint8)).sum())
    print(f"{a.views} candidate views, {len(vol.occupied_cells_)} occupied voxels: visibility {1e3*(t1-t0):.1f} ms, set cover {1e3*(t2-t1):.1f} ms "
          f"-> {len(selected)} views cover {n_cov} of the {n_any} voxels any view sees")
    if a.out:
        write_camera_locations(a.out, poses[selected])
        print("wrote", a.out)
    ctx.close()


if __name__ == "__main__":
    main()
