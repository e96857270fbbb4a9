#!/usr/bin/env python
"""The reference's single-view driver (tests/Raytracing.cpp:55-104), headless, for BASELINE.json configs[0] / configs[1]:
one 640x480 depth camera cast into the synthetic-box grid.  Same steps as the driver -- set up the volume, integrate the
cloud, place a camera with positionCamera-style poses, derive `resolution_single_dimension` from `voxel_size_`
(:84-85), call `reverseRayTraceFast(volume, pose, true, resolution)` (:91) or the commented `rayTraceAndClassify` (:92) --
with the viewer replaced by a count of the Voxel::view / Voxel::good marks it would colour.

    python examples/raytracing.py [--scene S128] [--forward] [--carve]

--forward also prints the simulated depth image statistics (first-hit z_depth per pixel) and, with --carve, the occupied /
free voxel counts of the observed grid (the north-star extension the reference does not have).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import dmf_b200 as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="S128", help="S128 = configs[0] (128^3, 8 mm voxels), S512 = configs[1]")
    ap.add_argument("--forward", action="store_true", help="rayTraceAndClassify (Raytracing.cpp:92) instead of reverseRayTraceFast (:91)")
    ap.add_argument("--carve", action="store_true", help="with --forward: also mark every visited voxel in the observed grid")
    a = ap.parse_args()

    sc = D.scenes.scene(a.scene)
    ctx = D.Context(0)
    volume = D.VoxelVolume(ctx)
    volume.setDimensions(*sc.bounds)                      # Raytracing.cpp:69
    volume.setVolumeSize(*sc.dims)                        # :75
    volume.constructVolume()                              # :76
    volume.integratePointCloud(sc.points, sc.normals)     # :77
    volume._commit(ctx)                                   # build + upload now (otherwise done lazily by the first engine call)
    print(f"Volume Integrated: {len(volume.occupied_cells_)} occupied voxels, voxel_size_ = {volume.voxel_size_:.3e} m^3")

    L = float(sc.bounds[1])
    camera_locations = D.scenes.poses_position_camera(L, 8)           # positionCameras(locations), :81
    cam = D.Camera(D.scenes.REFERENCE_K)                              # :61, :82
    resolution_single_dimension = int(round(np.cbrt(volume.voxel_size_ * 1e9)))     # :84-85
    print("Resolution Single Dim:", resolution_single_dimension)

    engine = D.RayTracingEngine(cam, ctx)                             # :90
    pose = camera_locations[0]
    t0 = time.perf_counter()
    if a.forward:
        res = engine.forward_views(volume, pose, D.MODE_CLASSIFY, resolution_single_dimension, False,
                                   want=("depth",), carve=a.carve)
        dt = time.perf_counter() - t0
        depth = res["depth"][0]
        hit = depth >= 0
        print(f"rayTraceAndClassify: {1e3 * dt:.2f} ms, {int(hit.sum())} of {depth.size} pixels hit"
              + (f", depth {int(depth[hit].min())}..{int(depth[hit].max())} mm" if hit.any() else ""))
        if a.carve:
            c = ctx.observed_counts()
            print(f"observed grid: {c['free']} voxels seen free, {c['hit']} seen occupied")
    else:
        found, ids = engine.reverseRayTraceFast(volume, pose, True, resolution_single_dimension)
        dt = time.perf_counter() - t0
        print(f"reverseRayTraceFast: {1e3 * dt:.2f} ms, found = {found}, {len(ids)} voxels visible with a good normal")
    view, good = volume.marks()                                       # what addVolumeWithVoxelsClassified colours (:93)
    print(f"Voxel::view set on {int((view != 0).sum())} voxels, Voxel::good on {int(good.sum())}")
    ctx.close()


if __name__ == "__main__":
    main()
