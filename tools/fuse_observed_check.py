"""Run under torchrun on >= 2 GPUs: every rank carves its interleaved share of a small sweep into its own observed-voxel
grid, fuse_observed() OR-all-reduces the grids over NCCL (all-gather + k_or_reduce), and every rank checks the result
against the grid it gets by carving ALL views itself.  Prints one line per rank; exit code 0 only if all match."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np
import torch
import torch.distributed as dist

import dmf_b200 as D
from dmf_b200.sweep import fuse_observed, shard_indices

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = D.Context(local)
sc = D.scenes.scene("S128")
vol = D.VoxelVolume(ctx)
vol.setDimensions(*sc.bounds); vol.setVolumeSize(*sc.dims); vol.constructVolume(); vol.integratePointCloud(sc.points, sc.normals)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K, 480, 640), ctx, D.GRID_BYTE)
poses = D.scenes.poses_sphere_lookat(float(sc.bounds[1]), 64)[::4]              # 16 views
mine = poses[shard_indices(len(poses), rank, world, "strided")]
eng.forward_views(vol, mine, D.MODE_POINTS, sc.zdelta, False, want=(), carve=True)
own = ctx.observed_counts()
fused = fuse_observed(ctx)
got = ctx.observed_words()
ctx.clear_observed()
eng.forward_views(vol, poses, D.MODE_POINTS, sc.zdelta, False, want=(), carve=True)
want = ctx.observed_words()
ok = bool(np.array_equal(got, want)) and fused["observed"] >= own["observed"] and (world == 1 or fused["observed"] > own["observed"])
print(f"rank {rank}/{world}: own {own['observed']} voxels, fused {fused['observed']} (free {fused['free']}, hit {fused['hit']}), "
      f"all views locally {ctx.observed_counts()['observed']}: {'OK' if ok else 'MISMATCH'}", flush=True)
t = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
dist.all_reduce(t)
ctx.close()
dist.destroy_process_group()
sys.exit(int(t.item() != 0))
