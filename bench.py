#!/usr/bin/env python
"""bench.py -- rays/s and voxel-updates/s of the RayTracingEngine march (BASELINE.json metric), through the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--views V] [--config 2|3|4] [--impl reference]

--config 2 (default; the headline, BASELINE.json configs[1..2]): 640x480 camera, 512^3 box-shell grid (scene S512 of SURVEY 8d),
  zdelta = 2 mm, sparse = false, rayTraceAndGetPoints semantics.  A step = V views per GPU of the CameraPlacement sphere sweep
  P1024, dealt round-robin over the GPUs (view g on GPU g mod N; at every N the step's views are spread evenly over the whole
  sphere, so every N marches the same view mix).
    value : whole-job rays/s, poses and outputs resident in HBM (dmf_sweep_forward_dev: first-hit depth image + simulated point
            cloud + hit voxel ids per pixel, visibility row + found flag per view), CUDA events on the launching stream.  At
            N > 1 the exchange of the visibility rows is inside the step: the march kernels push finished rows into the peers'
            gathered buffers over NVLink (dmf_comm.cuh); the step ends when every peer's rows have arrived.
    e2e   : the same sweep through the host-buffer call dmf_sweep_forward: poses from pinned host memory, the result the
            reference's routine returns -- per view, which voxels it sees (visibility row) and whether it hit -- back in pinned
            host memory, every step.  Variants that also bring the per-pixel depth images back are reported beside it, with the
            box's measured device-to-host ceiling.
--config 3: moving-camera fusion: 1024^3 grid, helix trajectory (P10k), rayTraceAndClassify marks + occupied/free update (carve),
  marks and observed grids fused over the GPUs (min / OR reduce-scatter over peer memory).
--config 4: high-res stress: 1920x1080 camera, Fibonacci sweep (P4096) into the 1024^3 grid, visibility rows gathered and OR-reduced.
--impl reference times the reference's own hot-path headers compiled against oracle/ref_shim (oracle/_ref; the oracle port only
if that prebuilt library is missing) on the host cores, same workload, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))

CONFIGS = {
    2: dict(scene="S512", H=480, W=640, kscale=1.0, views=128, pose_set="P1024",
            workload="config 2 (BASELINE.json configs[1..2]): S512 box shell (512^3 grid, 1 m cube), 640x480 camera, zdelta=2mm dense, "
                     "CameraPlacement sphere look-at sweep P1024, views dealt round-robin over the GPUs"),
    3: dict(scene="S1024", H=480, W=640, kscale=1.0, views=250, pose_set="P10k",
            workload="config 3 (BASELINE.json configs[3]): moving-camera fusion, S1024 box shell (1024^3 grid), 640x480 camera, zdelta=1mm dense, "
                     "helix trajectory P10k subsampled evenly, rayTraceAndClassify + occupied/free update"),
    4: dict(scene="S1024", H=1080, W=1920, kscale=3.0, views=64, pose_set="P4096",
            workload="config 4 (BASELINE.json configs[4]): S1024 box shell (1024^3 grid), 1920x1080 camera, zdelta=1mm dense, Fibonacci "
                     "sphere sweep P4096 subsampled evenly, visibility rows gathered and OR-reduced"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4])
    ap.add_argument("--views", type=int, default=0, help="views per step per GPU (default: 128 / 250 / 64 for config 2 / 3 / 4)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", default="byte", choices=["bit", "byte"], help="byte = per-voxel Chebyshev distance bytes (default), bit = packed bits + macro-cell clearance")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-skip", action="store_true", help="evaluate every probe (brute-force kernel)")
    ap.add_argument("--two-probe", action="store_true", help="byte grid: the previous skipping kernel (k_forward_dist) instead of the line-first one")
    ap.add_argument("--count-in-timed", action="store_true", help="keep the DMF_CNT_* probe counters on in the timed launches")
    ap.add_argument("--cpu-threads", type=int, default=0, help="reference arm / all-core baseline: host threads (default: every CPU this process may run on)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hot = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(hot)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per k_forward_line launch from the committed ncu capture, if any (profiles/forward_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "forward_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def host_threads(args) -> int:
    """CPU threads of the reference arm: every CPU this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    if args.cpu_threads > 0:
        return args.cpu_threads
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def sweep_indices(n_total: int, n_set: int) -> np.ndarray:
    """n_total indices into a pose set of n_set poses, spread evenly over the whole set (every N marches the same view mix)."""
    if n_total <= n_set:
        return (np.arange(n_total, dtype=np.int64) * n_set) // n_total
    return np.arange(n_total, dtype=np.int64) % n_set


def pose_set(scenes, cfg, L: float) -> np.ndarray:
    if cfg["pose_set"] == "P1024":
        return scenes.poses_sphere_lookat(L, 1024)
    if cfg["pose_set"] == "P10k":
        return scenes.poses_helix(L, 10000)
    return scenes.poses_fibonacci(L, 4096)


def camera_K(scenes, cfg):
    return scenes.REFERENCE_K if cfg["kscale"] == 1.0 else scenes.scaled_K(cfg["kscale"])


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_backend():
    """The reference's own hot-path headers compiled here against a minimal Eigen/PCL shim (oracle/_ref, kind "reference")
    when that prebuilt library is present, else the oracle port (kind "port")."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_py
    if ref_py.available():
        return ref_py, "reference", ("reference include/{Camera,Volume,RayTracingEngine}.hpp compiled unmodified (g++ -O3 -fopenmp) against "
                                     "oracle/ref_shim, a minimal Eigen/PCL look-alike; stack scrubbed before each call (found[][] VLA is uninitialised in the reference)")
    import oracle_py
    return oracle_py, "port", "oracle/dmf_oracle.hpp restatement on the reference's vector<vector<vector<Voxel*>>> grid"


def cpu_volume(mod, sc):
    if mod.__name__ == "ref_py":
        return mod.volume_from_scene(sc)
    return mod.volume_from_scene(sc, flat=False)


def cpu_baseline(args, scenes, cfg, sc, poses, gpu_ids, n_views_serial=2) -> dict:
    """The reference's CPU path timed on this box's host cores, on the first views of the SAME step the GPU marched: one thread
    (as the reference runs it) with the returned id lists compared against the GPU's, then the same code with one view per
    host thread.  Also counts the tie cases of the bench views: what changes under the alternative Eigen summation order."""
    M, kind, how = cpu_backend()
    vol = cpu_volume(M, sc)
    K, H, W = camera_K(scenes, cfg), cfg["H"], cfg["W"]
    t0 = time.perf_counter()
    ref_ids = [M.forward(vol, K, H, W, poses[i], 0, sc.zdelta, False)["ids"] for i in range(n_views_serial)]
    sec = time.perf_counter() - t0
    ids_equal = all(np.array_equal(a, b) for a, b in zip(ref_ids, gpu_ids))
    out = {"value": n_views_serial * H * W / sec, "unit": "rays/s", "cores": 1, "kind": kind,
           "sample": f"the first {n_views_serial} views of the step the GPU marched, 1 thread (the reference hot path is single-threaded); {how}",
           "sec_per_view": sec / n_views_serial, "ids_equal_gpu": bool(ids_equal), "ids_per_view": [int(len(a)) for a in ref_ids]}
    if cfg is CONFIGS[2]:
        secr, _ = M.time_views(vol, K, H, W, poses[:1], 10 if kind == "reference" else 11, sc.zdelta, False, threads=1)
        out["reverse_sweep"] = {"views_per_s": 1.0 / secr, "sample": "1 view of reverseRayTraceFast incl. its dead getNeighborHashes work, 1 thread", "sec_per_view": secr}
        nt = host_threads(args)
        nv = max(nt, 2)
        sec2, _ = M.time_views(vol, K, H, W, np.tile(poses, (-(-nv // len(poses)), 1))[:nv], 0, sc.zdelta, False, threads=nt)
        out["all_cores"] = {"value": nv * H * W / sec2, "cores": nt, "sample": f"{nv} views over {nt} OpenMP threads (num_threads clause, independent of OMP_NUM_THREADS)"}
    del vol
    # tie cases on the bench views (north_star: "any face-boundary tie cases counted and reported"): the oracle under Eigen
    # order 0 (canonical, Eigen >= 3.3) and order 1 (translation + linear*v, Eigen 3.2): pixels whose first-hit depth differs, ids that differ
    try:
        import oracle_py as O
        ov = O.volume_from_scene(sc, flat=True)
        diff_px = diff_ids = 0
        same_as_gpu = True
        for i in range(n_views_serial):
            O.set_eigen_order(0)
            a = O.forward(ov, K, H, W, poses[i], O.MODE_POINTS, sc.zdelta, False)
            same_as_gpu = same_as_gpu and np.array_equal(a["ids"], gpu_ids[i])
            O.set_eigen_order(1)
            b = O.forward(ov, K, H, W, poses[i], O.MODE_POINTS, sc.zdelta, False)
            O.set_eigen_order(0)
            diff_px += int((a["depth"] != b["depth"]).sum())
            diff_ids += len(np.setxor1d(a["ids"], b["ids"]))
        out["tie_cases"] = {"views": n_views_serial, "pixels_whose_depth_changes_under_eigen_order_1": diff_px, "ids_that_change_under_eigen_order_1": diff_ids,
                            "of_pixels": n_views_serial * H * W, "oracle_order0_ids_equal_gpu": bool(same_as_gpu)}
    except Exception as e:                                                # noqa: BLE001 - reported, not fatal
        out["tie_cases"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dmf_b200 import scenes
    cfg = CONFIGS[args.config]
    M, kind, how = cpu_backend()
    sc = scenes.scene(cfg["scene"])
    vol = cpu_volume(M, sc)
    K, H, W = camera_K(scenes, cfg), cfg["H"], cfg["W"]
    nt = host_threads(args)
    nv = max(nt, 1)                      # one view per host thread per step: a bounded sample of the V-view batch
    V = args.views or cfg["views"]
    world = max(1, args.gpus)
    allp = pose_set(scenes, cfg, float(sc.bounds[1]))
    step_poses = allp[sweep_indices(V * world, len(allp))]          # the step the B200 arm marches
    # honour --steps / --warmup up to a run of a few minutes (a step is ~1-2 s of all-core work at VGA, more at 1080p); say so if clamped
    cap_steps, cap_warm = (30, 5) if args.config == 2 else (4, 1)
    steps, warm = max(1, min(args.steps, cap_steps)), max(0, min(args.warmup, cap_warm))
    take = lambda i: np.take(step_poses, np.arange(i * nv, (i + 1) * nv) % len(step_poses), axis=0)      # noqa: E731
    one_s, _ = M.time_views(vol, K, H, W, take(0)[:1], 0, sc.zdelta, False, threads=1)
    for i in range(warm):
        M.time_views(vol, K, H, W, take(i), 0, sc.zdelta, False, threads=nt)
    total = 0.0
    for i in range(steps):
        s, _ = M.time_views(vol, K, H, W, take(warm + i), 0, sc.zdelta, False, threads=nt)
        total += s
    val = steps * nv * H * W / total
    line = {
        "impl": "reference", "metric": "rays/s", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "views_per_step": nv,
                   "note": "CPU: " + how + "; each step is a bounded sample of the B200 arm's step: one view per host thread, taken from the same view list"},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": nt, "kind": kind, "sample": f"{nv} views per step over {nt} OpenMP threads (num_threads clause: independent of OMP_NUM_THREADS); {how}",
                         "one_thread": {"value": H * W / one_s, "cores": 1, "sample": "1 view, 1 thread: how the reference actually runs"}},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm: shared set-up
class Rig:
    """One rank: context, group (dmf_comm_init_rank; the id travels over torch.distributed), replicated volume, this step's poses."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import dmf_b200 as D
        self.torch, self.dist, self.D = torch, dist, D
        self.args = args
        self.cfg = cfg = CONFIGS[args.config]
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU path)")
        torch.cuda.set_device(self.local)
        from dmf_b200.sweep import bind_to_gpu_numa_node
        self.numa = bind_to_gpu_numa_node(self.local)        # before any pinned allocation: host buffers on the GPU's own socket
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = D.Context(self.local)
        uid = [D.Comm.unique_id() if (self.rank == 0 and self.world > 1) else None]
        if self.world > 1:
            dist.broadcast_object_list(uid, src=0, device=self.dev)
        self.comm = D.Comm.init_rank(self.ctx, uid[0] or b"", self.rank, self.world)
        self.info = self.comm.info()
        self.H, self.W = cfg["H"], cfg["W"]
        self.K = camera_K(D.scenes, cfg)
        self.sc = D.scenes.scene(cfg["scene"])
        self.fmt = D.GRID_BIT if args.grid == "bit" else D.GRID_BYTE
        self.comm.set_camera(self.K, self.H, self.W)
        # rank 0 builds the volume; it reaches the other GPUs GPU to GPU (occupied ids + normals over NCCL, structures rebuilt on each device)
        t0 = time.perf_counter()
        if self.rank == 0:
            self.vol = D.VoxelVolume(self.ctx)
            self.vol.setDimensions(*self.sc.bounds); self.vol.setVolumeSize(*self.sc.dims); self.vol.constructVolume(); self.vol.integratePointCloud(self.sc.points, self.sc.normals)
            self.vol._commit(self.ctx)
        self.barrier()
        t1 = time.perf_counter()
        self.comm.replicate_volume(0)
        if self.rank != 0:
            self.vol = D.VoxelVolume.attach(self.ctx)
        self.barrier()
        self.replicate_ms = 1e3 * (time.perf_counter() - t1)
        self.eng = D.RayTracingEngine(D.Camera(self.K, self.H, self.W), self.ctx, self.fmt)
        self.eng._prepare(self.vol)
        D._lib.check(self.ctx.lib.dmf_prepare_grid(self.ctx.h, self.fmt))
        a, b = C.c_float(), C.c_float()
        D._lib.check(self.ctx.lib.dmf_volume_prepare_ms(self.ctx.h, C.byref(a), C.byref(b)))
        self.prepare = {"structures_ms": a.value, "distance_bytes_ms": b.value, "replicate_to_peers_ms": self.replicate_ms if self.world > 1 else 0.0,
                        "what": "one-off per volume, on the device: bit grid + rank directory + macro-cell clearance + centroid hashes from the occupied id list; "
                                "Chebyshev distance bytes (O(n) separable transform); replication = id list + normals GPU to GPU, structures rebuilt per GPU"}
        self.n_occ = len(self.vol.occupied_cells_)
        self.vw = (self.n_occ + 63) // 64
        self.V = args.views or cfg["views"]
        allp = pose_set(D.scenes, cfg, float(self.sc.bounds[1]))
        self.idx = sweep_indices(self.V * self.world, len(allp))
        self.all_poses = np.ascontiguousarray(allp[self.idx])                        # the step's views, in view order
        self.my_poses = np.ascontiguousarray(self.all_poses[self.rank::self.world])  # view g on GPU g mod N
        self.n_total = len(self.all_poses)
        self.stream = torch.cuda.Stream(device=self.dev)       # a dedicated non-default stream: torch.cuda.Event sees torch's current stream
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def pinned(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.ctx.lib.dmf_host_alloc(max(n, 8))
        if not p:
            raise SystemExit("dmf_host_alloc failed")
        return np.frombuffer((C.c_char * max(n, 8)).from_address(p), dtype=dtype, count=int(np.prod(shape))).reshape(shape), p

    def d2h_ceiling(self):
        """Device-to-host rate of this box with every rank copying at once (pinned memory, 256 MiB x 4): the ceiling of any variant
        that brings per-pixel images back."""
        torch = self.torch
        n = 256 << 20
        src = torch.empty(n, dtype=torch.uint8, device=self.dev)
        dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        mine = 4 * n / sec / 1e9
        self.barrier()
        tot = self.sum_over_ranks([mine])[0]
        return {"per_rank_GBps": mine, "aggregate_GBps": tot, "how": "4 x 256 MiB cudaMemcpyAsync device -> pinned host per rank, all ranks at once"}

    def close(self):
        self.comm.close()
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def ptr_array(*ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


# ------------------------------------------------------------------------------------------------ config 2: the headline
def run_config2(args):
    rig = Rig(args)
    torch, dist, D = rig.torch, rig.dist, rig.D
    from dmf_b200._lib import ForwardOut, ForwardParams, ReverseOut, SweepOut, check
    ctx, comm, lib = rig.ctx, rig.comm, rig.ctx.lib
    H, W, V, vw, n_occ, world, rank, dev, sc, fmt = rig.H, rig.W, rig.V, rig.vw, rig.n_occ, rig.world, rig.rank, rig.dev, rig.sc, rig.fmt
    n_total = rig.n_total

    # ---- value: device-resident sweep step; per-pixel outputs stay in HBM, visibility rows are exchanged inside the step ----------
    d_poses = torch.from_numpy(rig.my_poses).to(dev)
    d_depth = torch.empty((V, H, W), dtype=torch.int32, device=dev)
    d_points = torch.empty((V, H, W, 3), dtype=torch.float32, device=dev)
    d_voxel = torch.empty((V, H, W), dtype=torch.int64, device=dev)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base_flags = (1 if args.no_skip else 0) | (2 if args.two_probe else 0)
    # timed launches run without the DMF_CNT_* probe counters (instrumentation the reference does not have); one identical untimed
    # step with the counters on supplies the per-step probe counts reported below
    params = ForwardParams(D.MODE_POINTS, sc.zdelta, 0, 1, fmt, base_flags | (0 if args.count_in_timed else D.FWD_NO_COUNTERS))
    params_counted = ForwardParams(D.MODE_POINTS, sc.zdelta, 0, 1, fmt, base_flags)
    o = ForwardOut()
    o.depth_mm, o.points, o.hit_voxel = d_depth.data_ptr(), d_points.data_ptr(), d_voxel.data_ptr()
    outs = (C.POINTER(ForwardOut) * 1)(C.pointer(o))
    poses_dev = ptr_array(d_poses.data_ptr())

    def step_dev(p=None):
        st = torch.cuda.current_stream().cuda_stream
        assert st != 0
        check(lib.dmf_sweep_forward_dev(comm.h, C.byref(p or params), poses_dev, n_total, outs, ptr_array(st)))

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    ctx.reset_counters()
    sampler = ClockSampler(rig.local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    hot_ms = []
    rig.barrier(); torch.cuda.synchronize()
    t_wall = time.perf_counter()
    for a, b in evs:
        if flush is not None:
            flush.fill_(1)                      # evict the L2 between timed iterations (not timed)
        a.record()
        step_dev()
        b.record()
        if rank == 0 and len(hot_ms) < 4:
            hot_ms.append(ctx.last_hot_kernel_ms())   # synchronises; cheap, outside the event pair's GPU time
    torch.cuda.synchronize(); rig.barrier()
    wall = time.perf_counter() - t_wall
    dev_ms = rig.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
    launches_timed = ctx.counters()["launches"]
    ctx.reset_counters()
    step_dev(params_counted); torch.cuda.synchronize()                  # the counted step (untimed): probe counts of ONE step
    cnt = ctx.counters()
    rays_total = args.steps * V * H * W * world
    value = rays_total / (dev_ms * 1e-3)
    inb1, samp1, launches_total, skip1, f641, exact1 = rig.sum_over_ranks(
        [cnt["inbounds"], cnt["samples"], launches_timed, cnt["skipped"], cnt["f64_path"], cnt["exact_div"]])
    inbounds_total, samples_total, skipped_total, f64_total, exact_total = (x * args.steps for x in (inb1, samp1, skip1, f641, exact1))

    # ---- e2e: the host-buffer sweep.  Poses from pinned host memory in, the routine's result (visibility row + found flag per view)
    # back in pinned host memory, every step; H2D and D2H inside the timed region -----------------------------------------------
    h_poses, p0 = rig.pinned((n_total, 12), np.float32)
    h_poses[:] = rig.all_poses
    h_vis, p1 = rig.pinned((n_total, max(vw, 1)), np.uint64)
    h_found, p2 = rig.pinned((n_total,), np.int32)
    so = SweepOut()
    so.visibility, so.found_any, so.rows_to_host = h_vis.ctypes.data, h_found.ctypes.data, D.comm.ROWS_OWN
    fp_all = h_poses.ctypes.data_as(C.POINTER(C.c_float))
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(max(2, min(args.warmup, 5))):
        check(lib.dmf_sweep_forward(comm.h, C.byref(params), fp_all, n_total, C.byref(so)))
    rig.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        check(lib.dmf_sweep_forward(comm.h, C.byref(params), fp_all, n_total, C.byref(so)))
    torch.cuda.synchronize(); rig.barrier()
    e2e_s = rig.max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_steps * V * H * W * world / e2e_s
    own_rows = h_vis[rank::world].copy()
    own_found = h_found[rank::world].copy()

    # ---- parity inside the run ---------------------------------------------------------------------------------------------
    parity = {}
    # (a) this rank's rows through the sweep == the same views through the single-GPU host call dmf_forward
    single = rig.eng.forward_views(rig.vol, rig.my_poses, D.MODE_POINTS, sc.zdelta, False, want=("visibility", "depth"))
    parity["sweep_rows_equal_single_gpu_call"] = bool(np.array_equal(own_rows[:, :vw], single["visibility"]) and np.array_equal(own_found, single["found_any"]))
    parity["device_depth_equal_host_call"] = bool(np.array_equal(d_depth.cpu().numpy(), single["depth"]))
    # (b) N > 1: the gathered array of EVERY rank (all rows brought back) == a recompute of all views on rank 0 alone
    if world > 1:
        so_all = SweepOut()
        h_all = np.zeros((n_total, max(vw, 1)), np.uint64); h_fall = np.zeros(n_total, np.int32)
        so_all.visibility, so_all.found_any, so_all.rows_to_host = h_all.ctypes.data, h_fall.ctypes.data, D.comm.ROWS_ALL
        check(lib.dmf_sweep_forward(comm.h, C.byref(params), fp_all, n_total, C.byref(so_all)))
        digest = hashlib.sha1(h_all.tobytes() + h_fall.tobytes()).hexdigest()
        digests = [None] * world
        dist.all_gather_object(digests, digest)
        parity["gathered_rows_identical_on_all_ranks"] = len(set(digests)) == 1
        if rank == 0:
            full = rig.eng.forward_views(rig.vol, rig.all_poses, D.MODE_POINTS, sc.zdelta, False, want=("visibility",))
            parity["gathered_rows_equal_rank0_recompute"] = bool(np.array_equal(h_all[:, :vw], full["visibility"]) and np.array_equal(h_fall, full["found_any"]))
    # the sweep's consumer: Algorithms::greedySetCover over the gathered rows, on the device
    t0 = time.perf_counter()
    cover = comm.set_cover()
    cover_ms = 1e3 * (time.perf_counter() - t0)

    # ---- variants that bring per-pixel images back (per-rank host call dmf_forward; PCIe-bound), with the box's D2H ceiling ------
    h_p, q0 = rig.pinned((V, 12), np.float32); h_p[:] = rig.my_poses
    h_depth, q1 = rig.pinned((V, H, W), np.int32)
    h_depth16, q2 = rig.pinned((V, H, W), np.uint16)
    h_v, q3 = rig.pinned((V, max(vw, 1)), np.uint64)
    h_f, q4 = rig.pinned((V,), np.int32)
    fp = h_p.ctypes.data_as(C.POINTER(C.c_float))

    def time_host(compact: bool):
        oh = ForwardOut()
        oh.visibility, oh.found_any = h_v.ctypes.data, h_f.ctypes.data
        if compact:
            oh.depth_u16 = h_depth16.ctypes.data
        else:
            oh.depth_mm = h_depth.ctypes.data
        n = max(3, min(args.steps, 10))
        for _ in range(2):
            check(lib.dmf_forward(ctx.h, C.byref(params), fp, V, C.byref(oh)))
        rig.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            check(lib.dmf_forward(ctx.h, C.byref(params), fp, V, C.byref(oh)))
        torch.cuda.synchronize(); rig.barrier()
        return rig.max_over_ranks(time.perf_counter() - t0), n

    i32_s, i32_n = time_host(False)
    u16_s, u16_n = time_host(True)
    depth_ok = bool(np.array_equal(h_depth, single["depth"]) and np.array_equal(np.where(h_depth < 0, 0xFFFF, h_depth), h_depth16.astype(np.int32)))
    parity["depth_u16_equal_int32"] = depth_ok
    ceiling = rig.d2h_ceiling()
    u16_bytes = V * (H * W * 2 + vw * 8 + 4)
    i32_bytes = V * (H * W * 4 + vw * 8 + 4)
    for p in (q0, q1, q2, q3, q4):
        lib.dmf_host_free(p)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "device-resident timed loop + e2e timed loops"

    # ---- secondary: the sweep the shipped drivers run (tests/SetCover.cpp:218-240): reverseRayTraceFast per view, through dmf_sweep_reverse
    rv_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        check(lib.dmf_sweep_reverse(comm.h, 1, fp_all, n_total, C.byref(so)))
    rig.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(rv_steps):
        check(lib.dmf_sweep_reverse(comm.h, 1, fp_all, n_total, C.byref(so)))
    torch.cuda.synchronize(); rig.barrier()
    rv_s = rig.max_over_ranks(time.perf_counter() - t0)
    rv_hot = ctx.last_hot_kernel_ms()
    t0 = time.perf_counter()
    rv_cover = comm.set_cover()
    rv_cover_ms = 1e3 * (time.perf_counter() - t0)
    rv_single = rig.eng.reverse_views(rig.vol, rig.my_poses, fast=True, want=("visibility",))
    parity["reverse_sweep_rows_equal_single_gpu_call"] = bool(np.array_equal(h_vis[rank::world][:, :vw], rv_single["visibility"]))

    # ---- secondary: carve mode (DMF_FWD_CARVE, "occupied/free voxel marking"): every in-bounds sample of every ray updates the
    # observed-voxel bit grid.  Device-resident, CUDA events on the bench stream; at N > 1 the grids are fused over the group.
    carve = None
    if not (args.no_skip or args.two_probe) and fmt == D.GRID_BYTE:
        cparams = ForwardParams(D.MODE_POINTS, sc.zdelta, 0, 1, fmt, D.FWD_CARVE)

        def carve_step():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            check(lib.dmf_sweep_forward_dev(comm.h, C.byref(cparams), poses_dev, n_total, outs, ptr_array(torch.cuda.current_stream().cuda_stream)))
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b), ctx.last_hot_kernel_ms()

        carve_step()                                  # warm-up (allocates and zeroes the observed grid)
        ctx.clear_observed(); ctx.reset_counters()
        rig.barrier()
        first_ms, first_hot = carve_step()
        inb_per_step = ctx.counters()["inbounds"]
        n_c = max(3, min(args.steps, 10))
        steady = [carve_step() for _ in range(n_c)]
        steady_ms, steady_hot = float(np.mean([x[0] for x in steady])), float(np.mean([x[1] for x in steady]))
        oc = ctx.observed_counts()
        fused = None
        if world > 1:
            rig.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
            comm.fuse_observed()
            torch.cuda.synchronize(); rig.barrier()
            fuse_ms = 1e3 * (time.perf_counter() - t0)
            oc_f = ctx.observed_counts()
            digest = hashlib.sha1(ctx.observed_words().tobytes()).hexdigest()
            digests = [None] * world
            dist.all_gather_object(digests, digest)
            fused = {"observed_voxels": oc_f["observed"], "free_voxels": oc_f["free"], "hit_voxels": oc_f["hit"], "ms": fuse_ms,
                     "how": "reduce-scatter + all-gather over peer memory (k_peer_reduce): each GPU ORs 1/N of the words from all peers and pushes the result back",
                     "grid_bytes": int(lib.dmf_observed_words(ctx.h)) * 4, "identical_on_all_ranks": len(set(digests)) == 1}
        first_ms, steady_ms = rig.max_over_ranks(first_ms), rig.max_over_ranks(steady_ms)
        inb_all = rig.sum_over_ranks([inb_per_step])[0]
        c_bytes = inb_per_step * 1.125 + V * H * W * 24 + V * (vw * 8 + 48)
        carve = {"what": "same step with DMF_FWD_CARVE: k_forward_line finds the hit, carve_on_line_sign marks the voxel of every visited in-bounds sample in the observed bit grid",
                 "voxel_updates_per_step": inb_all, "first_pass_ms": first_ms, "steady_ms_per_step": steady_ms,
                 "voxel_updates_per_s_first_pass": inb_all / (first_ms * 1e-3), "voxel_updates_per_s": inb_all / (steady_ms * 1e-3),
                 "rays_per_s": V * H * W * world / (steady_ms * 1e-3), "kernel_ms_first_pass": first_hot, "kernel_ms": steady_hot,
                 "observed_voxels": oc["observed"], "free_voxels": oc["free"], "hit_voxels": oc["hit"], "fused_over_ranks": fused,
                 "roofline": {"bound": "hbm", "algorithmic_bytes_per_launch": c_bytes, "achieved": c_bytes / (steady_hot * 1e-3) / 1e9,
                              "frac": c_bytes / (steady_hot * 1e-3) / 1e9 / measured_peak()[0], "unit": "GB/s",
                              "note": "1 B occupancy read + 1/8 B observed-bit write per in-bounds sample + 24 B per ray + bitset/pose per view"}}

    # ---- single-view calls, the way the reference's drivers use the engine (one pose per call, id list returned) ----
    single_calls = None
    gpu_ids = []
    if rank == 0:
        gpu_ids = [rig.eng.rayTraceAndGetPoints(rig.vol, rig.all_poses[i], sc.zdelta, False)[1] for i in range(2)]
        for i in range(3):                       # warm-up: first calls allocate the id-list scratch
            rig.eng.rayTraceAndGetPoints(rig.vol, rig.my_poses[i % V], sc.zdelta, False)
            rig.eng.reverseRayTraceFast(rig.vol, rig.my_poses[i % V], False)
        t0 = time.perf_counter()
        for i in range(20):
            rig.eng.rayTraceAndGetPoints(rig.vol, rig.my_poses[i % V], sc.zdelta, False)
        t1 = time.perf_counter()
        for i in range(20):
            rig.eng.reverseRayTraceFast(rig.vol, rig.my_poses[i % V], False)
        t2 = time.perf_counter()
        single_calls = {"rayTraceAndGetPoints_ms_per_call": 1e3 * (t1 - t0) / 20, "reverseRayTraceFast_ms_per_call": 1e3 * (t2 - t1) / 20,
                        "note": "one pose per call through the host-buffer C ABI incl. the discovery-ordered id list (what the drop-in RayTracingEngine does per call)"}

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes of one march launch: 1/8 B (bit grid) or 1 B (byte grid) per in-bounds sample,
        # + 24 B per cast ray (int32 depth, float3 point, uint64 hit id), + n_occ/8 B visibility and 48 B pose per view
        per_launch_inb = cnt["inbounds"]
        grid_b = 0.125 if fmt == D.GRID_BIT else 1.0
        alg_bytes = per_launch_inb * grid_b + V * H * W * 24 + V * (vw * 8 + 48)
        hot = float(np.mean(hot_ms)) if hot_ms else dev_ms / args.steps
        achieved = alg_bytes / (hot * 1e-3) / 1e9
        traffic = ncu_traffic()
        exch = rig.info["exchange_name"]
        line = {
            "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": rig.cfg["workload"], "views_per_step_per_gpu": V, "views_per_step": n_total, "view_indices_into_P1024": f"{int(rig.idx[0])}, {int(rig.idx[1])}, ... {int(rig.idx[-1])} (evenly spread; view g on GPU g mod N)",
                       "mode": "rayTraceAndGetPoints", "grid_format": args.grid,
                       "outputs": "depth_mm+points+hit_voxel per pixel in HBM, visibility row + found flag per view gathered on every GPU", "n_occupied": n_occ,
                       "l2": "flushed between timed iterations (256 MiB fill, untimed)" if flush is not None else "not flushed",
                       "probe_counters": "on in the timed launches" if args.count_in_timed else "off in the timed launches (DMF_FWD_NO_COUNTERS: instrumentation the reference does not have); counted in one identical untimed step",
                       "parallelism": f"C-ABI group of {world} GPU(s) (dmf_comm_init_rank), volume replicated GPU to GPU, exchange of the visibility rows: {exch}",
                       "host": rig.numa},
            "voxel_updates_per_s": inbounds_total / (dev_ms * 1e-3),
            "samples_per_s": samples_total / (dev_ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if not traffic else traffic.get("dram_bytes_per_view", 0) * V,
                         "traffic_source": None if not traffic else traffic.get("source"),
                         "peak_source": peak_src, "kernel": "k_forward" if args.no_skip else (("k_forward_dist" if args.two_probe else "k_forward_line") if fmt == D.GRID_BYTE else "k_forward_skip"),
                         "achieved_dram": None if not traffic else traffic.get("dram_bytes_per_view", 0) * V / (hot * 1e-3) / 1e9, "kernel_ms": hot, "kernel_ms_covers": "k_view_start + k_tile_start + the march kernel of one step (CUDA events inside the library around the three launches; the march kernel alone is ~90 % of it)",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "algorithmic bytes are SURVEY 8(d)'s reference-equivalent ones (1 B per in-bounds sample of the reference + the per-ray and per-view outputs); the kernel proves ~98 % of those samples empty from one distance byte each without touching memory, so frac can exceed 1 and is not an HBM utilisation: achieved_dram (measured DRAM bytes / kernel time) is. The kernel is issue-bound, see DESIGN.md section 5"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": V * 48, "d2h_bytes_per_step": V * (vw * 8 + 4), "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "bytes_are": "per rank",
                    "result": "what rayTraceAndGetPoints returns, per view: the set of voxels seen (visibility row over occupied_cells_) + found flag, in pinned host memory; "
                              "host call dmf_sweep_forward (poses H2D, march, fused row exchange, rows D2H inside the timed region)",
                    "with_depth_u16": {"value": u16_n * V * H * W * world / u16_s, "d2h_bytes_per_step": u16_bytes, "ms_per_step": 1e3 * u16_s / u16_n,
                                       "d2h_GBps_aggregate": u16_bytes * world * u16_n / u16_s / 1e9, "frac_of_d2h_ceiling": u16_bytes * world * u16_n / u16_s / 1e9 / ceiling["aggregate_GBps"],
                                       "note": "per-rank dmf_forward: the first-hit depth image as uint16 mm (0xFFFF = no hit) comes back too; round 1's headline e2e"},
                    "with_depth_int32": {"value": i32_n * V * H * W * world / i32_s, "d2h_bytes_per_step": i32_bytes, "ms_per_step": 1e3 * i32_s / i32_n,
                                         "d2h_GBps_aggregate": i32_bytes * world * i32_n / i32_s / 1e9, "frac_of_d2h_ceiling": i32_bytes * world * i32_n / i32_s / 1e9 / ceiling["aggregate_GBps"]},
                    "host_d2h_ceiling": ceiling},
            "parity_in_run": parity,
            "set_cover_over_gathered_rows": {"selected_views": [int(x) for x in cover], "ms": cover_ms, "n_sets": n_total},
            "volume_prepare_ms": rig.prepare,
            "gpu_launches": int(launches_total),
            "probes": {"reference_equivalent_per_step": samples_total / args.steps, "in_bounds_per_step": inbounds_total / args.steps,
                       "skipped_as_provably_empty_per_step": skipped_total / args.steps, "redone_in_f64_per_step": f64_total / args.steps,
                       "exact_division_ties_per_step": exact_total / args.steps},
            "reverse_sweep": {"what": "reverseRayTraceFast over the same step via dmf_sweep_reverse (host poses in, visibility rows gathered on every GPU, own rows to the host), then greedy set cover on the device",
                              "views_per_s": rv_steps * V * world / rv_s, "voxel_rays_per_s": rv_steps * V * world * n_occ / rv_s,
                              "ms_per_step": 1e3 * rv_s / rv_steps, "kernel_ms_per_step": rv_hot, "set_cover_ms": rv_cover_ms, "selected_views": [int(x) for x in rv_cover]},
            "carve": carve,
            "single_view_calls": single_calls,
            "clocks": clocks,
            "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, D.scenes, rig.cfg, sc, rig.all_poses, gpu_ids)
            line["parity_in_run"]["reference_ids_equal_gpu"] = line["cpu_baseline"]["ids_equal_gpu"]
            line["tie_cases"] = line["cpu_baseline"].pop("tie_cases", None)
        print(json.dumps(line), flush=True)
    for p in (p0, p1, p2):
        lib.dmf_host_free(p)
    rig.close()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 2:
        run_config2(args)
    else:
        from bench_fusion import run_config34           # configs 3 and 4 (same contract, same Rig)
        run_config34(args, sys.modules[__name__])


if __name__ == "__main__":
    main()
