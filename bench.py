#!/usr/bin/env python
"""bench.py -- headline benchmark of the supervoxel-plus-merging hot path (BASELINE.json).

Metric: Mpoints/s end-to-end segmentation (voxelize -> VCCS -> edge weights -> merge) on the
C2 workload: synthetic 640x480 RGB-D frames (307,200 points each, seed 20020 + k), flags
--CVX --AL -t 0.2.  A "step" is one frame through f3ps_run.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU oracle port, rank 0 only)

`value`  : frames resident in HBM before the timed region, CUDA events on the stream the kernels run on.
`e2e`    : the same frames through the public API with HOST buffers (H2D of the points and D2H of the
           labelled voxel cloud + merge log inside the timed region).
Frames are sharded one stream per GPU with no collective: every rank runs K
steps on its own frames ("weak" scaling); the time is the max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200"))

WORKLOAD = "C2: synthetic 640x480 RGB-D frame (307200 points), --CVX --AL -t 0.2"
FLAGS = dict(color_mode=0, geom_mode=1, merge_mode=1, lam=0.5, bins=500)
THRESHOLD = 0.2
N_POOL = 4                     # distinct frames rotated through the steps
L2_FLUSH_BYTES = 512 << 20     # > 126 MB L2


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = "/tmp/f3ps_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); smax.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def make_frames(rank, count):
    from f3ps import synth
    return [synth.make_frame(seed=20020 + rank * 1000 + i) for i in range(count)]


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle port on the host cores, one frame per core per step (frames are independent
    in the reference: fresh SupervoxelClustering + Clustering per file, src/supervoxel_clustering.cpp:348,408)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    oracle_py.build()
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    frames = make_frames(0, 1)
    npts = len(frames[0])
    with ProcessPoolExecutor(max_workers=cores) as ex:
        def step():
            t0 = time.perf_counter()
            list(ex.map(_oracle_frame, [(frames[0], 1)] * cores))
            return time.perf_counter() - t0
        for _ in range(args.warmup):
            step()
        times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = cores * npts * args.steps / total / 1e6
    line = {"impl": "reference", "metric": "Mpoints/s end-to-end segmentation", "value": value, "unit": "Mpoints/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": cores},
            "cpu_baseline": {"value": value, "unit": "Mpoints/s", "cores": cores, "kind": "port",
                             "sample": "%d VGA frames per step, one per core, CPU oracle with the stamp-based merge "
                                       "(identical results to the literal std::multimap replay, which is ~30x slower)" % cores},
            "e2e": {"value": value, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def _oracle_frame(arg):
    pts, merge_impl = arg
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    o = oracle_py.Oracle()
    o.set_vccs_params()
    o.set_merge_params(merge_impl=merge_impl, **FLAGS)
    o.set_input(pts)
    t0 = time.perf_counter()
    o.run(0, THRESHOLD)
    return time.perf_counter() - t0, o.array("stage_ms")


def cpu_baseline(frame):
    """Rank 0, N=1: the literal CPU oracle (std::multimap replay, as the reference is written) on one frame."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    oracle_py.build()
    t_lit, st_lit = _oracle_frame((frame, 0))
    t_fix, st_fix = _oracle_frame((frame, 1))
    n = len(frame)
    return {"value": n / t_lit / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
            "sample": "1 VGA frame (307200 points), single thread, literal std::multimap merge replay: %.2f s" % t_lit,
            "stage_ms": {k: round(float(v), 3) for k, v in zip(["voxelize", "neighbors", "normals", "seeds", "expand", "graph", "merge", "total"], st_lit)},
            "fixed_merge_value": n / t_fix / 1e6,
            "fixed_merge_note": "same oracle with the stamp-based merge (identical results): %.3f s/frame" % t_fix,
            "host_cores_available": os.cpu_count()}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import f3ps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the f3ps path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    frames = make_frames(rank, N_POOL)
    npts = len(frames[0])
    stream = torch.cuda.current_stream()
    seg = f3ps.Segmenter(device=local_rank, stream=stream.cuda_stream)
    seg.set_vccs_params()
    seg.set_merge_params(**FLAGS)

    # resident inputs
    d_frames = [torch.from_numpy(f.view(np.uint8).reshape(-1, 32).copy()).to(dev) for f in frames]
    pinned = [torch.from_numpy(f.view(np.uint8).reshape(-1, 32).copy()).pin_memory() for f in frames]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def step_resident(i):
        t = d_frames[i % N_POOL]
        seg.set_input_device(t.data_ptr(), npts, 32)
        seg.run(THRESHOLD)

    out_bytes = [0]

    def step_e2e(i):
        seg.set_input(pinned[i % N_POOL].numpy().view(f3ps.synth.POINT_DTYPE).reshape(-1))
        seg.run(THRESHOLD)
        x = seg.array("out_xyz"); l = seg.array("out_label"); m = seg.array("merges_ab"); w = seg.array("merges_w")
        out_bytes[0] = x.nbytes + l.nbytes + x.shape[0] * 4 + m.nbytes + w.nbytes + m.nbytes

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = seg.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_acc = {}
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)                  # L2 flush between timed iterations (outside the events)
        ev[k][0].record(stream)
        step_resident(k)
        ev[k][1].record(stream)
        torch.cuda.synchronize()
        for name, ms in seg.stage_ms().items():
            stage_acc[name] = stage_acc.get(name, 0.0) + ms
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = seg.launch_count() - launches0
    clocks = sampler.stop()
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3          # seconds, this rank
    counts = seg.counts()

    # end-to-end through the public API with host buffers
    for i in range(2):
        step_e2e(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(k)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0

    tmax = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t_dev_max, t_e2e_max = float(tmax[0]), float(tmax[1])

    if rank == 0:
        peak, peak_src = load_peaks()
        value = npts * args.steps * world / t_dev_max / 1e6
        e2e = npts * args.steps * world / t_e2e_max / 1e6
        V, M = counts.n_voxels, counts.n_merges
        stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
        # dominant kernel = the persistent merge kernel (K7); algorithmic bytes per launch:
        # 12 E (edge list) + 40 S (region statistics) + 12 M (merge log) + 16 * fold_steps (voxels re-read by the folds)
        merge_ms = stage_ms.get("merge_kernel", stage_ms.get("merge", 0.0))
        alg_bytes = 12 * counts.n_edges + 40 * counts.n_supervoxels + 12 * M + 16 * counts.fold_steps
        ach = alg_bytes / (merge_ms * 1e-3) / 1e9 if merge_ms > 0 else 0.0
        e2e_bytes = 16 * npts + 16 * V + 12 * M
        line = {
            "metric": "Mpoints/s end-to-end segmentation", "value": value, "unit": "Mpoints/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev_max / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames": "one per step per GPU, %d distinct frames rotated" % N_POOL,
                       "l2": "L2 flushed (512 MB write) between timed iterations", "sharding": "frames per GPU, no collective",
                       "V": int(V), "S": int(counts.n_supervoxels), "E": int(counts.n_edges), "M": int(M)},
            "e2e": {"value": e2e, "unit": "Mpoints/s", "h2d_bytes_per_step": npts * 32, "d2h_bytes_per_step": int(out_bytes[0])},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "roofline": {"bound": "hbm", "kernel": "merge_kernel (K7, persistent single block, latency-bound by design)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         "e2e_algorithmic_bytes": e2e_bytes,
                         "e2e_achieved_gbs": e2e_bytes / (t_dev_max / args.steps) / 1e9 if t_dev_max > 0 else 0.0,
                         "e2e_frac": e2e_bytes / (t_dev_max / args.steps) / 1e9 / peak if t_dev_max > 0 else 0.0},
            "wall_s": wall,
        }
        if world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(frames[0])
            except Exception as e:     # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "Mpoints/s", "cores": 1, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
