#!/usr/bin/env python
"""bench.py -- the registration hot path on B200, measured.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One STEP = one complete ICP registration of the workload BASELINE.json quotes its metric on
(configs[1]): an ETH-Apartment-shaped synthetic scan pair (344 sweeps x 1077 beams, ~370k points
per scan), k-NN matching, point-to-plane linear ICP, 30 iterations, select-all, constant weights,
normal-angle rejection on -- including the index build (the reference does buildIndex inside
estimatePose, ICPOptimizer.h:532-535).

  value    registrations/s with the raw clouds already resident in HBM (device pointers in, pose out)
  e2e      the same registration through the host-pointer C ABI call a reference user would make
           (icp_gpu_set_target / set_source / estimate_pose with host arrays): the H2D copy of both
           clouds and the D2H read of the pose are inside the timed region
  roofline the dominant kernel (the fused k-NN match kernel), CUDA-event timed per launch
  cpu_baseline  the reference's own LinearICPOptimizer::estimatePose (oracle/_ref: the reference headers compiled in
           place against the stand-ins of oracle/ref_shim) on the same pair, a bounded number of iterations; the
           oracle port (oracle/icp_oracle.c) when that library is absent

N > 1 (torchrun): independent pairs sharded across ranks, one queue per GPU, no collective on the
data path -> weak scaling; time = max over ranks.

--impl reference: times the reference's own code path -- LinearICPOptimizer::estimatePose from
/root/reference/icp-variants/ICPOptimizer.h, compiled in place into oracle/_ref/libicp_ref.so -- on the box's host
cores, on a bounded sample (a few iterations) of the same workload.  Eigen, FLANN, Ceres and PCL are neither vendored
nor installed, so that build uses the stand-ins of oracle/ref_shim: the matcher is an EXACT kd-tree (OpenMP over the
queries, all host threads) instead of FLANN's approximate one, the 4M x 6 least squares a Gram-matrix SVD; every other
line (transforms with a 3x3 inverse per normal, weighting, rejection, gather, system assembly, pose composition) is
the reference's own, single-threaded as in the reference.  Without the library the arm falls back to the oracle port
(kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ITER = 30
WORKLOAD = "eth_apartment_shaped_pair_344x1077_knn_point_to_plane_linear_30it"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed ncu --set full capture
    (profiles/r1_ncu_traffic.json, produced by profiles/final_measure.sh); None when the summary is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")) as f:
            k = json.load(f)["kernels"]
        for name, v in k.items():
            if kernel_substr in name:
                return v["dram_bytes_read"] + v["dram_bytes_write"]
    except Exception:   # noqa: BLE001
        pass
    return None


def make_pair(pair_index=0, n_sweeps=344, n_beams=1077):
    """Synthetic ETH-shaped pair; cached under /tmp because k=5 PCA normals of 370k points take seconds."""
    from icp_variants_b200 import synth
    cache = f"/tmp/icp_b200_pair_{n_sweeps}x{n_beams}_{pair_index}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        return (synth.Cloud(z["sp"], z["sn"], z["sc"]), synth.Cloud(z["tp"], z["tn"], z["tc"]))
    src, tgt, _ = synth.eth_pair(seed=1234, n_sweeps=n_sweeps, n_beams=n_beams, pair_index=pair_index)
    try:
        np.savez(cache, sp=src.points, sn=src.normals, sc=src.colors, tp=tgt.points, tn=tgt.normals, tc=tgt.colors)
    except OSError:
        pass
    return src, tgt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(ns, nt, max_d2):
    """The workload both arms run (BASELINE.json configs[1])."""
    return {"workload": WORKLOAD, "n_source": ns, "n_target": nt, "iterations": N_ITER, "max_distance_sq": max_d2,
            "pairs_per_gpu_per_step": 1, "includes_index_build": True}


def ref_sample(src, tgt, max_d2, iters):
    """The reference's own LinearICPOptimizer::estimatePose (oracle/_ref) for `iters` iterations; None if unavailable."""
    try:
        from oracle import ref
        if ref.build() is None:
            return None
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
        t0 = time.perf_counter()
        n, pose, _ = ref.estimate_pose(0, 1, src.points, src.normals, src.colors, tgt.points, tgt.normals, tgt.colors,
                                       src.points[:4], tgt.points[:4], n_iterations=iters, max_distance_sq=max_d2)
        dt = time.perf_counter() - t0
        if n != iters:
            return None
        return dt, iters * len(src), os.cpu_count() or 1
    except Exception as e:   # noqa: BLE001 -- the baseline must never take the bench down
        print(f"bench: reference library unavailable ({e}); using the oracle port", file=sys.stderr)
        return None


def cpu_baseline_sample(src, tgt, max_d2, iters):
    """(seconds, queries, threads, kind): the reference build when present, else the oracle port."""
    r = ref_sample(src, tgt, max_d2, iters)
    if r is not None:
        return r + ("reference",)
    return cpu_sample(src, tgt, max_d2, iters) + ("port",)


def cpu_registration_seconds(src, tgt, max_d2, iters):
    """Seconds of one full N_ITER-iteration CPU registration, extrapolated from a 1-iteration and an `iters`-iteration run
    of the same pair: t(1) + (N_ITER - 1) * (t(iters) - t(1)) / (iters - 1), so that the once-per-registration work (index
    build, cloud copies) is counted once.  Returns (seconds, measured seconds, queries/iteration, threads, kind)."""
    iters = max(int(iters), 2)
    t1, _, cores, kind = cpu_baseline_sample(src, tgt, max_d2, 1)
    tk, nq, cores, kind = cpu_baseline_sample(src, tgt, max_d2, iters)
    per_iter = max(tk - t1, 0.0) / (iters - 1)
    return t1 + (N_ITER - 1) * per_iter, t1 + tk, nq // iters, cores, kind


def cpu_sample(src, tgt, max_d2, iters):
    """The oracle's whole registration loop for `iters` iterations on all host threads."""
    from oracle import oracle as orc
    orc.build()
    orc.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the baseline uses every host core
    cfg = orc.Config(metric=1, minimizer=0, max_distance_sq=max_d2, n_iterations=iters)
    t0 = time.perf_counter()
    rc, pose, hist, nq = orc.estimate_pose(cfg, src.points, src.normals, src.colors, tgt.points, tgt.normals, tgt.colors)
    dt = time.perf_counter() - t0
    return dt, nq, orc.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    src, tgt = make_pair(0, args.sweeps, args.beams)
    iters = args.cpu_iters
    if os.environ.get("OMP_NUM_THREADS") == "1":      # torchrun exports 1; the baseline may use every host core
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline_sample(src, tgt, args.max_dist2, 1)
    times, nq, cores, kind = [], 0, 1, "port"
    for _ in range(args.steps):
        dt, _, nq, cores, kind = cpu_registration_seconds(src, tgt, args.max_dist2, iters)
        times.append(dt)
    per_reg = statistics.mean(times)
    value = 1.0 / per_reg
    sample = (f"per step: a 1-iteration and a {max(iters, 2)}-iteration run of the same {len(src)}-point pair, extrapolated to {N_ITER} iterations "
              f"(index build counted once)")
    line = {"impl": "reference", "metric": "icp_registrations_per_s", "value": value, "unit": "reg/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_reg * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(len(src), len(tgt), args.max_dist2),
            "mcorr_per_s": nq * N_ITER / per_reg / 1e6,
            "cpu_baseline": {"value": value, "unit": "reg/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "reg/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": ("reference's own estimatePose compiled in place (oracle/_ref); Eigen/FLANN absent: exact kd-tree matcher on all host threads "
                     "instead of FLANN's approximate search, Gram-matrix SVD; the rest of the loop is the reference's single-threaded code"
                     if kind == "reference" else "oracle/_ref absent: oracle port, exact kd-tree instead of FLANN's approximate search")}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps", type=int, default=344)
    ap.add_argument("--beams", type=int, default=1077)
    ap.add_argument("--max-dist2", type=float, default=10.0, help="squared matching distance; alignETH uses 10 (main.cpp:361)")
    ap.add_argument("--cpu-iters", type=int, default=3, help="iterations per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from icp_variants_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the icp_gpu_* path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    src, tgt = make_pair(rank, args.sweeps, args.beams)       # each rank owns its own pair(s): no data-path collective
    ns, nt = len(src), len(tgt)

    ctx = capi.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)                        # torch.cuda.Event only sees torch's current stream
    cfg = capi.default_config()
    cfg.metric, cfg.minimizer, cfg.matching, cfg.n_iterations = 1, 0, 0, N_ITER
    cfg.max_distance_sq = args.max_dist2
    cfg.nn_algorithm = 2
    cfg.collect_stats = 0                                     # work counters cost device atomics: off while timing
    ctx.set_config(cfg)

    # device-resident raw clouds (reference layouts: packed float[3N], uint8[4N])
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         (("sp", src.points), ("sn", src.normals), ("sc", src.colors), ("tp", tgt.points), ("tn", tgt.normals), ("tc", tgt.colors))}
    # pinned host copies for the end-to-end arm
    h = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in
         (("sp", src.points), ("sn", src.normals), ("sc", src.colors), ("tp", tgt.points), ("tn", tgt.normals), ("tc", tgt.colors))}
    hn = {k: v.numpy() for k, v in h.items()}
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_resident():
        ctx.set_target_dev(d["tp"].data_ptr(), d["tn"].data_ptr(), d["tc"].data_ptr(), nt)
        ctx.set_source_dev(d["sp"].data_ptr(), d["sn"].data_ptr(), d["sc"].data_ptr(), ns)
        return ctx.estimate_pose(want_history=False)

    def step_e2e():
        ctx.set_target(hn["tp"], hn["tn"], hn["tc"])
        ctx.set_source(hn["sp"], hn["sn"], hn["sc"])
        return ctx.estimate_pose(want_history=False)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            if flush is not None:
                flush.fill_(1)
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms, launches, pose = 0.0, 0, None
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1)                                 # L2 flush between timed iterations (outside the events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            pose, _, n_it = fn()
            e1.record(stream)
            e1.synchronize()
            assert n_it == N_ITER
            total_ms += e0.elapsed_time(e1)
            launches += int(ctx.stats().n_kernel_launches)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
            l = torch.tensor([launches], dtype=torch.int64, device=dev)
            dist.all_reduce(l, op=dist.ReduceOp.SUM)
            launches = int(l.item())
        return total_ms, launches, pose

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, launches, pose = timed(step_resident, args.steps, args.warmup)
    e2e_ms, _, pose_e2e = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = launches // (args.steps * world)

    # per-kernel timing of the dominant kernel: one extra registration, launch by launch with CUDA events
    ctx.set_target_dev(d["tp"].data_ptr(), d["tn"].data_ptr(), d["tc"].data_ptr(), nt)
    ctx.set_source_dev(d["sp"].data_ptr(), d["sn"].data_ptr(), d["sc"].data_ptr(), ns)
    for _ in range(2):
        _, _, _, tm = ctx.estimate_pose(want_history=False, timings=True)
    # work counters (distance evaluations, staged points, deferred queries): one more, instrumented, registration
    cfg.collect_stats = 1
    ctx.set_config(cfg)
    ctx.set_target_dev(d["tp"].data_ptr(), d["tn"].data_ptr(), d["tc"].data_ptr(), nt)
    ctx.set_source_dev(d["sp"].data_ptr(), d["sn"].data_ptr(), d["sc"].data_ptr(), ns)
    ctx.estimate_pose(want_history=False)
    st_t = ctx.stats()
    st = st_t
    match_ms = tm.matching_ms / N_ITER
    prep_ms = tm.search_prep_ms / N_ITER
    walk_ms = max(match_ms - prep_ms, 1e-6)          # knn_bvh_kernel alone (the dominant kernel)
    solve_ms = tm.solver_ms / N_ITER

    if rank == 0:
        peak, peak_kind = load_peaks()
        regs = args.steps * world
        value = regs / (total_ms * 1e-3)
        # algorithmic bytes of the fused match kernel per query: source point 12 + source normal 12 +
        # matched target point 12 + target normal 12 = 48 B
        alg_bytes = 48.0 * ns
        achieved = alg_bytes / (walk_ms * 1e-3) / 1e9
        evals_per_launch = st_t.n_distance_evals / N_ITER
        line = {
            "metric": "icp_registrations_per_s", "value": value, "unit": "reg/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": dict(workload_config(ns, nt, args.max_dist2),
                           l2="flushed between timed steps (256 MB write)" if flush is not None else "not flushed"),
            "ms_per_registration": total_ms / args.steps,
            "mcorr_per_s": (st.n_queries * regs) / (total_ms * 1e-3) / 1e6,
            "e2e": {"value": regs / (e2e_ms * 1e-3), "unit": "reg/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int((ns + nt) * 28 + 64 + 32 * N_ITER), "d2h_bytes_per_step": int(64 + 16 * 4 * N_ITER + 1024)},
            "gpu_launches": launches,
            "roofline": {"kernel": "knn_bvh_kernel<false>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic("knn_bvh_kernel"), "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": walk_ms,
                         "note": "issue-bound tree search over L2-resident clouds; see roofline_fp32"},
            "roofline_fp32": {"kernel": "knn_bvh_kernel<false>", "distance_evals_per_launch": evals_per_launch,
                              "gevals_per_s": evals_per_launch / (match_ms * 1e-3) / 1e9, "flop_per_eval": 8,
                              "tflops": evals_per_launch * 8 / (match_ms * 1e-3) / 1e12, "kernels": "knn_prep_kernel (fast path) + knn_bvh_kernel (walk)", "nodes_per_launch": st_t.n_nodes_visited / N_ITER,
                              "matched_per_launch": st_t.n_matched / N_ITER},
            "stage_ms_per_iteration": {"match": match_ms, "match_prep_fast_path": prep_ms, "match_tree_walk": walk_ms, "reduce_solve": solve_ms, "index_build": tm.index_ms},
            "clocks": clocks,
            "pose_checksum": float(np.abs(pose).sum()),
        }
        if not args.no_cpu_baseline and world == 1:
            per_reg, dt, nq, cores, kind = cpu_registration_seconds(src, tgt, args.max_dist2, args.cpu_iters)
            line["cpu_baseline"] = {"value": 1.0 / per_reg, "unit": "reg/s", "cores": cores, "kind": kind,
                                    "sample": (f"a 1-iteration and a {max(args.cpu_iters, 2)}-iteration run of the same pair, extrapolated to {N_ITER} iterations "
                                               f"(index build counted once)"),
                                    "seconds": dt}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
