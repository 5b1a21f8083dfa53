#!/usr/bin/env python
"""bench.py — headline benchmark of the dftpav_b200 hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm  (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm = the reference's CPU algorithm on host cores

metric : full L-BFGS trajectory solves/s at batch = 4096 (BASELINE.json configs[2]: 24 pieces, per-sample rectangle
         corridor, curvature + vel/acc limits active).  One "step" = one pass of the hot path over one batch of 4096
         synthetic problems per GPU (weak scaling: every rank solves its own seeded batch, no data-path collective;
         one NCCL all-reduce of the 1-int convergence flag per step).
value  : solves/s with the batch already resident in HBM (dftpav_optimize only), device-timed with CUDA events.
e2e    : the same metric through the planner-facing call dftpav_solve (pinned host buffers in, H2D + layout kernels +
         solve + D2H of the results inside the timed region).
The CPU oracle (oracle/) is used here only for the cpu_baseline leg and for --impl reference.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "trajectory solves/sec (L-BFGS to tol) at batch=4096"
UNIT = "solves/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "_refbuild"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json config id (3 = the config the metric is quoted on)")
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems of the batch used for the CPU legs (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes read+write of the solve kernel from the committed ncu summary (per launch), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_summary_latest.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return d.get("solve_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def pin_batch(batch):
    """Moves every array of the batch into pinned host memory (so the e2e leg copies from pinned buffers)."""
    import torch

    keep = []
    for name in ("seg_ptr", "piece_num", "singul", "ini_state", "fin_state", "init_T", "inner_pts", "planes", "t_now",
                 "plane_cnt", "obs_ptr", "obs_idx", "obs_start_time", "obs_duration", "obs_piece_ptr", "obs_piece_dur",
                 "obs_coeffs"):
        a = getattr(batch, name)
        if a is None:
            continue
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(t)
        setattr(batch, name, t.numpy())
    batch.meta["_pinned"] = keep
    return batch


def cpu_leg(prm, batch, sample: int, threads: int, reps: int = 1, warm: bool = False):
    """Times the CPU oracle (the reference's algorithm restated, oracle/) on `sample` problems with `threads` threads.
    warm: one untimed pass over a few problems per thread first (page-in of the library, OpenMP team start-up, CPU clocks).
    Returns (solves/s of the MEDIAN pass, median seconds, solution of the last pass)."""
    from oracle import oracle_py

    oracle_py.build()
    if warm:
        oracle_py.optimize(prm, batch, 0, min(sample, 2 * threads), threads=threads)
    times = []
    sol = None
    for _ in range(max(1, reps)):
        t0 = time.perf_counter()
        sol = oracle_py.optimize(prm, batch, 0, sample, threads=threads)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return sample / med, med, sol


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    from dftpav_b200 import synth

    batch, prm = synth.make_config(args.config, B=args.batch, seed=0)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or min(args.batch, max(64, 64 * cores))
    for _ in range(args.warmup):
        cpu_leg(prm, batch, min(sample, 4 * cores), cores)
    times = []
    for _ in range(args.steps):
        _, dt, _ = cpu_leg(prm, batch, sample, cores)
        times.append(dt)
    total = sum(times)
    value = sample * args.steps / total
    ref_build = reference_build_leg(args, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config {args.config}: {batch.meta.get('description')}", "batch_per_gpu": args.batch,
                   "pieces": batch.meta.get("N"), "note": "CPU arm: each step solves a bounded sample of the same batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample} problems of the seeded batch per step, OpenMP over problems, one thread "
                                   f"per problem like the single-threaded reference",
                         "note": "the port is bit-identical to the reference's own sources compiled into oracle/_ref "
                                 "(tests/test_oracle_vs_reference_optimizer.py) and is the FASTER of the two CPU builds "
                                 "available here (see reference_build), so it is the one timed"},
        "reference_build": ref_build,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def reference_build_leg(args, cores):
    """The reference's OWN optimizer source compiled in place (oracle/_ref/libref_optimizer.so, test-only Eigen layer) on a
    small sample: reported for completeness - its eager matrix layer allocates per expression and runs ~100x slower than
    the bit-identical port, so timing it as THE baseline would flatter the GPU arm.  Runs in a SUBPROCESS with a timeout:
    whatever happens to it (round 1: a data race in the driver's cout sink aborted the process) cannot take the
    reference arm's result line with it."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "_refbuild", "--config", str(args.config),
                              "--batch", str(args.batch)], capture_output=True, text=True, timeout=300)
        for ln in reversed(res.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"available": False, "error": f"rc {res.returncode}: {res.stderr.strip()[-160:]}"}
    except Exception as e:  # never let the auxiliary leg break the bench line
        return {"available": False, "error": str(e)[:200]}


def run_refbuild(args):
    """Child process of reference_build_leg."""
    from dftpav_b200 import synth
    from oracle import oracle_py as O

    cores = os.cpu_count() or 1
    if O.ref_optimizer_lib() is None:
        print(json.dumps({"available": False}))
        return
    batch, prm = synth.make_config(args.config, B=args.batch, seed=0)
    m = min(batch.B, 2 * cores)
    t0 = time.perf_counter()
    r = O.ref_solve_batch(prm, batch, 0, m, threads=cores)
    dt = time.perf_counter() - t0
    port = O.optimize(prm, batch, 0, m, threads=cores)
    same = bool(np.array_equal(r["cost"], port.cost[:m]) and np.array_equal(r["iters"], port.iters[:m])
                and np.array_equal(r["status"], port.status[:m]))
    print(json.dumps({"available": True, "value": m / dt, "unit": UNIT, "cores": cores,
                      "sample": f"first {m} problems, one pass", "bit_identical_to_port": same}))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the dftpav_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        sys.stderr.write(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}\n")

    from dftpav_b200 import synth
    from dftpav_b200.optimizer import BatchOptimizer

    # weak scaling: every rank owns an independent seeded batch (problems are independent)
    batch, prm = synth.make_config(args.config, B=args.batch, seed=rank)
    pin_batch(batch)
    opt = BatchOptimizer(prm, device=local_rank)
    # a dedicated (non-default) torch stream: the library launches on it, the CUDA events below are recorded on it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    opt.set_stream(stream.cuda_stream)
    flag = torch.ones(1, dtype=torch.int32, device="cuda")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def convergence_flag(sol_ok: int):
        # the single collective of the path: global "every rank's shard finished and every solve was accepted" flag
        # (1 int, NCCL over NVLink; MIN over ranks of the local flag)
        flag.fill_(sol_ok)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    opt.upload(batch)
    sol = opt.new_solution(pinned=True)

    # ---------------- value: batch resident in HBM, dftpav_optimize only
    for _ in range(max(args.warmup, 3)):
        opt.optimize()
        convergence_flag(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        opt.optimize()
        st = opt.stats()
        kernel_ms.append(st.solve_ms)
        launches += int(st.kernel_launches - st.n_shapes)  # per shape: f(x0) pass + rank kernel (queue order) + the persistent solve kernel
        convergence_flag(int(opt.success_count() == args.batch))  # the real local flag (16 KB read-back), then MIN over ranks
    ev1.record(stream)
    sync_all()
    dev_ms = ev0.elapsed_time(ev1)
    all_accepted = int(flag.item())
    opt.download(sol)
    st = opt.stats()
    alg_bytes = int(st.algorithmic_bytes)
    total_evals, total_iters = int(st.total_evals), int(st.total_iters)
    success = int(sol.success.sum())

    # ---------------- e2e: pinned host batch in, results out, every step (ONE C-ABI call: dftpav_solve)
    for _ in range(2):
        opt.solve(batch, sol)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        opt.solve(batch, sol)
        convergence_flag(int(bool(np.all(sol.success != 0))))
    e1.record(stream)
    sync_all()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()  # sampled over both timed regions (value and e2e legs)
    st2 = opt.stats()
    e2e_launches = st2.kernel_launches
    e2e_breakdown = {"small_arrays_and_state_kernel_ms": st2.upload_ms, "solver_kernels_incl_waiting_for_plane_chunks_ms": st2.solve_ms,
                     "result_readback_ms": st2.download_ms}

    # max over ranks (device-timed); per-rank kernel times to name a straggler
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([args.batch * args.steps], dtype=torch.float64, device="cuda")
    km = torch.tensor([statistics.mean(kernel_ms), dev_ms / args.steps], dtype=torch.float64, device="cuda")
    km_all = [torch.zeros_like(km) for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_gather(km_all, km)
    else:
        km_all = [km]
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_solves = float(cnt[0])
    rank_kernel_ms = [float(x[0]) for x in km_all]
    rank_step_ms = [float(x[1]) for x in km_all]

    if rank == 0:
        peaks, peak_src = measured_peaks()
        kms = statistics.mean(kernel_ms)
        achieved = alg_bytes / (kms * 1e-3) / 1e9  # GB/s, algorithmic bytes of one launch / its duration
        peak = float(peaks.get("hbm_gbs", 6650.0))
        value = total_solves / (dev_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE config {args.config}: {batch.meta.get('description')}",
                       "batch_per_gpu": args.batch, "pieces": batch.meta.get("N"), "parallelism": f"dp{world} (independent shards)",
                       "l2": "inputs larger than L2 (corridor planes %.0f MB per GPU)" % (batch.planes.nbytes / 1e6),
                       "lbfgs": "mem 256, past 3, delta 1e-4 (reference defaults)"},
            "clocks": clocks,
            "e2e": {"value": total_solves / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(batch.nbytes()),
                    "d2h_bytes_per_step": int(sol.nbytes()), "ms_per_step": e2e_ms / args.steps,
                    "last_step_breakdown_rank0": e2e_breakdown,
                    "note": "one C-ABI call per step (dftpav_solve): the corridor planes are copied in 32 chunks on a second "
                            "stream while the persistent solver kernel already runs; the longest-predicted-first queue is "
                            "keyed by the plane-free part of f(x0) and only hands out problems whose chunk has arrived"},
            "gpu_launches": int(launches),
            "gpu_launches_e2e_per_step": int(e2e_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel": "solve_kernel<false>", "kernel_ms": kms},
            "evals_per_s": total_evals / (kms * 1e-3) * world,
            "iters_mean": total_iters / args.batch, "evals_mean": total_evals / args.batch,
            "success_frac": success / args.batch,
            "convergence_flag": {"all_ranks_all_accepted": all_accepted,
                                 "note": "MIN over ranks of (success_count == batch); 0 when any solve of any shard ends "
                                         "with a rejected L-BFGS code or a cost above the limit (traj_optimizer.cpp:176-201)"},
            "per_rank": {"kernel_ms": {"min": min(rank_kernel_ms), "median": statistics.median(rank_kernel_ms),
                                       "max": max(rank_kernel_ms), "all": [round(v, 3) for v in rank_kernel_ms]},
                         "step_ms": {"min": min(rank_step_ms), "max": max(rank_step_ms)},
                         "note": "every rank solves its own seeded batch (weak scaling); ms_per_step is the max over ranks"},
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.cpu_sample or min(args.batch, max(64, 64 * cores))
            v, dt, ref = cpu_leg(prm, batch, sample, cores, reps=3, warm=True)
            rel = np.abs(sol.cost[:sample] - ref.cost[:sample]) / np.maximum(1.0, np.abs(ref.cost[:sample]))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} problems of rank 0's batch, warmed, median of 3 passes "
                                              f"({dt:.2f} s each), OpenMP over problems"}
            line["parity_vs_oracle"] = {"problems": int(sample), "frac_cost_within_1e-5": float(np.mean(rel <= 1e-5)),
                                        "median_rel": float(np.median(rel)), "max_rel": float(rel.max()),
                                        "status_equal_frac": float(np.mean(sol.status[:sample] == ref.status[:sample]))}
        print(json.dumps(line))
    opt.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "_refbuild":
        run_refbuild(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
