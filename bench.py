#!/usr/bin/env python
"""bench.py — headline benchmark of the LM-CMA trajectory-optimisation hot path (BASELINE.json metric:
trajectory-cost evals/s + LM-CMA generations/s) on the C2 workload: one 2-D query on a 4096x4096 synthetic
occupancy grid, 200 waypoints (n = 400), lambda = 1024, m = 2*sqrt(n) = 40.

A "step" is ONE LM-CMA generation = k_cost (lambda trajectory evaluations) -> k_rank -> k_update -> k_sample,
replayed from a CUDA graph; `value` = trajectory evaluations per second with everything
resident in HBM; L2 is flushed (256 MiB write) before every timed step.  N > 1: every rank optimises its own
independent query on its own GPU (weak scaling, no data-path collective: SURVEY.md section 8e).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = {"workload": "C2: 2-D 4096x4096 synthetic occupancy grid (seed 42, 2048 rectangles), one query "
                        "(64,64)->(4032,4032), 200 waypoints (n=400), lambda=1024, m=40, sigma0=32, longsafe weights",
            "map": "4096x4096 f32 sign-tagged reciprocal clearance (64 MiB)", "n": 400, "lambda": 1024, "m": 40,
            "waypoints": 200}
W, LAM, M, SIGMA0 = 200, 1024, 40, 32.0
# dram__bytes_read.sum + dram__bytes_write.sum of one k_cost launch of this workload, from the committed
# `ncu --set full` capture (the map stays L2-resident between generations, so DRAM traffic is far BELOW the
# algorithmic bytes: the kernel is bound by issue slots / L1 gather rate, not by HBM)
NCU_DRAM_BYTES_PER_LAUNCH = 2.378e6
NCU_SOURCE = "profiles/r1f_full.md (k_cost<2,0,0>: dram_read 2.378 MB, dram_write 0)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(seed_offset=0):
    from lmcma_path_planner_b200 import maps
    dist, start, goal = maps.config2_map()
    lo, hi = maps.box_bounds((4096, 4096), W)
    x0 = maps.straight_line(start, goal, W)
    return dist, start, goal, lo, hi, x0


def run_b200(args):
    import torch
    import lmcma_path_planner_b200 as L
    from lmcma_path_planner_b200 import _capi as K

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist_pg = None
    if world > 1:
        import torch.distributed as dist_pg
        dist_pg.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist_pg is not None:
            dist_pg.barrier()
        torch.cuda.synchronize()

    dist, start, goal, lo, hi, x0 = build_problem()
    cmap = L.CostMap(dist, "f32", device=local)
    opt = L.Optimizer(2 * W, x0=x0, lam=LAM, m=M, lo=lo, hi=hi, sigma0=SIGMA0, seed=1000 + rank, rng="philox", device=local)
    opt.attach_cost(cmap, [start], [goal], W, L.LONGSAFE, 1e4)
    stream = torch.cuda.Stream()                     # a real (non-null) stream: the library enqueues on it
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    opt.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def one_step(timed):
        flush.fill_(1)                               # evict L2 (not timed)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(stream)
        opt.run(1, sync=False)
        e1.record(stream)
        return (e0, e1) if timed else None

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = K.lib().lmcma_b200_launch_count()
    t_wall0 = time.perf_counter()
    evs = [one_step(True) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = K.lib().lmcma_b200_launch_count() - launches0
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    nsamp_mean = float(opt.get("nsamp").mean())
    sigma_now = float(opt.get("sigma")[0])

    # steady state: back-to-back graph replays, L2 warm (the deployment mode: the map stays L2-resident)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(stream)
    opt.run(args.steps, sync=False)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_warm = e0.elapsed_time(e1)

    # per-kernel durations (CUDA events around every kernel), same flush policy as the timed region
    per_kernel = {}
    nsamp_k = []
    for _ in range(3):                               # un-graphed launch path: first use loads these kernel variants
        opt.profile_kernels(1)
    for _ in range(args.steps):
        flush.fill_(1)
        pk = opt.profile_kernels(1)
        nsamp_k.append(float(opt.get("nsamp").mean()))
        for k, v in pk.items():
            per_kernel[k] = per_kernel.get(k, 0.0) + v / args.steps
    clocks = sampler.stop()

    # end to end through the reference-facing protocol with HOST buffers: ask_all (D2H) -> cost_evaluate
    # (H2D, kernel, D2H) -> tell_all (H2D, update + sample), pinned host memory
    e2e_opt = L.Optimizer(2 * W, x0=x0, lam=LAM, m=M, lo=lo, hi=hi, sigma0=SIGMA0, seed=2000 + rank, rng="philox", device=local)
    Xh = torch.empty((LAM, 2 * W), dtype=torch.float32).pin_memory().numpy()
    fh = torch.empty(LAM, dtype=torch.float32).pin_memory().numpy()
    nch = torch.empty(LAM, dtype=torch.int32).pin_memory().numpy()
    nsh = torch.empty(LAM, dtype=torch.int32).pin_memory().numpy()
    import ctypes as C
    from lmcma_path_planner_b200.optimizer import _endpoints, _objective
    obj, ends = _objective(W, L.LONGSAFE, 1e4), _endpoints(start, goal)

    e2e_parts = [0.0, 0.0, 0.0]                                  # host wall clock inside ask_all / cost_evaluate / tell_all

    def e2e_step():
        t0 = time.perf_counter()
        K.check(K.lib().lmcma_b200_ask_all(e2e_opt._h, K.fptr(Xh)))
        t1 = time.perf_counter()
        K.check(K.lib().lmcma_b200_cost_evaluate(cmap._h, C.byref(obj), C.byref(ends), K.fptr(Xh), LAM, K.fptr(fh),
                                                K.iptr(nch), K.iptr(nsh)))
        t2 = time.perf_counter()
        K.check(K.lib().lmcma_b200_tell_all(e2e_opt._h, K.fptr(fh)))
        t3 = time.perf_counter()
        e2e_parts[0] += t1 - t0; e2e_parts[1] += t2 - t1; e2e_parts[2] += t3 - t2
        return float(fh[0])

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    barrier()
    e2e_t = 0.0
    e2e_parts[:] = [0.0, 0.0, 0.0]
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()
        e2e_t += time.perf_counter() - t0
    h2d = LAM * 2 * W * 4 + LAM * 4
    d2h = LAM * 2 * W * 4 + LAM * 12

    # max over ranks
    if dist_pg is not None:
        t = torch.tensor([ms_total, ms_warm, e2e_t], dtype=torch.float64, device="cuda")
        dist_pg.all_reduce(t, op=dist_pg.ReduceOp.MAX)
        ms_total, ms_warm, e2e_t = (float(v) for v in t.cpu())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist_pg.all_reduce(lt)
        launches = int(lt.item())

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        ms_step = ms_total / args.steps
        evals_per_s = world * LAM * args.steps / (ms_total * 1e-3)
        # cost kernel roofline: algorithmic bytes per trajectory = 4n (candidate) + S*b (map samples) + 8 (f, flag)
        S = float(np.mean(nsamp_k)) if nsamp_k else nsamp_mean
        bytes_per_launch = LAM * (4 * 2 * W + S * cmap.bytes_per_cell + 8)
        cost_ms = per_kernel["cost"]
        achieved = bytes_per_launch / (cost_ms * 1e-3) / 1e9
        out = {
            "metric": "trajectory-cost evals/s (LM-CMA generations/s = value / lambda)", "value": evals_per_s, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(WORKLOAD, parallelism="independent query per GPU (no collective)" if world > 1 else "1 GPU",
                           l2="flushed before every timed step (256 MiB fill)", rng="device Philox4x32-10",
                           mean_samples_per_trajectory=S, sigma_after_timed_region=sigma_now),
            "generations_per_s": world * args.steps / (ms_total * 1e-3),
            "steady_state_l2_warm": {"value": world * LAM * args.steps / (ms_warm * 1e-3), "unit": "evals/s",
                                     "ms_per_step": ms_warm / args.steps,
                                     "note": "back-to-back graph replays, map L2-resident (deployment mode)"},
            "kernel_ms": per_kernel,
            "roofline": {"kernel": "k_cost", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_SOURCE,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": cost_ms,
                         "note": "bytes = lambda*(4n + S*4 + 8), S = mean map samples per trajectory; duration = CUDA events "
                                 "around k_cost on the launching stream, L2 flushed before each generation"},
            "e2e": {"value": world * LAM * args.steps / e2e_t, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_t / args.steps * 1e3,
                    "calls_ms_rank0": {k: v / args.steps * 1e3 for k, v in zip(("ask_all", "cost_evaluate", "tell_all"), e2e_parts)},
                    "path": "lmcma_b200_ask_all (D2H X) -> lmcma_b200_cost_evaluate (H2D X, D2H f/flags) -> lmcma_b200_tell_all "
                            "(H2D f; update+sample), pinned host buffers, wall clock around synchronous calls"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        out["e2e"]["path"] += "; page-locked buffers are read / written by k_cost directly (no staging copy)"
        if world == 1:
            out["roofline_batched_queries"] = batched_cost_roofline(dist, cmap, lo, hi, peak, local)
        out["cpu_baseline"] = cpu_baseline(dist, start, goal, lo, hi, x0, budget_s=12.0)
        print(json.dumps(out))
    if dist_pg is not None:
        dist_pg.barrier()
        dist_pg.destroy_process_group()
    return out


def batched_cost_roofline(dist, cmap, lo, hi, peak, local, queries=256, lam=64, gens=12):
    """Explanatory extra (not the headline workload): the same k_cost on a C3-shaped batch — `queries` independent
    start/goal queries x lambda 64 on the C2 map — where many waves of CTAs overlap the per-trajectory phases that a
    single 1024-trajectory query runs in lock-step.  Same algorithmic-bytes definition, CUDA events around k_cost."""
    import lmcma_path_planner_b200 as L
    from lmcma_path_planner_b200 import maps
    starts, goals = maps.random_queries(dist, queries, seed=7, min_sep=1024)
    x0b = np.stack([maps.straight_line(starts[q], goals[q], W) for q in range(queries)])
    opt = L.Optimizer(2 * W, x0=x0b, lam=lam, m=M, batch=queries, lo=lo, hi=hi, sigma0=SIGMA0, seed=7, device=local)
    opt.attach_cost(cmap, starts, goals, W, L.LONGSAFE, 1e4)
    opt.run(5)
    cost_ms, nsamp = 0.0, 0.0
    for _ in range(gens):
        cost_ms += opt.profile_kernels(1)["cost"] / gens
        nsamp += float(opt.get("nsamp").mean()) / gens
    nbytes = queries * lam * (4 * 2 * W + nsamp * cmap.bytes_per_cell + 8)
    achieved = nbytes / (cost_ms * 1e-3) / 1e9
    return {"kernel": "k_cost", "workload": "%d queries x lambda %d (C3-shaped batch, one launch), L2 warm" % (queries, lam),
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "algorithmic_bytes_per_launch": nbytes, "launch_ms": cost_ms, "mean_samples_per_trajectory": nsamp,
            "samples_per_s": queries * lam * nsamp / (cost_ms * 1e-3)}


def cpu_baseline(dist, start, goal, lo, hi, x0, budget_s, steps=None, warmup=0):
    """The reference's CPU implementation of the path on the host cores: the UNMODIFIED reference LMCMA class
    (oracle/_ref, single serial stream as shipped; m = lambda is its only rule) driven through its ask/tell
    protocol, with the cost restatement (oracle/cost_oracle.c) as the user cost evaluated over the population
    on all host threads.  Bounded sample: generations of the C2 query until `budget_s` (or `steps`)."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    prob = po.CostProblem(dist, start, goal, W, 1.0, 1000.0, 1e4, threads=cores)
    kind = "reference"
    try:
        opt = po.RefLMCMA(2 * W, x0=x0, lam=LAM, lo=lo, hi=hi, sigma=SIGMA0, seed=1)
        step = lambda: opt.generation(prob)
    except Exception:
        kind = "port"
        opt = po.OracleLMCMA(2 * W, x0=x0, lam=LAM, m=M, lo=lo, hi=hi, sigma=SIGMA0, seed=1)

        def step():
            X = opt.array("X")
            opt.tell_all(prob.evaluate(X)["f"])
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    gens = 0
    per = []
    while True:
        t1 = time.perf_counter()
        step()
        per.append(time.perf_counter() - t1)
        gens += 1
        if steps is not None:
            if gens >= steps:
                break
        elif time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": LAM * gens / dt, "unit": "evals/s", "cores": cores, "kind": kind,
            "sample": "%d generations of the C2 query (lambda=1024, n=400) in %.1f s: reference LMCMA (oracle/_ref, serial, "
                      "m=lambda so at most %d live pairs) + cost restatement on %d threads" % (gens, dt, gens + warmup, cores),
            "ms_per_step": dt / gens * 1e3, "generations": gens}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    dist, start, goal, lo, hi, x0 = build_problem()
    cb = cpu_baseline(dist, start, goal, lo, hi, x0, budget_s=None, steps=args.steps, warmup=args.warmup)
    out = {"impl": "reference", "metric": "trajectory-cost evals/s (LM-CMA generations/s = value / lambda)",
           "value": cb["value"], "unit": "evals/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(WORKLOAD, parallelism="host CPU"),
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
