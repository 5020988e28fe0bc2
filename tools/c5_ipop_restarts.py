"""C5 workload (BASELINE.json configs[4]): IPOP-style restarts on the cluttered 2-D map (C2's map with 4x the
obstacle count): one query, lambda = lambda0 * 2^r up to lambda_max, restart r seeded 1000 + r, each restart runs until
sigma < 1e-20 (lmcma.cpp:428) or `max_gens` generations.  Restarts are independent jobs: they are assigned to the ranks
of a torchrun launch largest first (parallel.assign_largest_first; a restart costs ~lambda), no data-path collective;
the best (f, path) over all restarts is gathered at the end.

  python tools/c5_ipop_restarts.py [lambda_max] [max_gens] [lambda0]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/c5_ipop_restarts.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lmcma_path_planner_b200 as L
from lmcma_path_planner_b200 import maps, parallel

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
lam_max = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
max_gens = int(sys.argv[2]) if len(sys.argv) > 2 else 300
lam0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
W, m = 200, 40
size = int(sys.argv[4]) if len(sys.argv) > 4 else 4096          # map edge (tests use a small one)
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dmap, start, goal = maps.config2_map(size=size, n_rects=4 * 2048, seed=42)        # "cluttered": 4x the obstacle count
cmap = L.CostMap(dmap, "f32", device=local)
lo, hi = maps.box_bounds((size, size), W)
x0 = maps.straight_line(start, goal, W)
lams = parallel.ipop_schedule(lam0, lam_max)
mine = parallel.assign_largest_first([float(l) for l in lams], world)[rank]
torch.cuda.synchronize()
if dist is not None: dist.barrier()
t0 = time.perf_counter()
best_f, best_r, evals, log = float("inf"), -1, 0, []
for r in mine:
    lam = lams[r]
    opt = L.Optimizer(2 * W, x0=x0, lam=lam, m=m, lo=lo, hi=hi, sigma0=32.0, seed=1000 + r, device=local)
    opt.attach_cost(cmap, [start], [goal], W, L.LONGSAFE, 1e4)
    gens = 0
    while gens < max_gens:
        step = min(25, max_gens - gens)
        opt.run(step)                       # fused generations, no host round trip inside
        gens += step
        if float(opt.get("sigma")[0]) < 1e-20: break          # isBehaviorLearningDone (checked every 25 generations)
    xb, fb = opt.best()
    f = float(fb[0])
    nc = int(cmap.evaluate(xb[0], start, goal, W, L.LONGSAFE, 1e4)["ncoll"][0])    # colliding samples of the best path found
    evals += lam * gens
    log.append((r, lam, gens, f, float(opt.get("sigma")[0]), nc, time.perf_counter() - t0))
    if f < best_f: best_f, best_r = f, r
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if dist is not None:
    t = torch.tensor([dt, float(evals), best_f, float(best_r)], dtype=torch.float64, device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    rows = torch.stack(allt).cpu().numpy()
    dt, evals = float(rows[:, 0].max()), float(rows[:, 1].sum())
    k = int(np.argmin(rows[:, 2])); best_f, best_r = float(rows[k, 2]), int(rows[k, 3])
for r, lam, gens, f, sg, nc, tdone in log:
    print("  rank %d restart %d: lambda %6d, %3d generations, best f %.6g (%d colliding samples), sigma %.3g, done at %.3f s" %
          (rank, r, lam, gens, f, nc, sg, tdone), flush=True)
print("  rank %d wall clock %.3f s for restarts %s" % (rank, time.perf_counter() - t0, [r for r, *_ in log]), flush=True)
if dist is not None: dist.barrier()
if rank == 0:
    print("C5: %d restarts (lambda %d..%d) on %d GPU(s): %.3f s wall (max over ranks), %.3g evals, %.3g evals/s, best f %.6g from restart %d" %
          (len(lams), lams[0], lams[-1], world, dt, evals, evals / dt, best_f, best_r))
if dist is not None: dist.destroy_process_group()
