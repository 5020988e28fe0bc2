"""C3 workload (BASELINE.json configs[2]): 4096 independent 2-D queries batched on the C2 map, sharded over the
ranks of a torchrun launch (no data-path collective).  Prints evals/s and per-kernel times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lmcma_path_planner_b200 as L
from lmcma_path_planner_b200 import maps, parallel

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 50
storage = sys.argv[3] if len(sys.argv) > 3 else "f32"     # "u8": distance clamped at 63 cells, quantised to 1/4 cell (a different map)
W, lam, m = 200, 64, 40
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dmap, _, _ = maps.config2_map()
starts, goals = maps.random_queries(dmap, Q, seed=7, min_sep=1024)
off, cnt = parallel.shard_range(Q, world, rank)
cmap = L.CostMap(dmap if storage == "f32" else np.minimum(dmap, 63.0), storage, device=local)
lo, hi = maps.box_bounds((4096, 4096), W)
x0 = np.stack([maps.straight_line(starts[q], goals[q], W) for q in range(off, off + cnt)])
opt = L.Optimizer(2 * W, x0=x0, lam=lam, m=m, batch=cnt, lo=lo, hi=hi, sigma0=32.0, seed=7 + rank, device=local)
opt.attach_cost(cmap, starts[off:off + cnt], goals[off:off + cnt], W, L.LONGSAFE, 1e4)
opt.run(5)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
opt.run(gens)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ms = opt.last_run_ms()
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
pk = opt.profile_kernels(3)
if rank == 0:
    print("C3 (%s map): %d queries x lambda %d on %d GPU(s): %.3f ms/generation (device), %.3g evals/s, %.1f query-generations/s" %
          (storage, Q, lam, world, ms / gens, Q * lam * gens / (ms * 1e-3), Q * gens / (ms * 1e-3)))
    print("per-kernel ms:", {k: round(v, 4) for k, v in pk.items()}, "mean nsamp", float(opt.get("nsamp").mean()))
    f0 = opt.best()[1]
    print("best f: min %.4g median %.4g" % (f0.min(), np.median(f0)))
if world > 1: dist.destroy_process_group()
