"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two ways the path shards:

* independent units — planning queries (C3) or IPOP restarts (C5): disjoint work lists, the map replicated,
  NO data-path collective; only a final gather of (best cost, path).  ``shard_range`` /
  ``assign_largest_first``.
* one huge population (C4) — every rank owns lambda/G offspring rows; optimiser state is replicated and
  advanced redundantly and deterministically on every rank; per generation the lambda fitness scalars are
  all-gathered, then one (n+4)-float payload per rank (local weighted partial sums + the local count of
  the merged-ranking statistic), because with box bounds active the recombination is not linear in z
  (SURVEY.md appendix B.10).  ``SplitPopulation``.
"""
import numpy as np


def shard_range(total, world, rank):
    """Contiguous shard [offset, offset + count) of `total` units for `rank` (first ranks get the remainder)."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def assign_largest_first(costs, world):
    """Greedy longest-processing-time assignment of independent jobs (IPOP restarts whose cost doubles with
    lambda): jobs sorted by decreasing cost, each to the currently least-loaded rank.  Returns per-rank lists
    of job indices (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    jobs = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        jobs[r].append(i)
        load[r] += float(costs[i])
    return jobs


def ipop_schedule(lambda0, lambda_max):
    """lambda doubling per restart (IPOP): lambda0, 2*lambda0, ... <= lambda_max."""
    out, lam = [], int(lambda0)
    while lam <= lambda_max:
        out.append(lam)
        lam *= 2
    return out


class DeviceBackend:
    """Adapter from torch tensors to the C ABI's raw device pointers (lmcma_b200_mg_*)."""

    def __init__(self, opt, stream_ptr=None):
        """stream_ptr: the cudaStream_t (int) the three stages are enqueued on.  None = torch's CURRENT stream at the time
        of each call, i.e. the stream the exchanges of SplitPopulation are ordered with — never the optimiser's private
        stream, which nothing would order against the all-gathers."""
        self.opt, self._stream = opt, stream_ptr
        self.payload_floats = opt.mg_payload_floats()
        self.pop_count, self.lam = opt.pop_count, opt.lam

    @property
    def stream(self):
        if self._stream is not None:
            return self._stream
        import torch
        return torch.cuda.current_stream().cuda_stream

    def evaluate(self, f_local):
        self.opt.mg_evaluate(f_local.data_ptr(), self.stream)

    def rank(self, f_all, payload):
        self.opt.mg_rank(f_all.data_ptr(), payload.data_ptr(), self.stream)

    def update(self, payload_all, world):
        self.opt.mg_update(payload_all.data_ptr(), world, self.stream)


class SplitPopulation:
    """One LM-CMA population split over the ranks of a process group.  `backend` supplies the three device
    stages; this class owns the two per-generation exchanges and their buffers."""

    def __init__(self, backend, dist, device, group=None):
        import torch
        self.torch, self.dist, self.group, self.backend = torch, dist, group, backend
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.rank = dist.get_rank(group) if dist is not None else 0
        if backend.pop_count * self.world != backend.lam:
            raise ValueError("lambda (%d) must split evenly over %d ranks" % (backend.lam, self.world))
        pf = backend.payload_floats
        self.f_local = torch.zeros(backend.pop_count, dtype=torch.float32, device=device)
        self.f_all = torch.zeros(backend.lam, dtype=torch.float32, device=device)
        self.payload = torch.zeros(pf, dtype=torch.float32, device=device)
        self.payload_all = torch.zeros(self.world * pf, dtype=torch.float32, device=device)

    def _all_gather(self, out, inp):
        if self.dist is None or self.world == 1:
            out.copy_(inp)
        elif self.dist.get_backend(self.group) == "nccl":
            self.dist.all_gather_into_tensor(out, inp, group=self.group)
        else:
            self.dist.all_gather(list(out.chunk(self.world)), inp, group=self.group)

    def generation(self):
        b = self.backend
        b.evaluate(self.f_local)
        self._all_gather(self.f_all, self.f_local)          # exchange 1: lambda fitness scalars
        b.rank(self.f_all, self.payload)
        self._all_gather(self.payload_all, self.payload)    # exchange 2: G x (n+4) floats
        b.update(self.payload_all, self.world)

    def capture(self):
        """Capture one generation (3 device stages + 2 exchanges) into ONE CUDA graph on the current stream; run() then
        replays it: one launch per generation instead of five enqueue calls and two collective launches from Python.
        Returns False (and keeps the eager path) where the capture is not possible (e.g. a gloo group)."""
        t = self.torch
        self._graph = None
        if self.dist is not None and self.world > 1 and self.dist.get_backend(self.group) != "nccl":
            return False
        try:
            g = t.cuda.CUDAGraph()
            t.cuda.synchronize()
            with t.cuda.graph(g, stream=t.cuda.current_stream()):
                self.generation()
            self._graph = g
            return True
        except Exception as e:                       # keep going eagerly; the caller reports which path ran
            self.capture_error = str(e)
            self._graph = None
            return False

    def profile_stages(self, generations):
        """Mean device time (ms) of the five stages of the EAGER generation, CUDA events on the current stream."""
        t = self.torch
        b, acc = self.backend, [0.0] * 5
        for _ in range(generations):
            e = [t.cuda.Event(enable_timing=True) for _ in range(6)]
            e[0].record(); b.evaluate(self.f_local)
            e[1].record(); self._all_gather(self.f_all, self.f_local)
            e[2].record(); b.rank(self.f_all, self.payload)
            e[3].record(); self._all_gather(self.payload_all, self.payload)
            e[4].record(); b.update(self.payload_all, self.world)
            e[5].record()
            t.cuda.synchronize()
            for k in range(5):
                acc[k] += e[k].elapsed_time(e[k + 1]) / generations
        return acc

    def run(self, generations):
        g = getattr(self, "_graph", None)
        for _ in range(generations):
            if g is not None:
                g.replay()
            else:
                self.generation()
