"""CPU, world_size 2, gloo: the multi-GPU host logic — shard arithmetic and the split-population exchange
protocol (two all-gathers per generation) with a numpy stand-in for the three device stages.  The stand-in
re-derives LM-CMA's update from the gathered payloads exactly as k_update does, and must track the unsplit
FP64 oracle."""
import os
import socket

import numpy as np
import pytest

from lmcma_path_planner_b200 import parallel


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 64, 4096, 8192):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for off, cnt in spans:
                assert off == pos
                pos += cnt
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_ipop_schedule_and_largest_first():
    lams = parallel.ipop_schedule(1024, 65536)
    assert lams == [1024 * 2 ** r for r in range(7)]
    jobs = parallel.assign_largest_first(lams, 8)
    assert sorted(i for j in jobs for i in j) == list(range(7))
    assert jobs[0] == [6]                                   # the 65536 restart gets a GPU to itself
    jobs2 = parallel.assign_largest_first(lams, 2)
    loads = [sum(lams[i] for i in j) for j in jobs2]
    assert max(loads) == 65536                              # LPT: {65536} vs everything else (65024)
    assert parallel.assign_largest_first([3, 3, 3], 1) == [[0, 1, 2]]


class NumpySplitBackend:
    """CPU stand-in for lmcma_b200_mg_{evaluate,rank,update}: same payload layout (n_ld + 4 floats: partial
    weighted sum of (x - xmean) | S low word | S high word | 0 | 0), FP32 payload like the device."""

    def __init__(self, n, lam, m, x0, sigma, rank, world, seed):
        from oracle import pyoracle as po
        self.n, self.lam, self.m = n, lam, m
        self.pop_count = lam // world
        self.pop_offset = rank * self.pop_count
        self.ns = (n + 3) & ~3
        self.payload_floats = self.ns + 4
        self.rng_seed = seed
        self.gen = 0
        # the replicated state lives in an oracle instance that is only used as a state container + sampler
        self.po = po
        self.ref = po.OracleLMCMA(n, x0=x0, lam=lam, m=m, sigma=sigma, Z0=self._z(0))
        self.prev = None

    def _z(self, gen):   # counter-based: every rank can regenerate every row
        return np.random.default_rng([self.rng_seed, gen]).standard_normal((self.lam, self.n))

    def evaluate(self, f_local):
        X = self.ref.array("X")[self.pop_offset:self.pop_offset + self.pop_count]
        f = np.sum((1.0 + np.arange(self.n)) * X * X, axis=1).astype(np.float32)
        f_local.copy_(__import__("torch").from_numpy(f))

    def rank(self, f_all, payload):
        f = f_all.numpy().astype(np.float64)
        self.f_all = f.copy()
        order = np.argsort(f, kind="stable")
        rank = np.empty(self.lam, np.int64); rank[order] = np.arange(self.lam)
        w = self.ref.array("weights")
        X = self.ref.array("X"); xm = self.ref.array("xmean")
        acc = np.zeros(self.ns)
        for r in range(self.pop_offset, self.pop_offset + self.pop_count):
            if rank[r] < len(w):
                acc[:self.n] += w[rank[r]] * (X[r] - xm)
        S = 0
        if self.prev is not None:
            mine = f[self.pop_offset:self.pop_offset + self.pop_count]
            S = int(np.sum(self.prev[None, :] < mine[:, None]))
        out = np.zeros(self.payload_floats, np.float32)
        out[:self.ns] = acc
        out[self.ns:self.ns + 2] = np.array([S & 0xffffffff, S >> 32], np.uint32).view(np.float32)
        payload.copy_(__import__("torch").from_numpy(out))

    def update(self, payload_all, world):
        pay = payload_all.numpy().reshape(world, self.payload_floats)
        S = sum(int(p[self.ns:self.ns + 2].view(np.uint32)[0]) + (int(p[self.ns:self.ns + 2].view(np.uint32)[1]) << 32) for p in pay)
        self.S_total = S
        self.shift = pay[:, :self.n].astype(np.float64).sum(axis=0)
        self.gen += 1
        self.ref.tell_all(self.f_all, self._z(self.gen))      # the replicated update, identical on every rank
        self.prev = self.f_all.copy()


def _worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, lam, m = 12, 16, 6
    x0 = np.full(n, 0.4)
    be = NumpySplitBackend(n, lam, m, x0, 0.5, rank, world, seed=3)
    sp = parallel.SplitPopulation(be, dist, "cpu")
    checks = []
    for g in range(6):
        xm_before = be.ref.array("xmean").copy()
        sig_before = be.ref.doubles()["sigma"]
        sp.generation()
        # the gathered payloads must reproduce the replicated update: mean shift and merged-rank statistic
        shift_ref = be.ref.array("xmean") - xm_before
        checks.append(float(np.max(np.abs(be.shift - shift_ref))))
        if g > 0:
            L = lam
            sum_cur = L * (L - 1) // 2 + be.S_total
            sum_prev = L * (2 * L - 1) - sum_cur
            success = (sum_prev / L - sum_cur / L) / L
            s_new = be.ref.doubles()["s"]
            checks.append(abs(be.ref.doubles()["sigma"] - sig_before * np.exp(s_new)))
            checks.append(abs(s_new - ((1 - 0.3) * s_prev + 0.3 * (success - 0.25))))
        s_prev = be.ref.doubles()["s"]
    gathered = [None] * world
    dist.all_gather_object(gathered, (be.ref.array("xmean").tolist(), be.ref.doubles()["sigma"], sp.f_all.tolist()))
    if rank == 0:
        out_q.put((max(checks), gathered))
    dist.destroy_process_group()


def test_split_population_protocol_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    worst, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst < 1e-6, worst                               # FP32 payload vs FP64 replicated update
    assert gathered[0] == gathered[1]                        # replicas identical, fitness gathered in rank order
