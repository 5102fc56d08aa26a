"""world_size-2 gloo test of the multi-GPU commit orchestration (stwo-brainfuck_b200/sharded.py): the same code path as
on GPUs, with the CPU oracle standing in for the kernels.  Root must equal the unsharded tree's root."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROOT_LOG = 14


class OracleShardOps:
    def __init__(self, orc, torch):
        self.orc, self.torch = orc, torch

    def lde(self, host_cols, log_blowup):
        out = []
        for h in host_cols:
            c = self.orc.interpolate(h, ROOT_LOG)
            out.append(self.torch.from_numpy(self.orc.evaluate(c, log_blowup, ROOT_LOG).view(np.int32)))
        return out

    def empty(self, n):
        return self.torch.empty(n, dtype=self.torch.int32)

    def commit_on_layer(self, log_size, prev, cols):
        tn = lambda t: np.ascontiguousarray(t.numpy().view(np.uint32))
        out = self.orc.commit_on_layer(log_size, tn(prev) if prev is not None else None, [tn(c) for c in cols])
        return self.torch.from_numpy(out.view(np.int32))

    def to_numpy(self, t):
        return t.numpy().view(np.uint32)


def _worker(rank, world, port, cases, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module("stwo-brainfuck_b200.sharded")
    orc = Oracle()
    results = []
    for logs in cases:
        owner = sharded.assign_columns(logs, world)
        cols = {i: np.random.default_rng(100 + i).integers(0, P, size=1 << lg, dtype=np.uint32) for i, lg in enumerate(logs)}
        owned = {i: cols[i] for i in range(len(logs)) if owner[i] == rank}
        root = sharded.sharded_commit(OracleShardOps(orc, torch), dist, logs, owned, 1)
        if rank == 0:
            ldes = [orc.evaluate(orc.interpolate(cols[i], ROOT_LOG), 1, ROOT_LOG) for i in range(len(logs))]
            results.append((root.tolist(), orc.merkle_commit(ldes)[0].tolist()))
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cases", [(2, [[8, 6, 8, 4, 6, 6, 3, 8, 5], [3, 3, 4], [5]]), (4, [[7, 9, 9, 5, 4, 7, 7, 3]])])
def test_sharded_commit_root_matches_unsharded(world, cases):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(results) == len(cases)
    for got, want in results:
        assert got == want


def test_assign_columns_balances_by_size():
    sharded = importlib.import_module("stwo-brainfuck_b200.sharded")
    logs = [24] * 8 + [22] * 17 + [20] * 33 + [19] * 24 + [11] * 17 + [4] * 29   # fib19 main-trace tree
    owner = sharded.assign_columns(logs, 8)
    load = [sum(1 << logs[i] for i in range(len(logs)) if owner[i] == r) for r in range(8)]
    assert max(load) <= 1.3 * min(load)
