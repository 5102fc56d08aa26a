"""Multi-GPU LDE + commit: column-sharded interpolate/evaluate -> all-to-all -> row-sharded Blake2s Merkle commit.

This is the exchange step BASELINE.json's north_star describes ("trace and interaction columns are split by column for
interpolation and extension, then re-sharded by row range over NVLink (NCCL all-to-all) for Merkle leaf hashing ... the
top Merkle layers finish on one GPU"), applied to one commitment tree (TreeBuilder::commit, crates/brainfuck_prover/src/
brainfuck_air/mod.rs:583).  One process per GPU; `torch.distributed` supplies the collectives (NCCL on GPUs, gloo in the
CPU tests); every arithmetic step is a C-ABI call of libstwo_cuda.so.

  phase A  rank r interpolates and extends the columns it owns (columns are independent: no communication)
  exchange per LDE size: all_to_all_single so that rank r holds rows [r*R/N, (r+1)*R/N) of EVERY column of that size
           (bit-reversed row ranges: children (2i, 2i+1) of a node are always in the same range)
  phase B  rank r hashes its row range of every layer down to the layer with N nodes (commit_on_layer on the sub-range,
           which is the same node function), the N sub-roots are all-gathered (N x 32 bytes) and every rank finishes
           the top log2(N) layers — identical root on every rank, bit-identical to the single-GPU tree.

The orchestration is written against a tiny `ops` adapter so that tests/test_sharded_cpu.py can run the very same code
with world_size 2 over gloo, with the CPU oracle standing in for the kernels.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def assign_columns(logs: Sequence[int], world: int) -> List[int]:
    """owner[col]: columns sorted by size (descending, stable) are dealt round-robin — sizes are very uneven (SURVEY.md
    Table S), and FFT cost is ~N log N, so this balances phase A within one largest column."""
    order = sorted(range(len(logs)), key=lambda i: -logs[i])
    owner = [0] * len(logs)
    for k, i in enumerate(order):
        owner[i] = k % world
    return owner


class CudaShardOps:
    """ops adapter over CudaBackend + torch CUDA tensors."""

    def __init__(self, pkg, backend, twiddles, torch):
        self.pkg, self.be, self.tw, self.torch = pkg, backend, twiddles, torch

    def lde(self, host_cols: List, log_blowup: int):
        """host_cols: numpy arrays (uploaded here) or device-resident Columns (cloned: interpolate works in place)."""
        cols = [h.clone() if isinstance(h, self.pkg.Column) else self.be.column(h) for h in host_cols]
        self.be.interpolate_columns(cols, self.tw)
        ldes = self.be.evaluate_polynomials(cols, log_blowup, self.tw)
        return [self.as_tensor(c) for c in ldes]

    def as_tensor(self, col):
        class _Cai:
            pass
        o = _Cai()
        o.__cuda_array_interface__ = {"shape": (len(col),), "typestr": "<i4", "data": (col.device_ptr(), False), "version": 2}
        t = self.torch.as_tensor(o, device="cuda")
        t._sbf_owner = col  # the library owns the memory: keep the handle alive as long as the tensor
        return t

    def empty(self, n):
        return self.torch.empty(n, dtype=self.torch.int32, device="cuda")

    def commit_on_layer(self, log_size: int, prev, cols):
        wrap = lambda t: self.be.wrap(t.data_ptr(), t.numel(), keepalive=t)
        out = self.be.commit_on_layer(log_size, wrap(prev) if prev is not None else None, [wrap(c) for c in cols])
        return self.as_tensor(out)

    def to_numpy(self, t):
        return t.cpu().numpy().view(np.uint32)


def sharded_commit(ops, dist, logs: Sequence[int], owned: Dict[int, np.ndarray], log_blowup: int = 1, group=None):
    """Commits the tree over columns with trace log sizes `logs`; `owned[col]` are this rank's columns (assign_columns).
    Returns the 8-word root (same on every rank)."""
    torch = ops.torch
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    w = world.bit_length() - 1
    assert 1 << w == world, "world size must be a power of two"
    owner = assign_columns(logs, world)
    mine = [i for i in range(len(logs)) if owner[i] == rank]
    assert sorted(owned) == mine, "owned columns do not match the assignment"
    lde = dict(zip(mine, ops.lde([owned[i] for i in mine], log_blowup)))   # phase A

    # ---- exchange: per LDE size, rows -> ranks
    sizes = sorted({lg + log_blowup for lg in logs}, reverse=True)
    local_cols = {}  # L -> list of tensors (this rank's row range, or the full column when it is tiny), tree column order
    for L in sizes:
        group_cols = [i for i in range(len(logs)) if logs[i] + log_blowup == L]
        R = 1 << L
        full = R < 16 * world           # tiny columns are replicated instead of sliced
        seg = R if full else R // world
        n_from = [sum(1 for i in group_cols if owner[i] == s) for s in range(world)]
        own = [lde[i] for i in group_cols if owner[i] == rank]
        if own:
            stacked = torch.stack(own)                                   # (n_own, R)
            if full:
                send = stacked.reshape(1, -1).repeat(world, 1).reshape(-1)
            else:
                send = stacked.view(len(own), world, seg).permute(1, 0, 2).contiguous().reshape(-1)
        else:
            send = ops.empty(0)
        recv = ops.empty(sum(n_from) * seg)
        dist.all_to_all_single(recv, send, output_split_sizes=[n * seg for n in n_from],
                               input_split_sizes=[len(own) * seg] * world, group=group)
        # recv = for each source s: (n_from[s], seg); put back into tree column order
        blocks, off = {}, 0
        for s in range(world):
            src_cols = [i for i in group_cols if owner[i] == s]
            for k, i in enumerate(src_cols):
                blocks[i] = recv[off + k * seg: off + (k + 1) * seg]
            off += n_from[s] * seg
        local_cols[L] = (full, [blocks[i] for i in group_cols])
    del lde

    # ---- phase B: row-sharded layers down to the layer with `world` nodes
    max_L = sizes[0]
    prev = None
    L = max_L
    while L >= w:
        full, cols = local_cols.get(L, (False, []))
        if full:  # replicated tiny columns: take this rank's rows
            seg = (1 << L) // world
            cols = [c[rank * seg:(rank + 1) * seg].contiguous() for c in cols]
        prev = ops.commit_on_layer(L - w, prev, cols)
        L -= 1
    if max_L < w:
        prev_full = None
        L = max_L
    else:
        gathered = ops.empty(8 * world)
        dist.all_gather_into_tensor(gathered, prev.contiguous(), group=group)   # N sub-roots, 32 bytes each
        prev_full = gathered
        L = w - 1
    # ---- top log2(world) layers on every rank (columns this small are always replicated)
    while L >= 0:
        full, cols = local_cols.get(L, (True, []))
        assert full or not cols
        prev_full = ops.commit_on_layer(L, prev_full, cols)
        L -= 1
    return ops.to_numpy(prev_full)[:8].copy()
