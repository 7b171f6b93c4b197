"""Row-sharded item table across the GPUs of one box: one process per GPU, local fused
top-k with global ids, ONE all-gather of (float64 score, int64 id) pairs, on-device G-way
merge.  The reference has no counterpart (it scores on GPU 0 only, scripts/ms_marco_eval.py:205).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import engine
from .table import EmbeddingTable


def shard_bounds(n_items, world_size, rank):
    """Contiguous row shard [lo, hi) of rank: ceil(N/G) rows each, last shard shorter/empty."""
    per = (n_items + world_size - 1) // world_size
    lo = min(n_items, rank * per)
    hi = min(n_items, lo + per)
    return lo, hi


class ShardedIndex:
    """Each rank holds rows [lo, hi) of the global table in an EmbeddingTable with id_offset=lo."""

    def __init__(self, n_items, dim=768, normalize=False, device=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_items = int(n_items)
        self.lo, self.hi = shard_bounds(self.n_items, self.world, self.rank)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.table = self._make_table(self.hi - self.lo, dim, normalize)

    def _make_table(self, capacity, dim, normalize):
        return EmbeddingTable(capacity, dim, device=self.device, normalize=normalize, id_offset=self.lo)

    # ---- the two device steps; CPU gloo tests replace them to exercise the plumbing ----
    def _local_topk(self, q_encoded, k, mask):
        s, i, d = self.table.search(q_encoded, k, mask=mask, allow_short=True, want_f64=True, encoded=True)
        return d, i

    def _merge(self, scores64, ids, k):
        return engine.merge_topk(scores64, ids, k)

    def _encode(self, queries):
        return self.table.encode_queries(queries)

    def add_local(self, emb):
        """Append rows of this rank's shard (in global order)."""
        self.table.append(emb)
        return self

    def search(self, queries, k, mask: engine.SparseMask | None = None):
        """queries replicated on every rank; mask is the GLOBAL CSR (sharded here by column range).
        Returns (scores f32 [B,k], global ids [B,k], scores f64 [B,k]) on every rank."""
        if k > self.n_items:
            raise RuntimeError("selected index k out of range")
        q = self._encode(queries)
        local_mask = mask.column_shard(self.lo, self.hi) if mask is not None else None
        d, i = self._local_topk(q, k, local_mask)
        if self.world == 1:
            return self._merge(d.unsqueeze(0), i.unsqueeze(0), k)
        B = d.shape[0]
        gs = torch.empty((self.world * B, k), dtype=d.dtype, device=d.device)   # rank-major concatenation
        gi = torch.empty((self.world * B, k), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(gs, d.contiguous(), group=self.group)
        dist.all_gather_into_tensor(gi, i.contiguous(), group=self.group)
        return self._merge(gs.view(self.world, B, k), gi.view(self.world, B, k), k)
