"""world_size-2 gloo test (CPU) of the row-sharded exchange path: shard bounds, mask column
sharding, id offsets, all-gather layout and the merge contract.  The two device steps of
ShardedIndex (local fused top-k, G-way merge kernel) are replaced by the oracle here -- this
exercises the host plumbing only; the kernels are covered by tests/dist_check.py on GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "crowd-coachable-recommendations_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ccr_b200 import dist as cdist, engine
    from oracle import ccr_oracle as O

    N, D, B, k = 1003, 32, 7, 20
    g = torch.Generator().manual_seed(5)
    P = torch.randn((N, D), generator=g)
    Q = torch.randn((B, D), generator=g)
    rs = np.random.RandomState(1)
    rows = [np.unique(rs.randint(0, N, size=rs.randint(0, 30))) for _ in range(B)]
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    cols = np.concatenate(rows)
    vals = np.where(rs.rand(len(cols)) < 0.5, -1e10, 1e5)

    class HostMask(engine.SparseMask):  # host arrays only (no device on this box)
        def __init__(self, indptr, cols, vals, n_cols, mode, device=None):
            self.n_rows, self.n_cols, self.mode = len(indptr) - 1, int(n_cols), mode
            self.nnz = int(indptr[-1])
            self.host = (np.asarray(indptr, np.int64), np.asarray(cols, np.int32), np.asarray(vals, np.float64))
            self.max_row_nnz = int(np.diff(self.host[0]).max()) if self.n_rows else 0
            self.device = None

    engine.SparseMask = HostMask

    class CpuIndex(cdist.ShardedIndex):
        def _make_table(self, capacity, dim, normalize):
            return None

        def _encode(self, queries):
            return queries

        def _local_topk(self, q, kk, mask):
            n_local = self.hi - self.lo
            kl = min(kk, n_local)
            s, i = O.score_topk_ref(q, P[self.lo:self.hi], kl, mask=mask.host if mask else None, mode=O.MASK_ADD,
                                    id_offset=self.lo, return_f64=True)
            pad_s = torch.full((q.shape[0], kk - kl), float("-inf"), dtype=torch.float64)
            pad_i = torch.full((q.shape[0], kk - kl), -1, dtype=torch.int64)
            return torch.cat([s.double(), pad_s], 1), torch.cat([i, pad_i], 1)

        def _merge(self, scores64, ids, kk):
            G, Bq, _ = scores64.shape
            out_s = torch.empty(Bq, kk, dtype=torch.float64)
            out_i = torch.empty(Bq, kk, dtype=torch.int64)
            for b in range(Bq):
                ent = [(-float(scores64[g_, b, j]), int(ids[g_, b, j])) for g_ in range(G)
                       for j in range(scores64.shape[2]) if ids[g_, b, j] >= 0]
                ent.sort()
                out_s[b] = torch.tensor([-e[0] for e in ent[:kk]])
                out_i[b] = torch.tensor([e[1] for e in ent[:kk]])
            return out_s.float(), out_i, out_s

    idx = CpuIndex(N, D, device="cpu")
    assert (idx.lo, idx.hi) == cdist.shard_bounds(N, world, rank)
    mask = HostMask(indptr, cols, vals, N, engine.MASK_ADD)
    s, i, d = idx.search(Q, k, mask=mask)
    ref_s, ref_i = O.score_topk_ref(Q, P, k, mask=(indptr, cols, vals), mode=O.MASK_ADD, return_f64=True)
    ok = bool((i == ref_i).all()) and bool(torch.allclose(d, ref_s.double()))
    with pytest.raises(RuntimeError):
        idx.search(Q, N + 1)
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_plumbing_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
