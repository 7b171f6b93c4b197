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


def _worker_ranking(rank, world, port, ret):
    """ranking_sharded: every rank encodes only its corpus slice; result equals the single-table
    reference restatement (ids exactly, scores to fp32 round-off)."""
    for p in (ROOT, os.path.join(ROOT, "crowd-coachable-recommendations_b200"), os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CCREC_SIM_TYPE="dot")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    import importlib

    from ccr_b200 import dist as cdist, engine
    from oracle import ccr_oracle as O

    cranking = importlib.import_module("ccr_b200.ranking")  # the package attribute `ranking` is the function

    class HostMask(engine.SparseMask):
        def __init__(self, indptr, cols, vals, n_cols, mode, device=None):
            self.n_rows, self.n_cols, self.mode = len(indptr) - 1, int(n_cols), mode
            self.nnz = int(indptr[-1])
            self.host = (np.asarray(indptr, np.int64), np.asarray(cols, np.int32), np.asarray(vals, np.float64))
            self.max_row_nnz = int(np.diff(self.host[0]).max()) if self.n_rows else 0
            self.device = None

    engine.SparseMask = HostMask

    class CpuIndex(cdist.ShardedIndex):
        packed_calls = 0

        def _make_table(self, capacity, dim, normalize):
            self.rows_seen = []
            return None

        def add_local(self, emb):
            self.rows_seen.append(torch.as_tensor(emb))
            return self

        def _encode(self, queries):
            return queries

        def _local_topk(self, q, kk, mask):
            P_local = torch.cat(self.rows_seen)
            assert P_local.shape[0] == self.hi - self.lo  # exactly this rank's slice was encoded
            kl = min(kk, P_local.shape[0])
            s, i = O.score_topk_ref(q, P_local, kl, mask=mask.host if mask else None,
                                    mode=mask.mode if mask else O.MASK_NONE, id_offset=self.lo, return_f64=True,
                                    round_bf16=False)
            pad_s = torch.full((q.shape[0], kk - kl), float("-inf"), dtype=torch.float64)
            pad_i = torch.full((q.shape[0], kk - kl), -1, dtype=torch.int64)
            return torch.cat([s.double(), pad_s], 1), torch.cat([i, pad_i], 1)

        # packed exchange (block masks are SET -1e6: float32-exact): keys exactly as the ABI defines them
        def _local_topk_keys(self, q, kk, mask):
            d, i = self._local_topk(q, kk, mask)
            f = d.float().numpy().view(np.uint32).astype(np.uint64)
            o = np.where(f >> 31, f ^ 0xFFFFFFFF, f | 0x80000000)
            keys = (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - i.numpy().astype(np.uint64))
            keys = np.where(i.numpy() >= 0, keys, np.uint64(0))
            CpuIndex.packed_calls += 1
            return torch.as_tensor(keys.view(np.int64))

        def _merge_keys(self, keys, kk, packed=False):
            G, Bq, kin = keys.shape
            flat = np.ascontiguousarray(keys.permute(1, 0, 2).reshape(Bq, G * kin).numpy()).view(np.uint64)
            top = np.ascontiguousarray(np.sort(flat, axis=1)[:, ::-1][:, :kk])  # descending
            if packed:
                return torch.as_tensor(top.view(np.int64))
            return self._unpack_keys(torch.as_tensor(top.view(np.int64)))

        def _unpack_keys(self, keys):
            top = np.ascontiguousarray(keys.numpy()).view(np.uint64)
            o = (top >> np.uint64(32)).astype(np.uint32)
            f = np.where(o >> 31, o & 0x7FFFFFFF, ~o).astype(np.uint32).view(np.float32)
            ids = (np.uint64(0xFFFFFFFF) - (top & np.uint64(0xFFFFFFFF))).astype(np.int64)
            ids = np.where(top == 0, -1, ids)
            f = np.where(top == 0, -np.inf, f).astype(np.float32)
            return torch.as_tensor(f), torch.as_tensor(ids)

        def _merge(self, scores64, ids, kk):
            G, Bq, kin = scores64.shape
            flat_s = scores64.permute(1, 0, 2).reshape(Bq, G * kin)
            flat_i = ids.permute(1, 0, 2).reshape(Bq, G * kin)
            key = torch.where(flat_i >= 0, flat_s, torch.full_like(flat_s, float("-inf")))
            # order: score descending, id ascending (ids ascend with the rank-major layout only per run)
            order = np.lexsort((flat_i.numpy(), -key.numpy()), axis=1)[:, :kk]
            order = torch.as_tensor(order)
            out_s, out_i = torch.gather(key, 1, order), torch.gather(flat_i, 1, order)
            return out_s.float(), out_i, out_s

    c = cases.ranking_case("dot_block_tail_n1100")
    table = cases.TextTable(c["table"])
    prof = cranking.ranking_sharded(c["corpus"], c["queries"], table, c["batch_size"], c["block_dict"], device="cpu",
                                    index_cls=CpuIndex)
    want = O.ranking_ref(c["corpus"], c["queries"], cases.TextTable(c["table"]), c["batch_size"], c["block_dict"],
                         sim_type="dot")
    ok = list(prof.keys()) == list(want.keys())
    for qid in want:
        live = [p for p, s in want[qid].items() if s > -1e6]
        ok &= list(prof[qid].keys())[: len(live)] == live            # same order over the unblocked part
        ok &= set(prof[qid].keys()) == set(want[qid].keys()) or len(want[qid]) == 1001
        ok &= bool(np.allclose(list(prof[qid].values()), list(want[qid].values()), rtol=1e-5, atol=1e-4))
    # the encoder saw every query batch plus only this rank's share of the corpus
    n = len(c["corpus"])
    lo, hi = cdist.shard_bounds(n, world, rank)
    ok &= CpuIndex.packed_calls > 0  # the SET -1e6 block mask takes the one-gather packed exchange
    ret[rank] = (bool(ok), table.calls, -(-len(c["queries"]) // c["batch_size"]) + -(-(hi - lo) // c["batch_size"]))
    dist.barrier()
    dist.destroy_process_group()


def test_ranking_sharded_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29850 + os.getpid() % 100
    mp.spawn(_worker_ranking, args=(world, port, ret), nprocs=world, join=True)
    out = dict(ret)
    assert out[0][0] and out[1][0], out
    assert out[0][1] == out[0][2] and out[1][1] == out[1][2], out  # encoder calls: queries + own slice only
