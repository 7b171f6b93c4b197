"""Embedding-table residency layer: a device-resident bf16 [N, ld] item/passage table.

Replaces the reference's pageable host fp32 table (scripts/ms_marco_eval.py:123-152
``generate_embeddings`` -> ``.to("cpu")`` -> ``vstack``; src/ccrec/models/bbpr.py:466-483
``get_all_embeddings``): encoder batches are cast (and, for CCREC_SIM_TYPE=cos, L2-normalised
in fp32) straight into the device table, which then never crosses PCIe again.
"""
from __future__ import annotations

import torch

from . import engine


def _round_up(x, m):
    return (x + m - 1) // m * m


class EmbeddingTable:
    """Row-major bf16 table on one GPU.  ``id_offset`` is the global id of local row 0
    (row-sharded tables: scripts-level corpus position = id_offset + local row)."""

    def __init__(self, capacity, dim=768, device="cuda", normalize=False, id_offset=0):
        self.dim = int(dim)
        self.ld = _round_up(self.dim, 8)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("EmbeddingTable lives on a CUDA device (ccr_b200 has no CPU path)")
        self.normalize = bool(normalize)
        self.id_offset = int(id_offset)
        self.data = torch.empty((int(capacity), self.ld), dtype=torch.bfloat16, device=self.device)
        self.n = 0

    def __len__(self):
        return self.n

    @property
    def rows(self):
        return self.data[: self.n]

    def nbytes(self):
        return self.n * self.ld * 2

    def append(self, emb):
        """Append a batch of embeddings [m, dim] (any device / float dtype)."""
        emb = torch.as_tensor(emb)
        if emb.dim() == 1:
            emb = emb.unsqueeze(0)
        m = emb.shape[0]
        if emb.shape[1] != self.dim:
            raise ValueError(f"embedding dim {emb.shape[1]} != table dim {self.dim}")
        if self.n + m > self.data.shape[0]:
            grown = torch.empty((max(self.n + m, 2 * self.data.shape[0]), self.ld), dtype=torch.bfloat16,
                                device=self.device)
            grown[: self.n].copy_(self.data[: self.n])
            self.data = grown
        if emb.dtype not in (torch.float32, torch.bfloat16):
            emb = emb.float()
        emb = emb.to(self.device, non_blocking=True)
        if emb.stride(1) != 1:
            emb = emb.contiguous()
        engine.ingest_rows(emb, self.data[self.n : self.n + m], normalize=self.normalize)
        self.n += m
        return self

    @classmethod
    def from_tensor(cls, emb, device="cuda", normalize=False, id_offset=0, chunk=1 << 18):
        emb = torch.as_tensor(emb)
        t = cls(emb.shape[0], emb.shape[1], device=device, normalize=normalize, id_offset=id_offset)
        for s in range(0, emb.shape[0], chunk):
            t.append(emb[s : s + chunk])
        return t

    def encode_queries(self, q):
        """Query/user embeddings -> bf16 [B, ld] under the table's similarity convention."""
        q = torch.as_tensor(q)
        if q.dim() == 1:
            q = q.unsqueeze(0)
        if q.dtype not in (torch.float32, torch.bfloat16):
            q = q.float()
        q = q.to(self.device, non_blocking=True)
        if q.stride(1) != 1:
            q = q.contiguous()
        out = torch.empty((q.shape[0], self.ld), dtype=torch.bfloat16, device=self.device)
        return engine.ingest_rows(q, out, normalize=self.normalize)

    def search(self, queries, k, mask=None, algo=0, allow_short=False, want_f64=False, encoded=False,
               want_keys=False):
        """Top-k of queries against the resident rows.  Returns (scores, global ids[, scores64 | keys])."""
        q = queries if encoded else self.encode_queries(queries)
        return engine.score_topk(q, self.data, k, mask=mask, id_offset=self.id_offset, algo=algo,
                                 allow_short=allow_short, want_f64=want_f64, n_items=self.n, D=self.ld,
                                 want_keys=want_keys)

    def dense_scores(self, queries, encoded=False):
        q = queries if encoded else self.encode_queries(queries)
        return engine.score_dense(q, self.data, n_items=self.n, D=self.ld)

    # ---- persistence: the `name=` analogue of generate_embeddings (ms_marco_eval.py:150-151) ----
    def save(self, path):
        torch.save({"rows": self.rows.cpu(), "dim": self.dim, "normalize": self.normalize,
                    "id_offset": self.id_offset}, path)

    @classmethod
    def load(cls, path, device="cuda"):
        blob = torch.load(path)
        t = cls(blob["rows"].shape[0], blob["dim"], device=device, normalize=blob["normalize"],
                id_offset=blob["id_offset"])
        t.data[: blob["rows"].shape[0]].copy_(blob["rows"])
        t.n = blob["rows"].shape[0]
        return t
