"""Lazy score-matrix algebra with the same public surface as the reference's
src/rime_lite/util/score_array.py (``LazyScoreBase`` ops ``@ + - * / exp softplus sigmoid apply
T [] collate_fn as_tensor batch_size``, ``auto_cast_lazy_score``, ``score_op``), so that
``S = LazyDenseMatrix(U) @ LazyDenseMatrix(V).T + prior_csr`` written against the reference keeps
working.  What is new is ``fused_plan``: it recognises the hot-path shape

    MatMul(LazyDense, LazyDense)  [ +/- LazySparse ... ]

and lets ``ccr_b200.util._assign_topk`` hand it to the fused CUDA kernel instead of
materialising [batch, N] float64 blocks (score_array.py:291-293, :173-174 in the reference).

Out of scope here: the VAE / RandScore leaves and LazyScoreModel.
"""
from __future__ import annotations

import operator
import os

import numpy as np
import scipy.sparse as sps
import torch


def get_batch_size(shape):
    """Rows per dense batch, the reference's memory heuristic (score_array.py:8-18).  The fused
    path ignores it (nothing dense is materialised) but ``as_tensor`` consumers still use it."""
    n_users, n_items = shape
    frac = float(os.environ.get("BATCH_SIZE_FRAC", 0.1))
    if torch.cuda.device_count():
        total_memory = torch.cuda.get_device_properties(0).total_memory
    else:
        total_memory = 16e9
    max_batch_size = total_memory / 8 / max(1, n_items) * frac
    n_batches = int(n_users / max_batch_size) + 1
    return int(np.ceil(n_users / n_batches))


def sps_to_torch(x, device):
    coo = x.tocoo()
    idx = np.vstack((coo.row, coo.col))
    return torch.sparse_coo_tensor(idx, coo.data, coo.shape, device=device)


def auto_device():
    return "cuda" if torch.cuda.is_available() else "cpu"


def auto_tensor(x, device=None):
    if device is None:
        device = auto_device()
    if hasattr(x, "as_tensor"):
        return x.as_tensor(device)
    if sps.issparse(x):
        return sps_to_torch(x, device).to_dense()
    return torch.as_tensor(x, device=device)


def _row_index(key, n_rows):
    """slice / scalar / array -> int array, wrapped modulo n_rows (broadcast rows of size 1)."""
    if isinstance(key, slice):
        if key.stop is None:
            raise ValueError("row slices need an explicit stop")
        key = range(key.stop)[key]
    return np.array(key, ndmin=1) % n_rows


class LazyScoreBase:
    """Deferred score matrix: build with operators, slice rows, evaluate with ``as_tensor``."""

    def __init__(self, shape):
        self.shape = tuple(shape)

    def __repr__(self):
        return f"<{type(self).__name__} {self.shape}>"

    def __len__(self):
        return self.shape[0]

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def batch_size(self):
        return get_batch_size(self.shape)

    def numpy(self):
        return self.as_tensor().numpy()

    def as_tensor(self, device=None):
        raise NotImplementedError

    @property
    def T(self):
        raise NotImplementedError

    def __getitem__(self, key):
        raise NotImplementedError

    @staticmethod
    def collate_fn(parts):
        raise NotImplementedError

    # ---- expression builders ----
    def __matmul__(self, other):
        return MatMulExpression(operator.matmul, [self, other])

    def __add__(self, other):
        return ElementWiseExpression(operator.add, [self, other])

    def __sub__(self, other):
        return ElementWiseExpression(operator.sub, [self, other])

    def __mul__(self, other):
        return ElementWiseExpression(operator.mul, [self, other])

    def __truediv__(self, other):
        return ElementWiseExpression(operator.truediv, [self, other])

    def exp(self):
        return ElementWiseExpression(torch.exp, [self])

    def softplus(self):
        return ElementWiseExpression(torch.nn.functional.softplus, [self])

    def sigmoid(self):
        return ElementWiseExpression(torch.sigmoid, [self])

    def apply(self, op):
        return ElementWiseExpression(op, [self])


def auto_cast_lazy_score(other):
    if other is None:
        return None
    if isinstance(other, LazyScoreBase):
        return other
    if sps.issparse(other):
        return LazySparseMatrix(other)
    if hasattr(other, "values") and hasattr(other, "columns"):  # DataFrame
        return LazyDenseMatrix(other.values)
    if torch.is_tensor(other):
        return LazyDenseMatrix(other.detach().cpu().numpy())
    return LazyDenseMatrix(other)


class LazySparseMatrix(LazyScoreBase):
    def __init__(self, c):
        self.c = c.tocsr()
        self.shape = tuple(c.shape)

    def numpy(self):
        return self.c.toarray()

    def as_tensor(self, device=None):
        return sps_to_torch(self.c, device).to_dense()

    @property
    def T(self):
        return LazySparseMatrix(self.c.T)

    def __getitem__(self, key):
        if np.isscalar(key):
            lo, hi = self.c.indptr[key], self.c.indptr[key + 1]
            return _SparseRow(self.c.data[lo:hi], self.c.indices[lo:hi], self.c.shape[1])
        return LazySparseMatrix(self.c[key])

    @staticmethod
    def collate_fn(parts):
        return LazySparseMatrix(sps.vstack([p.c for p in parts]))


class _SparseRow(LazyScoreBase):
    """One CSR row picked by a scalar index (DataLoader-style access)."""

    def __init__(self, values, keys, n_cols):
        self.values, self.keys, self.n_cols = values, keys, n_cols
        self.shape = (1, n_cols)

    @staticmethod
    def collate_fn(parts):
        lens = [len(p.keys) for p in parts]
        csr = sps.csr_matrix(
            (np.hstack([p.values for p in parts]), np.hstack([p.keys for p in parts]),
             np.concatenate([[0], np.cumsum(lens)])),
            shape=(len(parts), parts[0].n_cols))
        return LazySparseMatrix(csr)


class LazyDenseMatrix(LazyScoreBase):
    """Scalars and arrays as 2-d arrays; row indices wrap so 1-row operands broadcast."""

    def __init__(self, c):
        self.c = np.array(c, ndmin=2)
        self.shape = self.c.shape

    def numpy(self):
        return self.c

    def as_tensor(self, device=None):
        return torch.as_tensor(self.c, device=device)

    @property
    def T(self):
        return LazyDenseMatrix(self.c.T)

    def __getitem__(self, key):
        return LazyDenseMatrix(self.c[_row_index(key, self.shape[0])])

    @staticmethod
    def collate_fn(parts):
        return LazyDenseMatrix(np.vstack([p.c for p in parts]))


def _op_name(op):
    return getattr(op, "__name__", repr(op))


class _Expression:
    def __init__(self, op, children):
        self.op = op
        self.children = [auto_cast_lazy_score(c) for c in children]
        self._setup()

    def _setup(self):
        pass

    def traverse(self, op_func=_op_name):
        out = ""
        for i, c in enumerate(self.children):
            out += f"({c.traverse(op_func)})" if hasattr(c, "traverse") else f"{c}"
            if i == 0:
                out += f" {op_func(self.op)} "
        return out

    def __repr__(self):
        return self.traverse()

    def as_tensor(self, device=None):
        return self.op(*[c.as_tensor(device) for c in self.children])


class ElementWiseExpression(_Expression, LazyScoreBase):
    """Element-wise op over children, broadcasting 1-row / 1-column operands."""

    def _setup(self):
        dims = np.array([c.shape for c in self.children])
        self.shape = (int(dims[:, 0].max()), int(dims[:, 1].max()))

    @property
    def T(self):
        return ElementWiseExpression(self.op, [c.T for c in self.children])

    def __getitem__(self, key):
        return ElementWiseExpression(self.op, [c[key] for c in self.children])

    @staticmethod
    def collate_fn(batch):
        first = batch[0]
        cols = zip(*[b.children for b in batch])
        return ElementWiseExpression(first.op, [c.collate_fn(list(parts)) for c, parts in zip(first.children, cols)])


class MatMulExpression(_Expression, LazyScoreBase):
    def _setup(self):
        self.left, self.right = self.children
        assert self.left.shape[1] == self.right.shape[0], (
            f"matmul shape check fail: {self.left.shape} vs {self.right.shape}")
        self.shape = (self.left.shape[0], self.right.shape[1])

    @property
    def T(self):
        return MatMulExpression(self.op, [self.right.T, self.left.T])

    def __getitem__(self, key):
        out = MatMulExpression(self.op, [self.left[key], self.right])
        return out

    @staticmethod
    def collate_fn(batch):
        left = type(batch[0].left).collate_fn([b.left for b in batch])
        return MatMulExpression(batch[0].op, [left, batch[0].right])

    def as_tensor(self, device=None):
        """score_array.py:291-293 of the reference evaluates ``left.as_tensor(device) @
        right.as_tensor(device)`` with torch.  On a CUDA device a dense x dense product runs through
        ``ccr_score_dense_f32`` instead (bf16 operands on the cached device table, fp32 accumulate; the
        TMA + tcgen05 pipeline for tiles of >= 2^20 scores); every other case keeps the generic op."""
        dev = torch.device(device) if device is not None else None
        if (dev is not None and dev.type == "cuda" and self.op is operator.matmul
                and isinstance(self.left, LazyDenseMatrix) and isinstance(self.right, LazyDenseMatrix)):
            from .util import _device_table_for

            table = _device_table_for(self.right)
            return table.dense_scores(torch.as_tensor(np.ascontiguousarray(self.left.c)))
        return super().as_tensor(device)


def batch_op_iter(S, op, device=None):
    if isinstance(op, str):
        op = getattr(torch, op)
    step = S.batch_size
    for i in range(0, len(S), step):
        yield op(S[i : min(len(S), i + step)].as_tensor(device))


def score_op(S, op, device=None, reduce_fn=None):
    """max / min / sum over the whole lazy matrix, streamed by row batches."""
    import functools

    if reduce_fn is None:
        reduce_fn = {"max": max, "min": min, "sum": operator.add}[op]
    return functools.reduce(reduce_fn, batch_op_iter(S, op, device))


# ----------------------------------------------------------------------------------------
# hot-path recognition
# ----------------------------------------------------------------------------------------
class FusedPlan:
    """left [B,D] and right [D,N] dense factors plus the summed sparse additive term (or None)."""

    def __init__(self, left, right, sparse_terms, shape):
        self.left, self.right, self.shape = left, right, shape
        self.sparse = None
        for sign, csr in sparse_terms:
            term = csr if sign > 0 else -csr
            self.sparse = term if self.sparse is None else self.sparse + term
        if self.sparse is not None:
            self.sparse = sps.csr_matrix(self.sparse, dtype=np.float64)


def fused_plan(S):
    """Return a FusedPlan when ``S`` is MatMul(LazyDense, LazyDense) optionally plus / minus
    LazySparse terms of the full shape; ``None`` for every other expression."""
    matmuls, sparse_terms = [], []

    def walk(node, sign):
        if isinstance(node, MatMulExpression):
            if isinstance(node.left, LazyDenseMatrix) and isinstance(node.right, LazyDenseMatrix) and sign > 0:
                matmuls.append(node)
                return True
            return False
        if isinstance(node, LazySparseMatrix):
            sparse_terms.append((sign, node.c))
            return True
        if isinstance(node, ElementWiseExpression) and len(node.children) == 2:
            if node.op is operator.add:
                return walk(node.children[0], sign) and walk(node.children[1], sign)
            if node.op is operator.sub:
                return walk(node.children[0], sign) and walk(node.children[1], -sign)
        return False

    if not walk(S, +1) or len(matmuls) != 1:
        return None
    mm = matmuls[0]
    shape = tuple(S.shape)
    if tuple(mm.shape) != shape or any(tuple(c.shape) != shape for _, c in sparse_terms):
        return None  # broadcasting terms are left to the generic path
    return FusedPlan(mm.left, mm.right, sparse_terms, shape)


class DensePlan:
    """An already materialised dense score matrix plus the summed sparse additive term (or None):
    what the reference's unmodified ``BertBPR.transform(D) + D.prior_score`` evaluates to
    (src/ccrec/models/bbpr.py:550, src/ccrec/models/bert_mt.py:376)."""

    def __init__(self, dense, sparse_terms, shape):
        self.dense, self.shape = dense, shape
        self.sparse = None
        for sign, csr in sparse_terms:
            term = csr if sign > 0 else -csr
            self.sparse = term if self.sparse is None else self.sparse + term
        if self.sparse is not None:
            self.sparse = sps.csr_matrix(self.sparse, dtype=np.float64)


def dense_plan(S):
    """Return a DensePlan when ``S`` is one LazyDenseMatrix optionally plus / minus LazySparse
    terms, all of the full shape; ``None`` for every other expression."""
    leaves, sparse_terms = [], []

    def walk(node, sign):
        if isinstance(node, LazyDenseMatrix):
            if sign < 0:
                return False
            leaves.append(node)
            return True
        if isinstance(node, LazySparseMatrix):
            sparse_terms.append((sign, node.c))
            return True
        if isinstance(node, ElementWiseExpression) and len(node.children) == 2:
            if node.op is operator.add:
                return walk(node.children[0], sign) and walk(node.children[1], sign)
            if node.op is operator.sub:
                return walk(node.children[0], sign) and walk(node.children[1], -sign)
        return False

    if not walk(S, +1) or len(leaves) != 1:
        return None
    shape = tuple(S.shape)
    if tuple(leaves[0].shape) != shape or any(tuple(c.shape) != shape for _, c in sparse_terms):
        return None
    return DensePlan(leaves[0], sparse_terms, shape)
