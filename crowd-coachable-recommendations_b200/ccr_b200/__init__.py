"""ccr_b200 -- B200-native candidate score-and-rank for Crowd Coachable Recommendations.

Public surface (mirrors the reference's call sites for this path):

* ``ranking`` / ``generate_embeddings`` / ``cos_sim``      <- scripts/ms_marco_eval.py
* ``_assign_topk`` / ``assign_topk``                       <- src/rime_lite/util/__init__.py
* ``LazyDenseMatrix`` ... ``auto_cast_lazy_score``         <- src/rime_lite/util/score_array.py
* ``evaluate_item_rec`` / ``evaluate_assigned``            <- src/rime_lite/metrics/__init__.py
* ``al_rank.rank_step`` / ``build_requests`` ...           <- scripts/al_0_rank.py:107-218
* ``BM25`` / ``ranking_bm25``                             <- scripts/bm_25.py, scripts/ms_marco_eval.py:165-186
* ``EmbeddingTable`` / ``ShardedIndex``                    (residency + row-sharding layer)
* ``score_topk`` / ``merge_topk`` / ``merge_topk_keys`` / ``argsort_scores`` / ``first_hit_rank``
                                                           (thin wrappers of the C ABI)
"""
from ._lib import ALGO_AUTO, ALGO_SIMT, ALGO_TCGEN05, MASK_ADD, MASK_NONE, MASK_SET, LIB_PATH  # noqa: F401
from .engine import (SparseMask, argsort_scores, device_status, first_hit_rank, ingest_rows, merge_topk,  # noqa: F401
                     merge_topk_keys, score_dense, score_topk, synchronize, topk_dense, unpack_topk_keys)
from .table import EmbeddingTable  # noqa: F401
from .score_array import (  # noqa: F401
    ElementWiseExpression,
    LazyDenseMatrix,
    LazyScoreBase,
    LazySparseMatrix,
    MatMulExpression,
    auto_cast_lazy_score,
    dense_plan,
    fused_plan,
    get_batch_size,
    score_op,
)
from .util import _argsort, _assign_topk, argsort, assign_topk, topk_lazy, transform_scores  # noqa: F401
from .metrics import evaluate_assigned, evaluate_item_rec  # noqa: F401
from .ranking import (  # noqa: F401
    build_block_mask,
    cos_sim,
    generate_embeddings,
    generate_embeddings_device,
    RankingProfile,
    mrr_at_k,
    qrels_csr,
    ranking,
    ranking_sharded,
    ranking_tensors,
)
from .dist import ShardedIndex, shard_bounds  # noqa: F401
from .bm25 import BM25, ranking_bm25  # noqa: F401
from . import al_rank  # noqa: F401
