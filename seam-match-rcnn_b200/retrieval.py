"""Retrieval front-ends over the engine: single-GPU search, the gallery-sharded multi-GPU
search (one process per GPU, NCCL all-gather of the per-shard candidate lists), and the
evaluation-script outputs (rank of the true shop item, top-k hit counts).

The reference scores one query at a time on the CPU (evaluate_movingfashion.py:157-277,
evaluate_multiDF2.py:209-230); here all queries go through one fused pass.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .engine import PreparedGallery, SeamEngine

K_THRESHOLDS = (1, 5, 10, 20)       # evaluate_movingfashion.py:15


# ----------------------------------------------------------------------------------------
# partitioning (pure host logic, exercised on CPU with gloo in tests/)
# ----------------------------------------------------------------------------------------
def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of n rows over `world` ranks: [lo, hi) of `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather tensors that differ in their first dimension (padded to the maximum)."""
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    sizes = [int(v) for v in ns]
    m = max(sizes)
    pad = x
    if x.shape[0] < m:
        pad = torch.cat([x, x.new_zeros((m - x.shape[0],) + tuple(x.shape[1:]))], 0)
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], 0)


# ----------------------------------------------------------------------------------------
# single GPU
# ----------------------------------------------------------------------------------------
def search(engine: SeamEngine, seq: torch.Tensor, mask: Optional[torch.Tensor], gallery, k: int = 20):
    """aggregation -> scorer -> top-k for all tracks.  Returns (scores, margins, idx int32)."""
    if not isinstance(gallery, PreparedGallery):
        gallery = engine.prepare_gallery(gallery)
    q = engine.aggregate(seq, mask)
    return engine.score_topk(q, gallery, k)


# ----------------------------------------------------------------------------------------
# gallery-sharded multi-GPU search
# ----------------------------------------------------------------------------------------
class ShardedRetriever:
    """Each rank holds a contiguous slice of the gallery (prepared once) and scores every
    query against it; queries are aggregated in slices and all-gathered; the per-rank
    (Q,k) candidate lists are all-gathered and merged.  `ops` is the compute provider -- a
    SeamEngine in production; tests inject a CPU stand-in to exercise the collective logic
    with gloo.
    """

    def __init__(self, ops, gallery_shard: torch.Tensor, shard_offset: int, group=None):
        self.ops = ops
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.gallery = ops.prepare_gallery(gallery_shard, index_offset=shard_offset)

    @classmethod
    def from_full_gallery(cls, ops, gallery: torch.Tensor, group=None):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(gallery.shape[0], world, rank)
        return cls(ops, gallery[lo:hi], lo, group)

    def aggregate(self, seq: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        """Every rank aggregates Q/N tracks; the (Q/N,256) results are all-gathered."""
        Q = seq.shape[1]
        if self.world == 1:
            return self.ops.aggregate(seq, mask)
        lo, hi = shard_bounds(Q, self.world, self.rank)
        part = self.ops.aggregate(seq[:, lo:hi], None if mask is None else mask[lo:hi])
        return all_gather_rows(part, self.group)

    def search_descriptors(self, q: torch.Tensor, k: int):
        sc, mg, ix = self.ops.score_topk(q, self.gallery, k)
        if self.world == 1:
            return sc, mg, ix
        packs = []
        for t in (sc, mg, ix):
            outs = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(outs, t.contiguous(), group=self.group)
            packs.append(torch.stack(outs, 0))
        return self.ops.merge_topk(packs[0], packs[1], packs[2])

    def search(self, seq: torch.Tensor, mask: Optional[torch.Tensor], k: int = 20):
        return self.search_descriptors(self.aggregate(seq, mask), k)


# ----------------------------------------------------------------------------------------
# evaluation-script outputs
# ----------------------------------------------------------------------------------------
@dataclass
class RetrievalReport:
    """What evaluate_movingfashion.py:268-277, 360-367 derives from the aggregated-descriptor
    ranking: per-query rank of the true shop item and the top-k accuracies."""
    ranks: torch.Tensor                 # (Q,) int32, 0 = best
    hits: List[int]                     # k_accs_aggr_desc
    accuracies: List[float]             # hits / Q
    topk_scores: torch.Tensor
    topk_idx: torch.Tensor


def evaluate_aggregated(engine: SeamEngine, seq, mask, gallery: torch.Tensor, target: torch.Tensor,
                        k_thresholds: Sequence[int] = K_THRESHOLDS) -> RetrievalReport:
    """The "AGGR DESC" block of the eval loop for all products at once.

    Per product the reference aggregates the track (evaluate_movingfashion.py:253-262), scores
    it against every shop descriptor (:263-267), argsorts (:268), reads the rank of the true
    shop item (:269) and counts hits for k in k_thresholds (:270-277).
    """
    q = engine.aggregate(seq, mask)
    gal = engine.prepare_gallery(gallery)
    kmax = min(max(k_thresholds), 32)
    sc, mg, ix = engine.score_topk(q, gal, kmax)
    ranks, _ = engine.rank_of_target(q, gal.g, target)
    r = ranks.cpu()
    hits = [int((r < k).sum()) for k in k_thresholds]
    n = max(1, q.shape[0])
    return RetrievalReport(ranks=ranks, hits=hits, accuracies=[h / n for h in hits], topk_scores=sc, topk_idx=ix)
