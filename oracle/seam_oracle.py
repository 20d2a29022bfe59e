"""CPU oracle for the SEAM Match-RCNN retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's hot path:
temporal aggregation (non-local block + frame-attention pooling), the
``(q-g)**2 -> Linear(256,2) -> softmax`` pair scorer, and per-query ranking.
It exists to CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it; the product package (``seam-match-rcnn_b200/``) never does, and has no CPU
fallback.

Where the arithmetic really lives: the reference is pure Python on top of
third-party ``torch`` (conv1d / conv2d / matmul / linear / softmax, fp32) and
``numpy`` (fp16 ufuncs, ``argsort``); neither is vendored nor pinned by the
reference (its README says "Pytorch 1.5.1 or more recent").  The module path is
therefore restated with the same torch CPU primitives in the same order, and the
evaluation-script path with the same numpy expressions, so that the oracle is
bit-comparable with the reference on one machine.

Parity pin: the reference ships no tests and no golden vectors ("parity unpinned"
by the reference itself).  The pin used here is outputs of the reference's own
modules, imported unmodified from /root/reference in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against them.

Each function cites the reference lines it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

D_MODEL = 256      # models/match_head.py:81  (d_model)
D_INTER = 128      # models/nlb.py:15-17      (in_channels // 2)

# state_dict keys of TemporalAggregationNLB that the hot path reads
# (probed from the reference module).
HOT_KEYS = (
    "newnlb.g.weight", "newnlb.g.bias",
    "newnlb.W.weight", "newnlb.W.bias",
    "newnlb.theta.weight", "newnlb.theta.bias",
    "newnlb.phi.weight", "newnlb.phi.bias",
    "newnlb.concat_project.0.weight",
    "attention_scorer.weight", "attention_scorer.bias",
    "last.weight", "last.bias",
)

Weights = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def random_weights(seed: int = 0, randomize_W: bool = True) -> Weights:
    """Random hot-path weights with the reference's default-init distributions.

    torch's Conv/Linear default init is U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and
    bias.  ``newnlb.W`` is zero-initialised by the reference (models/nlb.py:45-49) which
    makes the block an identity; ``randomize_W`` re-draws it U(+-1/sqrt(128)) so that a
    parity test exercises theta/phi/g/W.
    Drawn from numpy RandomState so that the stream is frozen across library versions.
    """
    rs = np.random.RandomState(seed)

    def u(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return torch.from_numpy(rs.uniform(-b, b, size=shape).astype(np.float32))

    w: Weights = {}
    for name in ("g", "theta", "phi"):
        w[f"newnlb.{name}.weight"] = u((D_INTER, D_MODEL, 1), D_MODEL)
        w[f"newnlb.{name}.bias"] = u((D_INTER,), D_MODEL)
    if randomize_W:
        w["newnlb.W.weight"] = u((D_MODEL, D_INTER, 1), D_INTER)
        w["newnlb.W.bias"] = u((D_MODEL,), D_INTER)
    else:
        w["newnlb.W.weight"] = torch.zeros(D_MODEL, D_INTER, 1)
        w["newnlb.W.bias"] = torch.zeros(D_MODEL)
    w["newnlb.concat_project.0.weight"] = u((1, 2 * D_INTER, 1, 1), 2 * D_INTER)
    w["attention_scorer.weight"] = u((1, D_MODEL), D_MODEL)
    w["attention_scorer.bias"] = u((1,), D_MODEL)
    w["last.weight"] = u((2, D_MODEL), D_MODEL)
    w["last.bias"] = u((2,), D_MODEL)
    return w


# --------------------------------------------------------------------------------------
# (a) temporal aggregation
# --------------------------------------------------------------------------------------
def nlb_forward(x: torch.Tensor, w: Weights) -> torch.Tensor:
    """Concatenation-form non-local block, 1-D, no sub-sampling, no BN.

    Follows ``_NonLocalBlockND.forward`` (models/nlb.py:66-101) as instantiated by
    ``NONLocalBlock1D(256, sub_sample=False, bn_layer=False)`` (models/match_head.py:87).
    x: (b, 256, t) fp32  ->  z: (b, 256, t).
    """
    b, _, t = x.shape
    # g(x) -> (b, t, 128)                                   nlb.py:74-75
    gx = F.conv1d(x, w["newnlb.g.weight"], w["newnlb.g.bias"]).reshape(b, D_INTER, -1).permute(0, 2, 1)
    # theta(x) (b,128,t,1) and phi(x) (b,128,1,t)           nlb.py:78-80
    th = F.conv1d(x, w["newnlb.theta.weight"], w["newnlb.theta.bias"]).reshape(b, D_INTER, -1, 1)
    ph = F.conv1d(x, w["newnlb.phi.weight"], w["newnlb.phi.bias"]).reshape(b, D_INTER, 1, -1)
    # broadcast both to (b,128,t,t), stack on channels, 1x1 conv to one channel, ReLU
    #                                                       nlb.py:82-90
    pair = torch.cat([th.repeat(1, 1, 1, t), ph.repeat(1, 1, t, 1)], dim=1)
    f = F.relu(F.conv2d(pair, w["newnlb.concat_project.0.weight"])).reshape(b, t, t)
    # normalise by the number of positions, aggregate g    nlb.py:92-95
    y = torch.matmul(f / f.size(-1), gx)
    y = y.permute(0, 2, 1).contiguous().reshape(b, D_INTER, t)
    # output projection + residual                          nlb.py:98-99
    return F.conv1d(y, w["newnlb.W.weight"], w["newnlb.W.bias"]) + x


def track_lengths_from_mask(mask: torch.Tensor) -> List[int]:
    """Number of real frames per track from the padding mask.

    models/match_head.py:136-139: the track ends at the first True of its mask row (or at
    the row length when there is none); row 0 of the sequence is a dummy, so a track with
    end index e owns rows 1..e-1.  Returns e-1 clamped at 0 for each track.
    """
    out = []
    m = mask.cpu().numpy().astype(bool)
    for i in range(m.shape[0]):
        nz = np.flatnonzero(m[i])
        end = int(nz[0]) if nz.size else m.shape[1]
        out.append(max(end - 1, 0))
    return out


def aggregate_tracks(seq: torch.Tensor, mask: torch.Tensor, w: Weights,
                     use_nlb: bool = True) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Seq-branch of ``TemporalAggregationNLB.forward`` up to the pooled descriptor.

    models/match_head.py:133-154.  seq: (1+Tmax, Q, 256) time-major, row 0 dummy;
    mask: (Q, 1+Tmax) bool, True = padding.  Returns (x3_1b (Q,256), [p_i (T_i,1)]).
    Per track: NLB when T_i > 1 (:144-147), attention logits = Linear(256,1), softmax over
    the T_i frames, weighted sum (:149-151).
    """
    lens = track_lengths_from_mask(mask)
    pooled, att = [], []
    wa, ba = w["attention_scorer.weight"], w["attention_scorer.bias"]
    for i, n in enumerate(lens):
        x = seq[1:1 + n, i]                                    # (T_i, 256)
        if use_nlb and x.shape[0] > 1:
            x = nlb_forward(x.transpose(0, 1).unsqueeze(0), w)[0].transpose(0, 1)
        p = F.softmax(F.linear(x, wa, ba), 0)                  # (T_i, 1)
        pooled.append((p * x).sum(0).unsqueeze(0))
        att.append(p)
    return torch.cat(pooled, 0), att


def nlb_collapsed_pool(x: torch.Tensor, w: Weights) -> torch.Tensor:
    """Algebraically collapsed NLB + attention pooling for ONE track, in float64.

    Not a reference function: the closed form the CUDA kernels implement (DESIGN.md,
    "K1 algebra"), restated here so the tests can bound the algebra's own rounding
    separately from the kernels'.  x: (T,256) -> (256,).
    """
    x = x.double()
    T = x.shape[0]
    if T == 0:
        return torch.zeros(D_MODEL, dtype=torch.float64)
    wa = w["attention_scorer.weight"].double()[0]
    ba = w["attention_scorer.bias"].double()[0]
    if T == 1:
        return x[0].clone()
    Wt = w["newnlb.theta.weight"].double()[:, :, 0]
    Wp = w["newnlb.phi.weight"].double()[:, :, 0]
    Wg = w["newnlb.g.weight"].double()[:, :, 0]
    WW = w["newnlb.W.weight"].double()[:, :, 0]
    bt, bp = w["newnlb.theta.bias"].double(), w["newnlb.phi.bias"].double()
    bg, bW = w["newnlb.g.bias"].double(), w["newnlb.W.bias"].double()
    wc = w["newnlb.concat_project.0.weight"].double().reshape(-1)
    a = x @ (Wt.T @ wc[:D_INTER]) + bt @ wc[:D_INTER]
    b = x @ (Wp.T @ wc[D_INTER:]) + bp @ wc[D_INTER:]
    v = WW.T @ wa
    c = x @ (Wg.T @ v) + bg @ v
    d = x @ wa
    f = torch.relu(a[:, None] + b[None, :]) / T
    s = d + f @ c + (bW @ wa + ba)
    p = torch.softmax(s, 0)
    q = p @ f
    r = q @ x
    return p @ x + WW @ (Wg @ r + bg * q.sum()) + bW


# --------------------------------------------------------------------------------------
# (b) pair scorer
# --------------------------------------------------------------------------------------
def pair_logits(q: torch.Tensor, g: torch.Tensor, w: Weights, chunk: int = 0) -> torch.Tensor:
    """x5 = last((q[:,None]-g[None])**2): models/match_head.py:156-162 (and :70-74).

    q: (Q,256), g: (G,256) fp32 -> (Q,G,2) fp32.  ``chunk`` bounds the (chunk,G,256)
    temporary; the arithmetic per pair is unchanged.
    """
    W, b = w["last.weight"], w["last.bias"]
    if chunk <= 0:
        chunk = max(1, min(q.shape[0], (1 << 27) // max(1, g.shape[0] * D_MODEL)))
    outs = []
    for s in range(0, q.shape[0], chunk):
        x4 = (q[s:s + chunk].unsqueeze(1) - g.unsqueeze(0)) ** 2
        outs.append(F.linear(x4, W, b))
    if not outs:
        return torch.zeros(0, g.shape[0], 2)
    return torch.cat(outs, 0)


def match_scores(x5: torch.Tensor) -> torch.Tensor:
    """Probability of the "match" class: softmax over the two logits, class 1.

    evaluate_movingfashion.py:265-267 (exp / sum-of-exp, class index 1), in fp32 torch.
    """
    return F.softmax(x5, dim=-1)[..., 1]


def logit_margin(x5: torch.Tensor) -> torch.Tensor:
    """d = l1 - l0.  softmax(l)[1] == sigmoid(d), so ranking by score == ranking by d
     but d does not saturate."""
    return x5[..., 1] - x5[..., 0]


def forward_seq_branch(seq, mask, x3_2, w: Weights, getatt: bool = False):
    """Whole ``TemporalAggregationNLB.forward`` seq-branch: models/match_head.py:133-169.

    Returns the reference's tuple (x3_1b, x3_2, x5, seq, mask, ids[, attention_scores]).
    """
    x3_1b, att = aggregate_tracks(seq, mask, w)
    x5 = pair_logits(x3_1b, x3_2.reshape(-1, D_MODEL) if x3_2.dim() == 1 else x3_2, w)
    ids = torch.zeros((1, 2))                                   # match_head.py:158
    if getatt:
        return x3_1b, x3_2, x5, seq, mask, ids, att
    return x3_1b, x3_2, x5, seq, mask, ids


# --------------------------------------------------------------------------------------
# (c) ranking
# --------------------------------------------------------------------------------------
def rank_topk(x5: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-query descending ranking, first k.

    evaluate_movingfashion.py:268 sorts the class-1 scores with ``np.argsort(...)[:, ::-1]``
    (non-stable quicksort then reversed: the order inside ties is unspecified).  The
    contract used for parity is: order by the logit margin d descending (a refinement of
    ordering by score, identical wherever fp32 scores differ), ties by lowest index.
    Returns (scores (Q,k) fp32, margins (Q,k) fp32, idx (Q,k) int64).
    """
    d = logit_margin(x5)
    s = match_scores(x5)
    k = min(k, d.shape[1])
    order = torch.argsort(d, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(s, 1, order), torch.gather(d, 1, order), order


def rank_of_target(x5: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Rank of a given gallery item for each query (0 = best):
    evaluate_movingfashion.py:268-269  ``(rankings == shop_prod_index).nonzero()[1]``.
    Counted on the margin: number of items strictly better, plus equal items with a lower
    index (the same tie contract as ``rank_topk``)."""
    d = logit_margin(x5)
    dt = d.gather(1, target.view(-1, 1))
    better = (d > dt).sum(1)
    idx = torch.arange(d.shape[1]).view(1, -1)
    tie_before = ((d == dt) & (idx < target.view(-1, 1))).sum(1)
    return better + tie_before


def merge_topk(score_lists: Sequence[torch.Tensor], margin_lists: Sequence[torch.Tensor],
               idx_lists: Sequence[torch.Tensor], k: int):
    """Top-k of a union of per-shard top-k lists (the multi-GPU merge; no reference
    counterpart -- the reference ranks one whole gallery, evaluate_movingfashion.py:268).
    Invalid entries carry idx < 0 and are ordered last."""
    s = torch.cat(list(score_lists), 1)
    d = torch.cat(list(margin_lists), 1)
    i = torch.cat(list(idx_lists), 1)
    dkey = torch.where(i < 0, torch.full_like(d, -float("inf")), d)
    # order by margin desc, then index asc
    o1 = torch.argsort(i, dim=1, stable=True)
    dk = torch.gather(dkey, 1, o1)
    o2 = torch.argsort(dk, dim=1, descending=True, stable=True)
    order = torch.gather(o1, 1, o2)[:, :k]
    return torch.gather(s, 1, order), torch.gather(d, 1, order), torch.gather(i, 1, order)


# --------------------------------------------------------------------------------------
# evaluation-script (numpy) variants -- secondary oracle
# --------------------------------------------------------------------------------------
def eval_aggr_scores_np(shop_aggr_f16: np.ndarray, aggr_desc_f32: np.ndarray,
                        aggrW_f16: np.ndarray, aggrB_f16: np.ndarray) -> np.ndarray:
    """Aggregated-descriptor scorer of the eval script, one query.

    evaluate_movingfashion.py:263-267 (evaluate_multiDF2.py:220-224): fp16 gallery and
    fp16 ``last`` weights (:123-124) against an fp32 query, so numpy promotes to fp32.
    Returns class-1 scores, shape (1, G).
    """
    sq = (shop_aggr_f16[np.newaxis] - aggr_desc_f32[np.newaxis, np.newaxis]) ** 2
    raw = sq @ aggrW_f16.transpose() + aggrB_f16
    e = np.exp(raw)
    return (e / e.sum(2)[:, :, np.newaxis])[:, :, 1]


def eval_frame_scores_np(shop_f16: np.ndarray, street_f16: np.ndarray,
                         w_f32: np.ndarray, b_f32: np.ndarray) -> np.ndarray:
    """Per-frame scorer of the eval script (pure fp16 numpy).

    evaluate_movingfashion.py:94-99 ``compute_ranking`` / :101-106 ``compute_distances``:
    shop_mat (G,256) fp16, street rows (n,256) fp16, ``match_predictor.last`` w, b cast to
    fp16.  Returns (n, G) fp16 class-1 scores.
    """
    sq = (shop_f16[np.newaxis] - street_f16[:, np.newaxis]) ** 2
    raw = sq @ w_f32.transpose().astype(np.float16) + b_f32.astype(np.float16)
    e = np.exp(raw)
    return (e / e.sum(2)[:, :, np.newaxis])[:, :, 1]


def eval_rankings_np(scores: np.ndarray) -> np.ndarray:
    """evaluate_movingfashion.py:98 / :268: ``np.argsort(scores, 1)[:, ::-1]``."""
    return np.argsort(scores, 1)[:, ::-1]


def topk_hits(rank: int, k_thresholds=(1, 5, 10, 20)) -> List[int]:
    """evaluate_movingfashion.py:270-272: hit at threshold k iff rank < k."""
    return [1 if rank < k else 0 for k in k_thresholds]


def eval_product_loop(frame_desc: torch.Tensor, frame_product: torch.Tensor, shop_desc: torch.Tensor,
                      target: torch.Tensor, w_frame: Weights, aggr_desc: torch.Tensor, shop_aggr: torch.Tensor,
                      w_aggr: Weights, k_thresholds=(1, 5, 10, 20)):
    """The per-product loop of the eval script, literally one product at a time, in fp32 (the reference
    runs the frame-level parts in numpy fp16, evaluate_movingfashion.py:82-100).

    For product p with frames F_p (``frame_product == p``) and true shop row t = target[p]:
      * every frame on its own: ``compute_ranking`` :95-100, rank :216-222, hits ``k_accs`` :223-232;
      * best frame of the product ("Product Max", ``k_accs_avg``) :233-241;
      * average descriptor ``street_mat[best_inds].mean(0)`` :279-292 (``k_accs_avg_desc``);
      * aggregated descriptor :252-277 (``k_accs_aggr_desc``), scored with the aggregator's ``last``;
      * average / maximum of the frames' class-1 probabilities :294-316 (``k_accs_avg_dist``, ``k_accs_max_dist``).
    Ranks are positions in the descending order with ties by lower index first (the reference's
    ``argsort(...)[::-1]`` leaves ties unspecified).  Returns a dict of int64 tensors: ``frame_ranks (N,)``,
    ``best``, ``avg_desc``, ``aggr``, ``avg_dist``, ``max_dist`` (each (P,), G for products without frames
    except ``aggr``), and ``hits`` (6, len(k_thresholds)) in the order frame, best, avg_desc, aggr, avg_dist,
    max_dist."""
    P, G = int(target.shape[0]), int(shop_desc.shape[0])
    col = torch.arange(G)

    def rank_in(scores_1d: torch.Tensor, t: int) -> int:
        return int(((scores_1d > scores_1d[t]) | ((scores_1d == scores_1d[t]) & (col < t))).sum())

    out = {k: torch.full((P,), G, dtype=torch.int64) for k in ("best", "avg_desc", "aggr", "avg_dist", "max_dist")}
    frame_ranks = torch.zeros(frame_desc.shape[0], dtype=torch.int64)
    for p in range(P):
        t = int(target[p])
        d_aggr = logit_margin(pair_logits(aggr_desc[p:p + 1], shop_aggr, w_aggr))[0]
        out["aggr"][p] = rank_in(d_aggr, t)
        rows = (frame_product == p).nonzero().flatten()
        if rows.numel() == 0:
            continue
        x5 = pair_logits(frame_desc[rows], shop_desc, w_frame)
        d = logit_margin(x5)
        for i, r in enumerate(rows.tolist()):
            frame_ranks[r] = rank_in(d[i], t)
        out["best"][p] = int(frame_ranks[rows].min())
        avg = frame_desc[rows].mean(0, keepdim=True)
        out["avg_desc"][p] = rank_in(logit_margin(pair_logits(avg, shop_desc, w_frame))[0], t)
        prob = match_scores(x5)
        out["avg_dist"][p] = rank_in(prob.mean(0), t)
        out["max_dist"][p] = rank_in(prob.max(0).values, t)
    rows_for_hits = [frame_ranks, out["best"], out["avg_desc"], out["aggr"], out["avg_dist"], out["max_dist"]]
    out["hits"] = torch.tensor([[int((r < k).sum()) for k in k_thresholds] for r in rows_for_hits])
    out["frame_ranks"] = frame_ranks
    return out


# --------------------------------------------------------------------------------------
# synthetic workloads; shared by tests and bench
# --------------------------------------------------------------------------------------
def synth_tracks(Q: int, Tmax: int, seed: int, ragged: Optional[Tuple[int, int]] = None):
    """(seq (1+Tmax,Q,256) fp32 time-major with dummy row 0, mask (Q,1+Tmax) bool, lens).
    ``ragged=(lo,hi)`` draws T_i uniformly in [lo,hi]; padding rows stay zero."""
    rs = np.random.RandomState(seed)
    seq = np.zeros((1 + Tmax, Q, D_MODEL), np.float32)
    seq[1:] = rs.randn(Tmax, Q, D_MODEL).astype(np.float32)
    if ragged is None:
        lens = np.full(Q, Tmax, np.int64)
    else:
        lens = rs.randint(ragged[0], ragged[1] + 1, size=Q)
    mask = np.zeros((Q, 1 + Tmax), bool)
    for i, n in enumerate(lens):
        mask[i, 1 + n:] = True
        seq[1 + n:, i] = 0.0
    return torch.from_numpy(seq), torch.from_numpy(mask), lens


def synth_gallery(G: int, seed: int, planted: Optional[torch.Tensor] = None, noise: float = 0.1):
    """randn(G,256); the first min(Q,G) rows are overwritten with ``planted[i] + noise*randn``
    (a true match per query: non-trivial top-1 and worst-case cancellation for the expanded
    form)."""
    rs = np.random.RandomState(seed + 1000003)
    g = rs.randn(G, D_MODEL).astype(np.float32)
    if planted is not None:
        n = min(planted.shape[0], G)
        g[:n] = planted[:n].numpy() + noise * rs.randn(n, D_MODEL).astype(np.float32)
    return torch.from_numpy(g)
