"""Shared helpers for the parity tests."""
import torch

from oracle import seam_oracle as so

# name -> generation parameters; must match tests/golden/make_golden.py::CASES
GOLDEN_CASES = {
    "cfg1":   dict(Q=64, Tmax=10, ragged=None,   G=1000, seed=0, planted=True),
    "ragged": dict(Q=37, Tmax=4,  ragged=(0, 4), G=300,  seed=2, planted=True),
    "t1":     dict(Q=5,  Tmax=1,  ragged=None,   G=33,   seed=5, planted=False),
    "t64":    dict(Q=8,  Tmax=64, ragged=None,   G=513,  seed=3, planted=True),
    "tiny_g": dict(Q=9,  Tmax=3,  ragged=(1, 3), G=7,    seed=7, planted=False),
}

# Stated tolerances (DESIGN.md "Parity"): fp32 path end to end.
TOL_EMB = 2e-5      # aggregated embedding, absolute, on O(1) values
TOL_ATT = 2e-6      # attention weights
TOL_LOGIT = 3e-5    # logits / margins, absolute, on O(1..30) values
TOL_SCORE = 1e-5    # softmax score


def case_inputs(case, w):
    """Regenerate the inputs of a golden case exactly as make_golden.py did."""
    seq, mask, lens = so.synth_tracks(case["Q"], case["Tmax"], case["seed"], case["ragged"])
    planted = None
    if case["planted"]:
        planted, _ = so.aggregate_tracks(seq, mask, w)
    gal = so.synth_gallery(case["G"], case["seed"], planted)
    return seq, mask, lens, gal


def assert_topk_matches(idx, margins, scores, d_full, s_full, k, tol=TOL_LOGIT):
    """Top-k parity up to ties: `idx/margins/scores` (Q,k) from the device against the
    oracle's full (Q,G) margin and score matrices.

    - reported margins / scores equal the oracle's at the reported indices (within tol);
    - the list is ordered by margin descending;
    - wherever the index list differs from the oracle's, the oracle margins at the
      differing positions are within tol of each other (a tie inside the tolerance).
    """
    Q, G = d_full.shape
    kk = min(k, G)
    idx = idx.long().cpu()
    margins, scores = margins.cpu(), scores.cpu()
    assert (idx[:, kk:] == -1).all(), "entries beyond G must be padded with -1"
    idx, margins, scores = idx[:, :kk], margins[:, :kk], scores[:, :kk]
    assert (idx >= 0).all() and (idx < G).all()
    # no duplicates
    assert all(len(set(r.tolist())) == kk for r in idx)
    d_at = torch.gather(d_full, 1, idx)
    s_at = torch.gather(s_full, 1, idx)
    assert (margins - d_at).abs().max() <= tol, f"margin error {(margins - d_at).abs().max():.3e}"
    assert (scores - s_at).abs().max() <= TOL_SCORE, f"score error {(scores - s_at).abs().max():.3e}"
    assert (margins[:, :-1] >= margins[:, 1:]).all(), "not sorted by margin"
    order = torch.argsort(d_full, dim=1, descending=True, stable=True)[:, :kk]
    d_ref = torch.gather(d_full, 1, order)
    differs = (order != idx)
    if differs.any():
        assert ((d_ref - d_at).abs()[differs] <= 2 * tol).all(), \
            "index lists differ where the oracle margins are NOT tied within tolerance"
    return int(differs.any(1).sum())
