"""The CPU oracle against the golden vectors produced by the UNMODIFIED reference modules
(tests/golden/make_golden.py).  The reference has no tests of its own; these goldens are the pin."""
import numpy as np
import pytest
import torch

from oracle import seam_oracle as so
from util import GOLDEN_CASES, case_inputs

torch.set_num_threads(1)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_seq_branch_matches_reference(name, golden, weights):
    case = GOLDEN_CASES[name]
    seq, mask, lens, gal = case_inputs(case, weights)
    assert np.array_equal(np.asarray(lens, np.int32), golden[f"{name}.lens"])
    x3_1b, x3_2, x5, s_out, m_out, ids, att = so.forward_seq_branch(seq, mask, gal, weights, getatt=True)
    assert ids.shape == (1, 2)                      # models/match_head.py:158
    assert s_out is seq and m_out is mask
    # same library, same op order: expect (near) bit equality with the reference outputs
    np.testing.assert_allclose(x3_1b.numpy(), golden[f"{name}.x3_1b"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(x5[:4].numpy(), golden[f"{name}.x5_head"], rtol=0, atol=1e-5)
    attp = np.zeros((case["Q"], case["Tmax"]), np.float32)
    for i, p in enumerate(att):
        attp[i, :p.shape[0]] = p[:, 0].numpy()
    np.testing.assert_allclose(attp, golden[f"{name}.att"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(so.match_scores(x5).double().sum(1).numpy(), golden[f"{name}.score_sum"], rtol=1e-6)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_ranking_matches_reference(name, golden, weights):
    case = GOLDEN_CASES[name]
    seq, mask, _, gal = case_inputs(case, weights)
    x5 = so.forward_seq_branch(seq, mask, gal, weights)[2]
    s, d, idx = so.rank_topk(x5, 20)
    assert np.array_equal(idx.numpy().astype(np.int32), golden[f"{name}.topk_idx"])
    np.testing.assert_allclose(s.numpy(), golden[f"{name}.topk_score"], atol=1e-6)
    np.testing.assert_allclose(d.numpy(), golden[f"{name}.topk_margin"], atol=1e-5)
    # the margin ordering refines the reference's score ordering (evaluate_movingfashion.py:268)
    sc = so.match_scores(x5)
    ref_rank = so.eval_rankings_np(sc.numpy())
    k = idx.shape[1]
    for i in range(x5.shape[0]):
        assert np.allclose(sc[i, ref_rank[i, :k].copy()].numpy(), s[i].numpy(), atol=1e-7)


@pytest.mark.parametrize("t", [2, 7, 10])
def test_nlb_matches_reference(t, golden, weights):
    x = torch.from_numpy(golden[f"nlb.t{t}.x"])
    np.testing.assert_allclose(so.nlb_forward(x, weights).numpy(), golden[f"nlb.t{t}.z"], rtol=0, atol=1e-6)


def test_default_init_is_identity(weights):
    """models/nlb.py:48-49 zero-initialises W: the block is the identity at default init."""
    w0 = so.random_weights(seed=0, randomize_W=False)
    x = torch.randn(2, 256, 6)
    assert torch.equal(so.nlb_forward(x, w0), x)


@pytest.mark.parametrize("T", [0, 1, 2, 4, 10, 64])
def test_collapsed_algebra(T, weights):
    """The closed form the kernels implement equals the reference formulation (fp64 vs fp32)."""
    rs = np.random.RandomState(T)
    x = torch.from_numpy(rs.randn(T, 256).astype(np.float32))
    seq = torch.zeros(1 + T, 1, 256)
    seq[1:, 0] = x
    mask = torch.zeros(1, 1 + T, dtype=torch.bool)
    ref, _ = so.aggregate_tracks(seq, mask, weights)
    got = so.nlb_collapsed_pool(x, weights).float()
    assert (ref[0] - got).abs().max() < 5e-6
    if T == 1:
        assert torch.equal(ref[0], x[0])            # NLB bypass: models/match_head.py:145-147


def test_mask_semantics():
    """First True ends the track, later False entries are ignored, row 0 is the dummy
    (models/match_head.py:136-139)."""
    m = torch.tensor([[0, 0, 0, 0], [0, 0, 1, 0], [1, 0, 0, 0], [0, 1, 1, 1], [0, 0, 0, 1]], dtype=torch.bool)
    assert so.track_lengths_from_mask(m) == [3, 1, 0, 0, 2]


def test_eval_script_scorer_matches_golden(golden, weights):
    """evaluate_movingfashion.py:263-267 restated (fp16 gallery / weights, fp32 query)."""
    case = GOLDEN_CASES["cfg1"]
    seq, mask, _, gal = case_inputs(case, weights)
    q, _ = so.aggregate_tracks(seq, mask, weights)
    g16 = gal.numpy().astype(np.float16)
    aW = weights["last.weight"].numpy().astype(np.float16)
    aB = weights["last.bias"].numpy().astype(np.float16)
    ev = np.stack([so.eval_aggr_scores_np(g16, q[i].numpy(), aW, aB)[0] for i in range(4)])
    np.testing.assert_allclose(ev, golden["cfg1.eval_scores_head"], atol=1e-6)
    # looser agreement with the fp32 module path: same top-1 (the planted match)
    sc = so.match_scores(so.pair_logits(q[:4], gal, weights)).numpy()
    assert (ev.argmax(1) == sc.argmax(1)).all()
    assert np.abs(ev - sc).max() < 2e-2


def test_rank_of_target_and_hits(weights):
    rs = np.random.RandomState(3)
    q = torch.from_numpy(rs.randn(6, 256).astype(np.float32))
    g = torch.from_numpy(rs.randn(50, 256).astype(np.float32))
    x5 = so.pair_logits(q, g, weights)
    rk = so.eval_rankings_np(so.logit_margin(x5).numpy())
    tgt = torch.tensor([0, 7, 49, 3, 3, 20])
    mine = so.rank_of_target(x5, tgt)
    for i in range(6):
        assert int(mine[i]) == int((rk[i] == int(tgt[i])).nonzero()[0][0])    # evaluate_movingfashion.py:269
    assert so.topk_hits(0) == [1, 1, 1, 1] and so.topk_hits(5) == [0, 0, 1, 1] and so.topk_hits(20) == [0, 0, 0, 0]


def test_merge_topk_equals_global(weights):
    rs = np.random.RandomState(5)
    q = torch.from_numpy(rs.randn(7, 256).astype(np.float32))
    g = torch.from_numpy(rs.randn(90, 256).astype(np.float32))
    x5 = so.pair_logits(q, g, weights)
    bounds = [(0, 10), (10, 55), (55, 90)]
    lists = [so.rank_topk(x5[:, a:b], 20) for a, b in bounds]
    # shard 0 has only 10 items: pad to k with invalid entries
    S, M, I = [], [], []
    for (s, d, i), (a, b) in zip(lists, bounds):
        pad = 20 - s.shape[1]
        S.append(torch.cat([s, torch.zeros(7, pad)], 1))
        M.append(torch.cat([d, torch.full((7, pad), -float("inf"))], 1))
        I.append(torch.cat([i + a, torch.full((7, pad), -1, dtype=torch.long)], 1))
    s, d, i = so.merge_topk(S, M, I, 20)
    s_ref, d_ref, i_ref = so.rank_topk(x5, 20)
    assert torch.equal(i, i_ref) and torch.equal(d, d_ref) and torch.equal(s, s_ref)


def test_eval_product_loop_matches_the_numpy_formulas(weights):
    """The fp32 restatement of the eval script's per-product loop against the script's own numpy
    expressions (evaluate_movingfashion.py:94-106, 263-268, 279-316) evaluated in fp32/fp64: identical
    ranks wherever the margins are not tied, and the hit counts that follow."""
    rs = np.random.RandomState(21)
    P, G = 9, 60
    lens = [3, 1, 0, 4, 2, 5, 1, 2, 3]
    frame_product = torch.tensor([p for p in range(P) for _ in range(lens[p])])
    frame_desc = torch.from_numpy(rs.randn(len(frame_product), 256).astype(np.float32))
    shop_desc = torch.from_numpy(rs.randn(G, 256).astype(np.float32))
    shop_aggr = torch.from_numpy(rs.randn(G, 256).astype(np.float32))
    aggr_desc = torch.from_numpy(rs.randn(P, 256).astype(np.float32))
    target = torch.from_numpy(rs.permutation(G)[:P].astype(np.int64))
    wf = dict(weights)
    wf["last.weight"] = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2, 256)).astype(np.float32))
    wf["last.bias"] = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2,)).astype(np.float32))
    got = so.eval_product_loop(frame_desc, frame_product, shop_desc, target, wf, aggr_desc, shop_aggr, weights)

    def np_scores(shop, street, w, b):                       # compute_distances, :101-106, in float64
        sq = (shop[np.newaxis] - street[:, np.newaxis]) ** 2
        raw = sq @ w.transpose() + b
        e = np.exp(raw)
        return (e / e.sum(2)[:, :, np.newaxis])[:, :, 1]

    w64, b64 = wf["last.weight"].double().numpy(), wf["last.bias"].double().numpy()
    aw, ab = weights["last.weight"].double().numpy(), weights["last.bias"].double().numpy()
    shop64, saggr64 = shop_desc.double().numpy(), shop_aggr.double().numpy()
    for p in range(P):
        t = int(target[p])
        rows = (frame_product == p).nonzero().flatten().numpy()
        s_aggr = np_scores(saggr64, aggr_desc[p:p + 1].double().numpy(), aw, ab)[0]
        assert int(got["aggr"][p]) == int((np.argsort(s_aggr)[::-1] == t).nonzero()[0][0])     # :268-269
        if len(rows) == 0:
            assert int(got["best"][p]) == G
            continue
        sc = np_scores(shop64, frame_desc[rows].double().numpy(), w64, b64)
        ranks = [(np.argsort(sc[i])[::-1] == t).nonzero()[0][0] for i in range(len(rows))]   # :95-100, :216-222
        assert got["frame_ranks"][rows].tolist() == [int(r) for r in ranks]
        assert int(got["best"][p]) == int(min(ranks))
        avg = frame_desc[rows].double().numpy().mean(0, keepdims=True)                         # :280
        assert int(got["avg_desc"][p]) == int((np.argsort(np_scores(shop64, avg, w64, b64)[0])[::-1] == t).nonzero()[0][0])
        assert int(got["avg_dist"][p]) == int((np.argsort(sc.mean(0))[::-1] == t).nonzero()[0][0])   # :296-298
        assert int(got["max_dist"][p]) == int((np.argsort(sc.max(0))[::-1] == t).nonzero()[0][0])    # :307-309
    assert got["hits"].shape == (6, 4)
    assert got["hits"][0, 3] == int((got["frame_ranks"] < 20).sum())
