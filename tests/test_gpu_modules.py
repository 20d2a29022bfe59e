"""The drop-in modules on the GPU: same signatures, state_dict keys and return tuples as the
reference's (models/match_head.py:47-169, models/nlb.py:104-109), results against the oracle."""
import numpy as np
import pytest
import torch

import seam_match_rcnn_b200 as pkg
from oracle import seam_oracle as so
from util import GOLDEN_CASES, TOL_ATT, TOL_EMB, TOL_LOGIT, case_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def model(weights):
    m = pkg.TemporalAggregationNLB().to(DEV).eval()
    missing, unexpected = m.load_state_dict(weights, strict=False)
    assert not unexpected
    return m


@pytest.mark.parametrize("name", ["cfg1", "ragged", "t1"])
def test_seq_branch_tuple(name, model, weights, golden):
    case = GOLDEN_CASES[name]
    seq, mask, _, gal = case_inputs(case, weights)
    out = model(None, None, None, x3_1_seq=seq.to(DEV), x3_1_mask=mask.to(DEV), x3_2=gal.to(DEV), getatt=True)
    assert len(out) == 7
    x3_1b, x3_2, x5, s_out, m_out, ids, att = out
    assert ids.shape == (1, 2) and x5.shape == (case["Q"], case["G"], 2)
    assert (x3_1b.cpu() - torch.from_numpy(golden[f"{name}.x3_1b"])).abs().max() <= TOL_EMB
    assert (x5[:4].cpu() - torch.from_numpy(golden[f"{name}.x5_head"])).abs().max() <= TOL_LOGIT
    assert len(att) == case["Q"]
    for i, p in enumerate(att):
        n = int(golden[f"{name}.lens"][i])
        assert p.shape == (n, 1)
        if n:
            assert (p[:, 0].cpu() - torch.from_numpy(golden[f"{name}.att"][i, :n])).abs().max() <= TOL_ATT
    assert len(model(None, None, None, x3_1_seq=seq.to(DEV), x3_1_mask=mask.to(DEV), x3_2=gal.to(DEV))) == 6


def test_eval_script_call_shape(model, weights):
    """The exact call of evaluate_movingfashion.py:253-262: one track, 1-D x3_2, [0][0]."""
    seq, mask, _ = so.synth_tracks(1, 10, seed=9)
    shop = torch.randn(256)
    r = model(None, None, None, x3_1_seq=seq.to(DEV), x3_1_mask=mask.to(DEV), x3_2=shop.to(DEV))
    ref = so.forward_seq_branch(seq, mask, shop, weights)
    assert (r[0][0].cpu() - ref[0][0]).abs().max() <= TOL_EMB
    assert r[2].shape == ref[2].shape == (1, 1, 2)
    assert (r[2].cpu() - ref[2]).abs().max() <= TOL_LOGIT


def test_score_topk_entry(model, weights, golden):
    case = GOLDEN_CASES["cfg1"]
    seq, mask, _, gal = case_inputs(case, weights)
    scores, idx = model.score_topk(seq.to(DEV), mask.to(DEV), gal.to(DEV), k=20)
    assert idx.dtype == torch.int64
    assert torch.equal(idx.cpu(), torch.from_numpy(golden["cfg1.topk_idx"]).long())
    assert (scores.cpu() - torch.from_numpy(golden["cfg1.topk_score"])).abs().max() <= 1e-5
    with pytest.raises(pkg.SeamError):          # the dense tuple is refused beyond the size limit
        model(None, None, None, x3_1_seq=torch.zeros(2, 9000, 256, device=DEV),
              x3_1_mask=torch.zeros(9000, 2, dtype=torch.bool, device=DEV), x3_2=torch.zeros(9000, 256, device=DEV))


def test_x_branch_groups_tracks_like_the_reference(model, weights):
    """x-branch (models/match_head.py:92-131): ROI features -> tower -> group street rows by id."""
    torch.manual_seed(0)
    x = torch.randn(9, 256, 14, 14, device=DEV)
    types = torch.tensor([1, 0, 0, 0, 1, 0, 0, 0, 0])
    ids = torch.tensor([0, 5, 2, 5, 0, 2, 5, 9, 2])
    x3_1b, x3_2, x5, seq, mask, out_ids = model(x, types, ids)
    x3 = model.embed(x)
    st_ids = ids[types == 0]
    uniq = st_ids.unique()
    assert seq.shape == (1 + 3, 3, 256) and mask.shape == (3, 4)
    for i, idd in enumerate(uniq):
        rows = x3[types.to(DEV) == 0][st_ids.to(DEV) == idd]
        n = rows.shape[0]
        assert torch.equal(seq[1:1 + n, i], rows) and not mask[i, :n + 1].any() and mask[i, n + 1:].all()
    ref, _ = so.aggregate_tracks(seq.cpu(), mask.cpu(), weights)
    assert (x3_1b.cpu() - ref).abs().max() <= TOL_EMB
    assert x5.shape == (3, 2, 2) and torch.equal(x3_2, x3[types.to(DEV) == 1])
    assert torch.equal(out_ids.cpu(), st_ids)
    # no street items at all -> None outputs (match_head.py:126-128, 163-164)
    r = model(x[:2], torch.tensor([1, 1]), torch.tensor([0, 1]))
    assert r[0] is None and r[2] is None


def test_match_predictor_forward(weights):
    m = pkg.MatchPredictor().to(DEV).eval()
    m.last.load_state_dict({"weight": weights["last.weight"], "bias": weights["last.bias"]})
    torch.manual_seed(1)
    x = torch.randn(7, 256, 14, 14, device=DEV)
    types = torch.tensor([0, 0, 1, 1, 1, 0, 1])
    x3, x5 = m(x, types)
    assert x3.shape == (7, 256) and x5.shape == (3, 4, 2)
    ref = so.pair_logits(x3[types.to(DEV) == 0].cpu(), x3[types.to(DEV) == 1].cpu(), weights)
    assert (x5.cpu() - ref).abs().max() <= TOL_LOGIT


def test_nonlocal_block_module(weights, golden):
    nlb = pkg.NONLocalBlock1D(in_channels=256, sub_sample=False, bn_layer=False).to(DEV).eval()
    x = torch.from_numpy(golden["nlb.t7.x"]).to(DEV)
    assert torch.equal(nlb(x), x)               # zero-initialised W: identity (models/nlb.py:48-49)
    nlb.load_state_dict({k[len("newnlb."):]: v for k, v in weights.items() if k.startswith("newnlb.")})
    assert (nlb(x).cpu() - torch.from_numpy(golden["nlb.t7.z"])).abs().max() <= TOL_EMB
    with pytest.raises(NotImplementedError):
        pkg.NONLocalBlock1D(in_channels=64)


def test_weight_updates_are_picked_up(model, weights):
    seq, mask, _ = so.synth_tracks(4, 5, seed=1)
    a = model.aggregate(seq.to(DEV), mask.to(DEV)).clone()
    saved = model.newnlb.W.weight.detach().clone()
    with torch.no_grad():
        model.newnlb.W.weight.mul_(0.5)
    b = model.aggregate(seq.to(DEV), mask.to(DEV))
    assert not torch.equal(a, b)
    w2 = dict(weights)
    w2["newnlb.W.weight"] = weights["newnlb.W.weight"] * 0.5
    ref, _ = so.aggregate_tracks(seq, mask, w2)
    assert (b.cpu() - ref).abs().max() <= TOL_EMB
    with torch.no_grad():
        model.newnlb.W.weight.copy_(saved)
    model.nlb = False                            # reference flag: skip the block (match_head.py:113)
    c = model.aggregate(seq.to(DEV), mask.to(DEV))
    ref0, _ = so.aggregate_tracks(seq, mask, weights, use_nlb=False)
    assert (c.cpu() - ref0).abs().max() <= TOL_EMB
    model.nlb = True


def test_evaluate_aggregated_report(model, weights):
    """evaluate_movingfashion.py:252-277 for all products at once."""
    Q, T, G = 40, 6, 120
    seq, mask, _ = so.synth_tracks(Q, T, seed=2, ragged=(1, 6))
    qref, _ = so.aggregate_tracks(seq, mask, weights)
    gal = so.synth_gallery(G, 2, None)
    target = torch.arange(Q) * 2
    eng = model._engine_for(torch.device(DEV))
    model._sync_weights(eng)                     # an earlier test left the engine with nlb=False weights
    rep = pkg.evaluate_aggregated(eng, seq.to(DEV), mask.to(DEV), gal.to(DEV), target)
    x5 = so.pair_logits(qref, gal, weights)
    ranks = so.rank_of_target(x5, target)
    assert torch.equal(rep.ranks.cpu().long(), ranks)
    assert rep.hits == [int((ranks < k).sum()) for k in (1, 5, 10, 20)]


@pytest.mark.parametrize("pinned", [True, False])
def test_search_from_host_memory(model, weights, pinned):
    """retrieval.search_host: host tensors in, host tensors out, tracks streamed in slices (row 0 of
    x3_1_seq, the dummy frame, filled with NaN: it must never be read or copied); same results as
    the device-resident search and as the oracle."""
    Q, T, G, k = 203, 7, 777, 20
    seq, mask, _ = so.synth_tracks(Q, T, seed=11, ragged=(1, 7))
    seq[0] = float("nan")
    qref, _ = so.aggregate_tracks(seq, mask, weights)
    gal = so.synth_gallery(G, 11, qref)
    eng = model._engine_for(torch.device(DEV))
    model._sync_weights(eng)
    hs, hm, hg = (t.pin_memory() if pinned else t for t in (seq, mask, gal))
    out = pkg.search_host(eng, hs, hm, hg, k, stream=pkg.HostTrackStream(eng, nchunk=3))
    torch.cuda.synchronize()
    s_d, m_d, i_d = pkg.search(eng, seq.to(DEV), mask.to(DEV), gal.to(DEV), k)
    assert torch.equal(out[2], i_d.cpu())
    assert torch.equal(out[1], m_d.cpu()) and torch.equal(out[0], s_d.cpu())
    s_ref, d_ref, i_ref = so.rank_topk(so.pair_logits(qref, gal, weights), k)
    assert torch.equal(out[2].long(), i_ref)
    assert (out[1] - d_ref).abs().max() <= TOL_LOGIT


def test_evaluate_products_rows(model, weights):
    """evaluate_movingfashion.py:157-330 for all products at once: per-frame, best-frame, average-descriptor
    and aggregated-descriptor ranks / accuracy rows against the same formulas in torch fp32."""
    P, T, G = 31, 6, 140
    seq, mask, lens = so.synth_tracks(P, T, seed=4, ragged=(1, 6))
    lens = [int(x) for x in lens]
    rs = np.random.RandomState(4)
    frame_product = torch.tensor([p for p in range(P) for _ in range(lens[p])])
    frame_desc = torch.from_numpy(rs.randn(len(frame_product), 256).astype(np.float32))
    shop_desc = torch.from_numpy(rs.randn(G, 256).astype(np.float32))
    target = torch.from_numpy(rs.permutation(G)[:P].astype(np.int64))
    shop_desc[target[frame_product[::3]]] = frame_desc[::3] + 0.3 * torch.from_numpy(
        rs.randn(len(frame_desc[::3]), 256).astype(np.float32))          # some frames resemble their product
    qref, _ = so.aggregate_tracks(seq, mask, weights)
    shop_aggr = so.synth_gallery(G, 4, None)
    shop_aggr[target] = qref + 0.2 * torch.from_numpy(rs.randn(P, 256).astype(np.float32))
    fw = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2, 256)).astype(np.float32))
    fb = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2,)).astype(np.float32))
    aw, ab = weights["last.weight"], weights["last.bias"]

    eng = model._engine_for(torch.device(DEV))
    model._sync_weights(eng)
    rep = pkg.evaluate_products(eng, frame_desc, frame_product, shop_desc, target, (fw, fb),
                                seq.to(DEV), mask.to(DEV), shop_aggr, (aw, ab))

    wf = dict(weights)
    wf["last.weight"], wf["last.bias"] = fw, fb
    fr = so.rank_of_target(so.pair_logits(frame_desc, shop_desc, wf), target[frame_product])
    best = torch.stack([fr[frame_product == p].min() for p in range(P)])
    avg = torch.stack([frame_desc[frame_product == p].mean(0) for p in range(P)])
    ar = so.rank_of_target(so.pair_logits(avg, shop_desc, wf), target)
    gr = so.rank_of_target(so.pair_logits(qref, shop_aggr, weights), target)
    assert torch.equal(rep.frame_ranks.cpu().long(), fr)
    assert torch.equal(rep.product_ranks.cpu().long(), torch.stack([best, ar, gr]))
    ks = (1, 5, 10, 20)
    want = torch.tensor([[float((r < k).sum()) / n * 100 for k in ks]
                         for r, n in ((fr, len(fr)), (best, P), (ar, P), (gr, P))])
    assert torch.allclose(rep.perf, want, atol=1e-4)
    assert abs(rep.ret[2] - float((gr < 1).sum()) / P) < 1e-6
    assert rep.rank_median == float(torch.quantile(fr.float(), 0.5))
    # somebody else's scorer in the module's engine must not leak into the module's next call
    eng.load_scorer(fw, fb)
    case = GOLDEN_CASES["t1"]
    seq1, mask1, _, gal1 = case_inputs(case, weights)
    _, idx1 = model.score_topk(seq1.to(DEV), mask1.to(DEV), gal1.to(DEV), k=5)
    ref_q1, _ = so.aggregate_tracks(seq1, mask1, weights)
    assert torch.equal(idx1.cpu(), so.rank_topk(so.pair_logits(ref_q1, gal1, weights), 5)[2])


def test_evaluate_distance_fusions(model, weights):
    """evaluate_movingfashion.py:294-316: ranks of the true shop item under the per-product average /
    maximum of the frames' class-1 probabilities, against the same formulas in torch fp32 on the CPU."""
    P, G = 23, 90
    rs = np.random.RandomState(6)
    lens = rs.randint(1, 6, size=P)
    lens[5] = 0                                              # a product without tracked frames
    frame_product = torch.tensor([p for p in range(P) for _ in range(int(lens[p]))])
    perm = torch.from_numpy(rs.permutation(len(frame_product)))
    frame_product = frame_product[perm]                      # frames arrive in any order
    frame_desc = torch.from_numpy(rs.randn(len(frame_product), 256).astype(np.float32))
    shop_desc = torch.from_numpy(rs.randn(G, 256).astype(np.float32))
    target = torch.from_numpy(rs.permutation(G)[:P].astype(np.int64))
    fw = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2, 256)).astype(np.float32))
    fb = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2,)).astype(np.float32))
    eng = model._engine_for(torch.device(DEV))
    ranks, hits = pkg.evaluate_distance_fusions(eng, frame_desc, frame_product, shop_desc, target, (fw, fb),
                                                max_pairs=7 * G)      # forces several chunks
    wf = dict(weights)
    wf["last.weight"], wf["last.bias"] = fw, fb
    prob = so.match_scores(so.pair_logits(frame_desc, shop_desc, wf))
    for p in range(P):
        rows = (frame_product == p).nonzero().flatten()
        if len(rows) == 0:
            assert ranks[0, p] == G and ranks[1, p] == G
            continue
        for r, fused in enumerate((prob[rows].mean(0), prob[rows].max(0).values)):
            t = int(target[p])
            want = int(((fused > fused[t]) | ((fused == fused[t]) & (torch.arange(G) < t))).sum())
            near = int(((fused - fused[t]).abs() <= 1e-6).sum())          # ties inside fp32 rounding may swap
            assert abs(int(ranks[r, p]) - want) <= near, (p, r, int(ranks[r, p]), want)
    assert hits.shape == (2, 4)
    for r in range(2):
        assert hits[r].tolist() == [int((ranks[r] < k).sum()) for k in (1, 5, 10, 20)]


def test_distance_fusions_against_the_eval_loop(model, weights):
    """Larger case with many frames per product (several shared-memory frame chunks) and several gallery chunks,
    against the oracle's literal per-product loop (eval_product_loop: avg_dist / max_dist)."""
    P, G = 40, 2600
    rs = np.random.RandomState(16)
    lens = rs.randint(1, 40, size=P)
    lens[7] = 0
    frame_product = torch.tensor([p for p in range(P) for _ in range(int(lens[p]))])
    perm = torch.from_numpy(rs.permutation(len(frame_product)))
    frame_product = frame_product[perm]
    frame_desc = torch.from_numpy(rs.randn(len(frame_product), 256).astype(np.float32))
    shop_desc = torch.from_numpy(rs.randn(G, 256).astype(np.float32))
    target = torch.from_numpy(rs.permutation(G)[:P].astype(np.int64))
    fw = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2, 256)).astype(np.float32))
    fb = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2,)).astype(np.float32))
    eng = model._engine_for(torch.device(DEV))
    before = eng._last
    ranks, hits = pkg.evaluate_distance_fusions(eng, frame_desc, frame_product, shop_desc, target, (fw, fb))
    assert eng._last[0].data_ptr() == before[0].data_ptr()       # the caller's scorer is back
    wf = dict(weights)
    wf["last.weight"], wf["last.bias"] = fw, fb
    ref = so.eval_product_loop(frame_desc, frame_product, shop_desc, target, wf,
                               torch.zeros(P, 256), shop_desc, wf)
    prob = so.match_scores(so.pair_logits(frame_desc, shop_desc, wf))
    for r, key in enumerate(("avg_dist", "max_dist")):
        for p in range(P):
            rows = (frame_product == p).nonzero().flatten()
            if len(rows) == 0:
                assert int(ranks[r, p]) == G
                continue
            fused = prob[rows].mean(0) if r == 0 else prob[rows].max(0).values
            near = int(((fused - fused[int(target[p])]).abs() <= 1e-6).sum())
            assert abs(int(ranks[r, p]) - int(ref[key][p])) <= near, (key, p)


def test_self_distances(model, weights):
    """compute_selfdist (evaluate_movingfashion.py:115-121): street x street class-1 probabilities."""
    rs = np.random.RandomState(26)
    x = torch.from_numpy(rs.randn(77, 256).astype(np.float32))
    fw = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2, 256)).astype(np.float32))
    fb = torch.from_numpy(rs.uniform(-1 / 16, 1 / 16, (2,)).astype(np.float32))
    eng = model._engine_for(torch.device(DEV))
    got = pkg.self_distances(eng, x, (fw, fb)).cpu()
    wf = dict(weights)
    wf["last.weight"], wf["last.bias"] = fw, fb
    ref = so.match_scores(so.pair_logits(x, x, wf))
    assert got.shape == (77, 77) and (got - ref).abs().max() <= 1e-5
    # numpy-fp16 values of the script's own formula stay within fp16 resolution of the fp32 values
    f16 = so.eval_frame_scores_np(x.numpy().astype(np.float16), x.numpy().astype(np.float16), fw.numpy(), fb.numpy())
    assert np.abs(f16.astype(np.float32) - got.numpy()).max() <= 2e-2
    model._sync_weights(eng)
