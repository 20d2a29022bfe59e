"""Host-side logic that needs no GPU: partitioning, module surface (constructor, parameter
names = the reference's state_dict keys), track grouping, and the world_size-2 gloo path of
the sharded retriever with the oracle standing in for the device kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import seam_match_rcnn_b200 as pkg
from oracle import seam_oracle as so

# TemporalAggregationNLB().state_dict() keys of the reference (probed)
REFERENCE_KEYS = {
    "conv_seq.0.weight": (256, 256, 3, 3), "conv_seq.0.bias": (256,),
    "conv_seq.2.weight": (256, 256, 3, 3), "conv_seq.2.bias": (256,),
    "conv_seq.4.weight": (256, 256, 3, 3), "conv_seq.4.bias": (256,),
    "conv_seq.6.weight": (1024, 256, 3, 3), "conv_seq.6.bias": (1024,),
    "linear.0.weight": (256, 1024), "linear.0.bias": (256,),
    "linear.1.weight": (256,), "linear.1.bias": (256,),
    "linear.1.running_mean": (256,), "linear.1.running_var": (256,), "linear.1.num_batches_tracked": (),
    "last.weight": (2, 256), "last.bias": (2,),
    "attention_scorer.weight": (1, 256), "attention_scorer.bias": (1,),
    "newnlb.g.weight": (128, 256, 1), "newnlb.g.bias": (128,),
    "newnlb.W.weight": (256, 128, 1), "newnlb.W.bias": (256,),
    "newnlb.theta.weight": (128, 256, 1), "newnlb.theta.bias": (128,),
    "newnlb.phi.weight": (128, 256, 1), "newnlb.phi.bias": (128,),
    "newnlb.concat_project.0.weight": (1, 256, 1, 1),
}


def test_state_dict_keys_match_reference():
    sd = pkg.TemporalAggregationNLB().state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == REFERENCE_KEYS
    mp_keys = {k for k in REFERENCE_KEYS if not k.startswith(("attention_scorer", "newnlb"))}
    assert set(pkg.MatchPredictor().state_dict()) == mp_keys
    m = pkg.TemporalAggregationNLB()
    assert m.nlb is True and m.n_frames == -1                     # models/match_head.py:85-88
    assert float(m.newnlb.W.weight.abs().sum()) == 0.0            # models/nlb.py:48-49
    assert set(so.HOT_KEYS) <= set(sd)


def test_state_dict_loads_a_reference_style_checkpoint():
    """Checkpoints are {'model_state_dict': ...} with the aggregator under
    roi_heads.temporal_aggregator. (evaluate_movingfashion.py:502-503, models/video_matchrcnn.py:37)."""
    src = pkg.TemporalAggregationNLB()
    ckpt = {"model_state_dict": {"roi_heads.temporal_aggregator." + k: v for k, v in src.state_dict().items()}}
    dst = pkg.TemporalAggregationNLB()
    prefix = "roi_heads.temporal_aggregator."
    dst.load_state_dict({k[len(prefix):]: v for k, v in ckpt["model_state_dict"].items() if k.startswith(prefix)})
    for k, v in src.state_dict().items():
        assert torch.equal(v, dst.state_dict()[k])


def test_shard_bounds():
    for n in (0, 1, 7, 8, 1000, 15000, 1000000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for r in range(world):
                lo, hi = pkg.shard_bounds(n, world, r)
                assert lo == prev and hi >= lo
                prev = hi
                sizes.append(hi - lo)
            assert prev == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pkg.shard_bounds(10, 2, 2)


def test_group_tracks_matches_reference_grouping():
    """Device-side grouping == the reference's per-id loop (models/match_head.py:97-111)."""
    rs = np.random.RandomState(0)
    ids = torch.from_numpy(rs.randint(0, 6, size=40)) * 3
    x = torch.from_numpy(rs.randn(40, 256).astype(np.float32))
    seq, mask, counts = pkg.TemporalAggregationNLB._group_tracks(x, ids)
    maxlen = int((ids == ids.mode()[0]).sum())
    uniq = ids.unique()
    ref_seq = torch.zeros((1 + maxlen, uniq.numel(), 256))
    ref_mask = torch.zeros((uniq.numel(), 1 + maxlen), dtype=torch.bool)
    for i, idd in enumerate(uniq):
        n = int((ids == idd).sum())
        ref_seq[1:n + 1, i] = x[ids == idd]
        ref_mask[i, n + 1:] = True
    assert torch.equal(seq, ref_seq) and torch.equal(mask, ref_mask)


# ------------------------------------------------------------------ gloo, world_size = 2
class OracleOps:
    """CPU stand-in for SeamEngine (TESTS ONLY) so the collective logic runs under gloo."""

    def __init__(self, w):
        self.w = w

    def aggregate(self, seq, mask=None, lens=None, getatt=False):
        if mask is None:
            mask = torch.zeros(seq.shape[1], seq.shape[0], dtype=torch.bool)
        return so.aggregate_tracks(seq, mask, self.w)[0] if seq.shape[1] else torch.zeros(0, 256)

    def prepare_gallery(self, g, index_offset=0):
        return (g, index_offset)

    def score_topk(self, q, gal, k):
        g, off = gal
        s, d, i = so.rank_topk(so.pair_logits(q, g, self.w), k)
        pad = k - s.shape[1]
        if pad:
            s = torch.cat([s, torch.zeros(s.shape[0], pad)], 1)
            d = torch.cat([d, torch.full((d.shape[0], pad), -float("inf"))], 1)
            i = torch.cat([i + off, torch.full((i.shape[0], pad), -1, dtype=torch.long)], 1)
        else:
            i = i + off
        return s, d, i.int()

    def merge_topk(self, S, M, I):
        k = S.shape[2]
        s, d, i = so.merge_topk(list(S), list(M), [x.long() for x in I], k)
        return s, d, i.int()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, Q, T, G, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        w = so.random_weights(0)
        seq, mask, _ = so.synth_tracks(Q, T, seed=13, ragged=(1, T))
        gal = so.synth_gallery(G, 13, None)
        r = pkg.ShardedRetriever.from_full_gallery(OracleOps(w), gal)
        lo, hi = pkg.shard_bounds(G, world, rank)
        assert r.gallery[0].shape[0] == hi - lo and r.gallery[1] == lo
        s, d, i = r.search(seq, mask, k)
        torch.save((s, d, i), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("Q,G", [(9, 101), (4, 3)])
def test_sharded_search_gloo_world2(tmp_path, Q, G):
    T, k, world = 4, 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, Q, T, G, k, str(tmp_path)), nprocs=world, join=True)
    w = so.random_weights(0)
    seq, mask, _ = so.synth_tracks(Q, T, seed=13, ragged=(1, T))
    gal = so.synth_gallery(G, 13, None)
    q, _ = so.aggregate_tracks(seq, mask, w)
    s_ref, d_ref, i_ref = so.rank_topk(so.pair_logits(q, gal, w), k)
    kk = s_ref.shape[1]
    for rank in range(world):
        s, d, i = torch.load(os.path.join(str(tmp_path), f"r{rank}.pt"))
        assert torch.equal(i[:, :kk].long(), i_ref) and (i[:, kk:] == -1).all()
        assert torch.allclose(d[:, :kk], d_ref, atol=1e-6) and torch.allclose(s[:, :kk], s_ref, atol=1e-6)


# ----------------------------------------------------------------------------------------
# scorer work decomposition (pure host logic inside the C library: no device needed)
# ----------------------------------------------------------------------------------------
def _partition(num_sms, Q, G, rank_variant=0):
    import ctypes as C
    lib = pkg.load_library()
    bounds = (C.c_int32 * (num_sms + 1))()
    out = (C.c_int32 * 6)()
    assert lib.seam_score_partition(num_sms, Q, G, rank_variant, bounds, num_sms + 1, out) == 0
    mt, nt, grid, P, cap, nseed = list(out)
    return list(bounds)[:grid + 1], mt, nt, grid, P, cap, nseed


@pytest.mark.parametrize("Q,G", [(64, 1000), (15000, 15000), (10000, 50000), (10000, 6250), (10000, 125000),
                                 (1, 1), (130, 257), (1000, 1000000), (200000, 300)])
@pytest.mark.parametrize("rank_variant", [0, 1])
def test_scorer_partition_is_a_balanced_cover(Q, G, rank_variant):
    """CTA b sweeps tiles [tb[b], tb[b+1]): the ranges must tile the (query tile x gallery tile) grid
    exactly, P must cover the CTAs that share any row, and the cost model (tiles + sample tiles of the
    segments that hold a row's head or are swept first + a constant per segment) must be balanced."""
    num_sms = 148
    tb, mt, nt, grid, P, cap, nseed = _partition(num_sms, Q, G, rank_variant)
    assert mt == -(-Q // 128) and nt == -(-G // 256)
    total = mt * nt
    assert grid == min(total, num_sms)
    assert tb[0] == 0 and tb[-1] == total and all(a <= b for a, b in zip(tb, tb[1:]))
    assert nseed == (0 if rank_variant else 4)
    assert cap >= 128 and cap & (cap - 1) == 0
    # pieces per row
    first = {}
    pieces = 1
    for b in range(grid):
        lo, hi = tb[b], tb[b + 1]
        if lo == hi:
            continue
        for m in range(lo // nt, (hi - 1) // nt + 1):
            first.setdefault(m, b)
            pieces = max(pieces, b - first[m] + 1)
    assert len(first) == mt and pieces <= P

    def cost(lo, hi):
        c, t, swept_first = 0.0, hi, True
        while t > lo:                                   # segments, last first (score_tc.cuh: segment_before)
            m = (t - 1) // nt
            start = max(lo, m * nt)
            n = t - start
            c += 0.35 + n + (min(nseed, n) if (swept_first or start == m * nt) else 0)
            swept_first, t = False, start
        return c
    costs = [cost(tb[b], tb[b + 1]) for b in range(grid) if tb[b + 1] > tb[b]]
    if total >= 4 * num_sms:
        # the last CTA takes the remainder (possibly less); nobody is far above the mean
        assert max(costs) <= 1.15 * (sum(costs) / len(costs)) + 2.0


def test_modules_copy_and_pickle_without_the_engine():
    """copy.deepcopy / pickle of a module that has run once (and so caches a ctypes-backed engine) work: the
    engine is transient state, the copy re-creates its own on first use."""
    import copy
    import pickle
    import seam_match_rcnn_b200 as pkg
    m = pkg.TemporalAggregationNLB()
    m.__dict__["_seam_engine"] = object()          # stands in for a SeamEngine (ctypes pointers are not picklable)
    m.__dict__["_seam_key"] = ("k",)
    c = copy.deepcopy(m)
    assert "_seam_engine" not in c.__dict__ and "_seam_key" not in c.__dict__
    r = pickle.loads(pickle.dumps(m))
    assert "_seam_engine" not in r.__dict__
    assert sorted(r.state_dict()) == sorted(m.state_dict())
    assert "_seam_engine" in m.__dict__            # the original keeps its engine


def test_match_predictor_embedding_keeps_its_graph():
    """The conv tower runs under autograd as in the reference (models/match_head.py:67-69): only the kernel
    outputs are detached."""
    import seam_match_rcnn_b200 as pkg
    m = pkg.MatchPredictor()
    x = torch.randn(2, 256, 14, 14)
    assert m.embed(x).requires_grad
