"""The gallery-sharded search with the exchange done by the kernels (csrc/exchange.cuh) on ONE GPU: with a single
rank, and with two "ranks" living in one process on two streams (their buffers are ordinary device memory instead
of NVLink-mapped peer memory -- the kernels, flags, parity double-buffering and the step counter are the real
ones).  The real multi-GPU run is checked by scripts/multi_gpu_check.py under torchrun."""
import pytest
import torch

import seam_match_rcnn_b200 as pkg
from oracle import seam_oracle as so
from util import TOL_EMB, assert_topk_matches

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(step, Q, T, G, ragged):
    seq, mask, lens = so.synth_tracks(Q, T, seed=100 + step, ragged=ragged)
    return seq.to(DEV), mask.to(DEV)


def test_sharded_search_single_rank(weights, engine):
    """world = 1: the sharded entry points reduce to aggregate + score_topk (signals to itself, waits on its own
    flags); three steps with different queries exercise both parity halves and the step counter."""
    Q, T, G, k = 700, 10, 3000, 20
    gal = torch.randn(G, 256, generator=torch.Generator().manual_seed(5)).to(DEV)
    peer = pkg.PeerExchange(engine, Q, k)
    retr = pkg.ShardedRetriever(engine, gal, 0)
    for step in range(3):
        seq, mask = _inputs(step, Q, T, G, (1, 10))
        sc, mg, ix = retr.search_peer(seq, mask, peer)
        ref = pkg.search(engine, seq, mask, gal, k)
        assert torch.equal(ix, ref[2]) and torch.equal(mg, ref[1]) and torch.equal(sc, ref[0])
        assert torch.equal(peer.descriptors()[(step + 1) & 1], engine.aggregate(seq, mask))
    assert peer.steps_done == 3


@pytest.mark.parametrize("replicate", [True, False])
def test_sharded_search_two_ranks_one_gpu(weights, engine, replicate):
    """Two ranks in one process: each owns half of the queries and half of the gallery, runs on its own stream with
    its own engine, and the kernels exchange descriptors / lists / merged rows through the flag protocol.  Ragged
    query count (odd Q), four steps, results equal the single-GPU search over the whole gallery."""
    Q, T, G, k = 301, 4, 2500, 20
    w = {kk: v.to(DEV) for kk, v in weights.items()}
    engines = [engine, pkg.SeamEngine(DEV)]
    engines[1].load_weights(w)
    gal = torch.randn(G, 256, generator=torch.Generator().manual_seed(6)).to(DEV)
    shared = {}
    peers = [pkg.PeerExchange(engines[r], Q, k, replicate=replicate,
                              local_peers={"world": 2, "rank": r, "shared": shared}) for r in range(2)]
    for p in peers:
        p._fill_struct()
    bounds = [pkg.shard_bounds(G, 2, r) for r in range(2)]
    retrs = [pkg.ShardedRetriever(engines[r], gal[bounds[r][0]:bounds[r][1]], bounds[r][0], world=2, rank=r)
             for r in range(2)]
    streams = [torch.cuda.Stream(DEV) for _ in range(2)]
    for step in range(4):
        seq, mask = _inputs(10 + step, Q, T, G, (0, 4))
        ref = pkg.search(engine, seq, mask, gal, k)
        torch.cuda.synchronize()
        res = [None, None]
        for r in (1, 0) if step & 1 else (0, 1):             # either rank may be first to enqueue
            streams[r].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(streams[r]):
                res[r] = retrs[r].search_peer(seq, mask, peers[r])
        torch.cuda.synchronize()
        assert engines[0].watchdog_records() == []
        for r in range(2):
            lo, hi = peers[r].q_lo[r], peers[r].q_lo[r + 1]
            sc, mg, ix = res[r]
            if replicate:
                assert torch.equal(ix, ref[2]) and torch.equal(mg, ref[1]) and torch.equal(sc, ref[0])
            else:
                assert torch.equal(ix, ref[2][lo:hi]) and torch.equal(mg, ref[1][lo:hi]) and torch.equal(sc, ref[0][lo:hi])
    assert peers[0].steps_done == 4 and peers[1].steps_done == 4


def test_sharded_aggregate_in_chunks(weights, engine):
    """A rank may hand its tracks over in several calls (host streaming): rows land at row0 of the descriptor
    buffer, only the last call signals."""
    Q, T, G, k = 256, 10, 1500, 10
    gal = torch.randn(G, 256, generator=torch.Generator().manual_seed(7)).to(DEV)
    peer = pkg.PeerExchange(engine, Q, k)
    g = engine.prepare_gallery(gal)
    seq, mask = _inputs(3, Q, T, G, None)
    for lo, hi in ((0, 100), (100, 101), (101, 256)):
        engine.sharded_aggregate(peer, seq[:, lo:hi], mask[lo:hi], row0=lo, last=hi == Q)
    engine.sharded_score_topk(peer, g)
    sc, mg, ix = engine.sharded_merge(peer)
    ref = pkg.search(engine, seq, mask, gal, k)
    assert torch.equal(ix, ref[2]) and torch.equal(mg, ref[1])
