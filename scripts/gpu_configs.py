"""Device-timed stage times for the BASELINE.json configurations (per-GPU shapes), L2 flushed between
iterations.  Developer / documentation tool: python scripts/gpu_configs.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
HBM, TF = 6542.1, 1632.4
CFG = [("cfg1 64x10 vs 1,000", 64, 10, None, 1000),
       ("cfg2 15,000x10 vs 15,000", 15000, 10, None, 15000),
       ("cfg3 10,000x(2..4) vs 50,000", 10000, 4, (2, 4), 50000),
       ("cfg3 per GPU at N=8: vs 6,250", 10000, 4, (2, 4), 6250),
       ("cfg4 100,000x64 (aggregation only)", 100000, 64, None, 0),
       ("cfg5 per GPU at N=8: 10,000x10 vs 125,000", 10000, 10, None, 125000)]
for name, Q, T, rag, G in CFG:
    g = torch.Generator(device=dev).manual_seed(1)
    seq = torch.zeros(1 + T, Q, 256, device=dev); seq[1:] = torch.randn(T, Q, 256, device=dev, generator=g)
    lens = None
    if rag:
        lens = torch.randint(rag[0], rag[1] + 1, (Q,), device=dev, generator=g).int()
    nfr = int(lens.sum()) if lens is not None else Q * T
    q = e.aggregate(seq, None, lens=lens)
    gal = e.prepare_gallery(torch.randn(G, 256, device=dev, generator=g)) if G else None
    for _ in range(3):
        e.aggregate(seq, None, lens=lens)
        if G: e.score_topk(q, gal, 20)
    e.profile(True)
    for _ in range(10):
        flush.fill_(1)
        e.aggregate(seq, None, lens=lens)
        if G: e.score_topk(q, gal, 20)
    torch.cuda.synchronize()
    pr = {k: v[0] / v[1] * 1e3 for k, v in e.profile_read().items() if v[1]}
    e.profile(False)
    agg_bytes = (nfr + Q) * 1024
    msg = f"{name}: aggregate {pr['aggregate']:.1f} us = {agg_bytes / pr['aggregate'] / 1e3:.0f} GB/s ({agg_bytes / pr['aggregate'] / 1e3 / HBM * 100:.0f} %)"
    if G:
        tfl = Q * G * 512 / pr['score'] / 1e6
        tot = sum(pr.values())
        msg += (f", prep {pr['prep_queries']:.1f}, score {pr['score']:.1f} us = {tfl:.0f} TFLOP/s ({tfl / TF * 100:.0f} %), "
                f"rescore {pr['rescore']:.1f}, exact {pr['exact']:.1f}; all {tot:.0f} us = {Q * G / tot / 1e6:.2f} Tpairs/s, {Q / tot:.1f} Mqueries/s")
    print(msg, flush=True)
