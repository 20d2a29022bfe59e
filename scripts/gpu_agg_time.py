"""Aggregation kernel timing over track counts / lengths (developer tool): python scripts/gpu_agg_time.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
clean = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
CLEAN = os.environ.get("SEAM_FLUSH_CLEAN", "1") != "0"   # read 256 MiB after the write: L2 holds clean lines, no write-backs during the kernel
HBM = 6542.1
CASES = [(15000, 10), (100000, 10), (100000, 4), (10000, 4), (100000, 16), (50000, 32), (100000, 64), (64, 10)]
if len(sys.argv) > 2:
    CASES = [(int(q), int(sys.argv[-1])) for q in sys.argv[1:-1]]
for Q, T in CASES:
    seq = torch.randn(1 + T, Q, 256, device=dev)
    for _ in range(3):
        e.aggregate(seq)
    e.profile(True)
    for _ in range(10):
        flush.fill_(1)
        if CLEAN: clean.sum()
        e.aggregate(seq)
    torch.cuda.synchronize()
    pr = e.profile_read()
    e.profile(False)
    us = pr["aggregate"][0] / pr["aggregate"][1] * 1e3
    by = Q * (T + 1) * 1024
    print(f"aggregate Q={Q} T={T}: {us:.1f} us = {by / us / 1e3:.0f} GB/s ({by / us / 1e3 / HBM * 100:.0f} %)", flush=True)
    del seq
