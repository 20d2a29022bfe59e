"""Per-CTA timeline of the fused aggregation kernel (SEAM_AGG_TIMELINE build):
SEAM_B200_LIB=.../libseam_b200.tl.so python scripts/gpu_agg_timeline.py [Q T]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
Q, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (15000, 10)
seq = torch.randn(1 + T, Q, 256, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
clean = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
CLEAN = os.environ.get("SEAM_FLUSH_CLEAN", "1") != "0"   # read 256 MiB after the write: L2 holds clean lines, no write-backs during the kernel
for _ in range(3):
    e.aggregate(seq, getatt=True)
flush.fill_(1)
if CLEAN: clean.sum()
out, att = e.aggregate(seq, getatt=True)
torch.cuda.synchronize()
tl = att.view(-1)[:148 * 16].view(torch.int64).view(148, 8).cpu().double()
t0 = tl[:, 0].min()
if os.environ.get('SEAM_TL3'):
    tl = tl - tl[:, :1] + t0          # per CTA, relative to its own first stamp
names2 = ["entry", "setup done", "registers re-balanced", "first frames requested", "M share loaded", "track 0 landed", "track 1 landed", "track 2 landed"]
names3 = ["batch 3 complete (tile_full)", "batch 3 MMAs issued", "batch 3 accumulator ready", "batch 3 accumulator read",
          "batch 3 written", "batch 4 complete (tile_full)", "producer enters claim", "producer has its slot"]
names = names3 if os.environ.get("SEAM_TL3") else names2 if os.environ.get("SEAM_TL2") else ["entry", "M share loaded (warp 0)", "first frames landed (warp 0)", "M in tensor memory (MMA lane)", "first MMA batch issued",
         "last MMA batch issued", "last batch written", "exit"]
rel = (tl - t0) / 1e3
for i, n in enumerate(names):
    c = rel[:, i]
    print(f"{n:36s} min {c.min():7.2f}  median {c.median():7.2f}  max {c.max():7.2f} us")
