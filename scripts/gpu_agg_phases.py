"""Per-phase cycle budget of the long-track aggregation kernel's producer iteration (SEAM_AGG_PHASES build):
python scripts/build_variants.py ph:SEAM_AGG_PHASES; SEAM_B200_LIB=.../libseam_b200.ph.so python scripts/gpu_agg_phases.py [Q T]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
Q, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100000, 64)
seq = torch.randn(1 + T, Q, 256, device=dev)
for _ in range(3):
    out, att = e.aggregate(seq, getatt=True)
torch.cuda.synchronize()
ph = att.view(-1)[:148 * 32].view(torch.int64).view(148, 16).cpu().double()
its = ph[:, 15].clamp(min=1)
names_w = ["loop top", "wait for the frames", "frames -> registers, dots, butterflies", "issue the next track's copy, peek", "interaction 1, max, e",
           "interaction 2 + z, {p,q} store", "weighted sums, |r| max", "wait for the batch slot", "publish (r split, pooled, fence, arrive)"]
names = ["loop top", "wait for the frames", "frames -> registers, dots, butterflies", "issue the next block's copy", "barrier 1 (scalars)",
         "interaction pass 1 + local softmax", "barrier 2", "p_t, barrier 4", "interaction pass 2, q sum, {p,q} store",
         "weighted sums, partials, |r| max", "barrier 5", "finishing: wait for the batch buffer", "finishing: sums, store, publish"]
if T <= 16:
    names = names_w
tot = 0.0
for i, n in enumerate(names):
    c = (ph[:, i] / its)
    tot += float(c.mean())
    print(f"{n:44s} mean {c.mean():8.0f}  min {c.min():8.0f}  max {c.max():8.0f} cycles per iteration")
print(f"{'total':44s} mean {tot:8.0f} cycles; iterations per warp {its.mean():.0f}")
