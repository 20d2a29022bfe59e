import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
from oracle import seam_oracle as so
dev = torch.device("cuda:0")
W = so.random_weights(0)
e = pkg.SeamEngine(dev); e.load_weights({k: v.to(dev) for k, v in W.items()})
Q, T, G = 40, 6, 120
seq, mask, lens = so.synth_tracks(Q, T, seed=2, ragged=(1, 6))
qref, _ = so.aggregate_tracks(seq, mask, W)
gal = so.synth_gallery(G, 2, None)
target = torch.arange(Q) * 2
q = e.aggregate(seq.to(dev), mask.to(dev))
print("lens", lens)
print("|q-qref| per row", (q.cpu() - qref).abs().max(1).values)
x5 = so.pair_logits(qref, gal, W)
ranks = so.rank_of_target(x5, target)
r1, m1 = e.rank_of_target(qref.to(dev), gal.to(dev), target)
r2, m2 = e.rank_of_target(q, gal.to(dev), target)
print("oracle", ranks.tolist()); print("gpu(qref)", r1.cpu().tolist()); print("gpu(q)", r2.cpu().tolist())
d = so.logit_margin(x5); dt = d.gather(1, target.view(-1, 1))[:, 0]
print("margin err (qref)", (m1.cpu() - dt).abs().max().item())
