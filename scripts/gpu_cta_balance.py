import os, sys, torch
os.environ["SEAM_DEBUG_CTA_NS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
Q, G = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (15000, 15000)
torch.manual_seed(0)
q = torch.randn(Q, 256, device=dev); g = torch.randn(G, 256, device=dev)
gal = e.prepare_gallery(g)
for _ in range(3): e.score_topk(q, gal, 20)
torch.cuda.synchronize()
plan = e.score_plan(Q, G); ws = e._ws["score"]
off = plan["off_rows"] + ((Q * 4 - 4096) & ~7)
nc = plan['ctas']
t = ws[off:off + nc * 16].view(torch.int64).view(nc, 2).cpu()
ns, seg = t[:, 0].float() / 1e3, t[:, 1]
print("per-CTA us: min %.1f mean %.1f max %.1f" % (ns.min(), ns.mean(), ns.max()))
for sgc in sorted(set(seg.tolist())):
    m = seg == sgc
    print(f"  segments={sgc}: n={int(m.sum())} mean={ns[m].mean():.1f} min={ns[m].min():.1f} max={ns[m].max():.1f}")
srt = torch.argsort(ns, descending=True)[:12]
print("slowest CTAs:", [(int(i), round(float(ns[i]), 1), int(seg[i])) for i in srt])
print("fastest CTAs:", [(int(i), round(float(ns[i]), 1), int(seg[i])) for i in torch.argsort(ns)[:8]])
