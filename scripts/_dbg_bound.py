"""A/B of two builds of the library on single-CTA sweeps (no cross-CTA exchange)."""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
which = sys.argv[1]
import seam_match_rcnn_b200 as pkg
from seam_match_rcnn_b200 import _lib
lib = ctypes.CDLL(os.path.join(ROOT, "scripts", "_libs", which + ".so"))
_lib._declare(lib); _lib._lib = lib
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
for (Q, G) in [(128, 47 * 256), (128, 200 * 256)]:
    torch.manual_seed(0)
    q = torch.randn(Q, 256, device=dev); g = torch.randn(G, 256, device=dev)
    gal = e.prepare_gallery(g)
    e.score_topk(q, gal, 20); torch.cuda.synchronize()
    plan = e.score_plan(Q, G)
    ws = e._ws["score"]; nl = plan["ctas_per_query_tile"] * 4
    cn = ws[plan["off_rowcnt"]:plan["off_rowcnt"] + Q * nl * 4].view(torch.int32).view(Q, nl).float()
    print(which, "grid", os.environ.get("SEAM_DEBUG_SCORE_GRID"), f"Q={Q} G={G} ctas={plan['ctas']} P={plan['ctas_per_query_tile']} records/row mean={cn.sum(1).mean():.1f} max={int(cn.sum(1).max())}")
