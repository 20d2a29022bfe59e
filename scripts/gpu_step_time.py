"""One-GPU step time, graph-replayed: seam_search (fused query preparation) against aggregate + score_topk
(developer tool): python scripts/gpu_step_time.py [Q T G]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
Q, T, G = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (15000, 10, 15000))]
k = 20
gen = torch.Generator(device=dev).manual_seed(1)
seq = torch.zeros(1 + T, Q, 256, device=dev); seq[1:] = torch.randn(T, Q, 256, device=dev, generator=gen)
g = torch.randn(G, 256, device=dev, generator=gen)
n = min(Q, G)
g[:n] = e.aggregate(seq)[:n] + 0.1 * torch.randn(n, 256, device=dev, generator=gen)
gal = e.prepare_gallery(g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def two_calls():
    q = e.aggregate(seq)
    return (q,) + tuple(e.score_topk(q, gal, k))
def one_call():
    return e.search(seq, None, gal, k)
res = {}
for name, fn in (("aggregate + score_topk", two_calls), ("seam_search", one_call)):
    side = torch.cuda.Stream(device=dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        out = fn()
    tot = 0.0
    for _ in range(20):
        flush.fill_(1); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gph.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    res[name] = [t.clone() for t in out]
    print(f"{name}: {tot / 20 * 1e3:.1f} us per step (Q={Q} T={T} G={G})", flush=True)
a, b = res["aggregate + score_topk"], res["seam_search"]
print("identical:", all(torch.equal(x, y) for x, y in zip(a, b)))
