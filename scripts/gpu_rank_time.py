import os, sys, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
dev = torch.device("cuda:0")
e = pkg.SeamEngine(dev); e.load_weights(random_init_weights(dev))
for Q, G in [(15000, 15000), (10000, 125000)]:
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(Q, 256, device=dev, generator=g)
    gal_t = torch.randn(G, 256, device=dev, generator=g)
    n = min(Q, G)
    gal_t[:n] = q[:n] + 0.1 * torch.randn(n, 256, device=dev, generator=g)
    gal = e.prepare_gallery(gal_t)
    for name, target in (("planted (true match)", torch.arange(Q, device=dev) % G), ("random item", torch.randint(0, G, (Q,), device=dev, generator=g))):
        for _ in range(2): r, m, st = e.rank_of_target(q, gal, target, return_stats=True)
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): r, m, st = e.rank_of_target(q, gal, target, return_stats=True)
        b.record(); torch.cuda.synchronize()
        t_fast = a.elapsed_time(b) / 5
        msg = f"Q={Q} G={G} target={name}: tensor-core path {t_fast*1e3:.0f} us, uncertified rows {int(st[0])}, median rank {int(r.float().median())}"
        if Q * G <= 3e8:
            a.record(); r2, m2 = e.rank_of_target(q, gal_t, target); b.record(); torch.cuda.synchronize()
            msg += f", exhaustive fp32 kernel {a.elapsed_time(b)*1e3:.0f} us, identical={bool(torch.equal(r, r2))}"
        print(msg, flush=True)
