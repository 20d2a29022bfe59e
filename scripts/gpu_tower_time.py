"""Conv tower timing (developer tool): python scripts/gpu_tower_time.py [K ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
dev = torch.device("cuda:0")
m = pkg.MatchPredictor().to(dev).eval()
eng = m._engine_for(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
FLOP_TRUE = 2 * (144 * 256 * 2304 + 100 * 256 * 2304 + 64 * 256 * 2304 + 36 * 1024 * 2304 + 1024 * 256)
for K in [int(a) for a in sys.argv[1:]] or [100, 1000, 3000]:
    x = torch.randn(K, 256, 14, 14, device=dev)
    for _ in range(3):
        m.embed(x)
    eng.profile(True)
    for _ in range(5):
        flush.fill_(1)
        m.embed(x)
    torch.cuda.synchronize()
    pr = eng.profile_read()["tower"]
    eng.profile(False)
    ms = pr[0] / pr[1]
    # the PyTorch modules (cuDNN, TF32 allowed as by default) for comparison
    with torch.no_grad():
        for _ in range(3):
            m.embed_torch(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            m.embed_torch(x)
        b.record()
        torch.cuda.synchronize()
    ms_t = a.elapsed_time(b) / 5
    print(f"tower K={K}: {ms * 1e3:.0f} us = {K * FLOP_TRUE / ms / 1e9:.0f} TFLOP/s of convolution arithmetic "
          f"({K / ms * 1e3:.0f} ROIs/s); PyTorch / cuDNN modules: {ms_t * 1e3:.0f} us", flush=True)
