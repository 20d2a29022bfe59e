"""Aggregation kernel against the oracle on small shapes with a capped grid (several tracks per warp,
several batches per CTA).  Developer tool: SEAM_DEBUG_AGG_GRID=1 python scripts/gpu_agg_debug.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seam_match_rcnn_b200 as pkg
from oracle import seam_oracle as so
dev = torch.device("cuda:0")
w = so.random_weights(0)
e = pkg.SeamEngine(dev); e.load_weights({k: v.to(dev) for k, v in w.items()})
cases = [(12, 10, None), (16, 10, None), (40, 10, None), (100, 10, (0, 10)), (16, 4, None), (37, 4, (0, 4)), (200, 4, (0, 4)),
         (60, 16, (1, 16)), (9, 32, (1, 32)), (40, 64, (1, 64))]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for Q, T, rag in cases:
    seq, mask, lens = so.synth_tracks(Q, T, seed=Q + T, ragged=rag)
    ref, _ = so.aggregate_tracks(seq, mask, w)
    try:
        out = e.aggregate(seq.to(dev), mask.to(dev))
        torch.cuda.synchronize()
        err = (out.cpu() - ref).abs().max().item()
        print(f"Q={Q} T={T} ragged={rag}: max err {err:.3e}", flush=True)
    except Exception as ex:                      # noqa: BLE001
        print(f"Q={Q} T={T} ragged={rag}: FAILED {type(ex).__name__}: {str(ex).splitlines()[0]}", flush=True)
        for r in e.watchdog_records():
            print("   watchdog", r, flush=True)
        break
