"""Per-kernel times of the sharded search under torchrun (developer tool):
   torchrun --nproc-per-node N scripts/multi_gpu_time.py"""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seam_match_rcnn_b200 as pkg
from bench import random_init_weights
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = pkg.SeamEngine(dev); eng.load_weights(random_init_weights(dev))
Q, T, Gs, k = 15000, 10, 15000, 20
seq = torch.randn(1 + T, Q, 256, device=dev)
gal = eng.prepare_gallery(torch.randn(Gs, 256, device=dev), index_offset=rank * Gs)
peer = pkg.PeerExchange(eng, Q, k)
lo, hi = peer.q_lo[rank], peer.q_lo[rank + 1]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def ev():
    return torch.cuda.Event(enable_timing=True)
names = ["local aggregate (own tracks)", "sharded aggregate", "sharded score_topk (prep+K2+K3+exact)", "sharded merge"]
tot = [0.0] * 4
N = 10
for it in range(N + 3):
    flush.fill_(1)
    dist.barrier(); torch.cuda.synchronize()
    e = [ev() for _ in range(6)]
    e[0].record(); eng.aggregate(seq[:, lo:hi]); e[1].record()
    torch.cuda.synchronize(); dist.barrier(); flush.fill_(1); torch.cuda.synchronize(); dist.barrier()
    e[2].record(); eng.sharded_aggregate(peer, seq[:, lo:hi]); e[3].record()
    eng.sharded_score_topk(peer, gal); e[4].record()
    eng.sharded_merge(peer); e[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i, (a, b) in enumerate(((0, 1), (2, 3), (3, 4), (4, 5))):
            tot[i] += e[a].elapsed_time(e[b])
t = torch.tensor(tot, device=dev, dtype=torch.float64) / N
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world={world}: " + "; ".join(f"{n} {float(v) * 1e3:.1f} us" for n, v in zip(names, t)))
dist.barrier(); dist.destroy_process_group()
