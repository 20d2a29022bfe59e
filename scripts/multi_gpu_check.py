"""N-GPU gallery-sharded search against the single-GPU search of the whole gallery (NCCL).
   torchrun --nproc-per-node N scripts/multi_gpu_check.py"""
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
Q, T, G, k = 3000, 10, 40000 + 13, 20                      # G not divisible by the world size
g = torch.Generator(device="cpu").manual_seed(7)
seq = torch.zeros(1 + T, Q, 256); seq[1:] = torch.randn(T, Q, 256, generator=g)
gal = torch.randn(G, 256, generator=g)
seq, gal = seq.to(dev), gal.to(dev)
r = pkg.ShardedRetriever.from_full_gallery(eng, gal)
sc, mg, ix = r.search(seq, None, k)
ok = True
if rank == 0:
    sc1, mg1, ix1 = pkg.search(eng, seq, None, gal, k)
    same = torch.equal(ix, ix1)
    derr = (mg - mg1).abs().max().item()
    print(f"world={world}: sharded vs single-GPU top-{k}: indices identical={same} |margin diff|max={derr:.2e}")
    ok = same and derr <= 3e-5
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
