"""The fused peer exchange on ONE GPU (1-rank process group: every 'peer' is the local arena): lets ncu see the kernel.
python tools/r2_exch1.py K"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from distdiff_b200 import prototypes
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29578")
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
coll = prototypes.PeerCollective()
N, D, C = 100_000, 2048, 100
g = torch.Generator(device=dev).manual_seed(7)
feats = torch.randn(N, D, generator=g, device=dev); labels = torch.arange(N, device=dev) % C
for it in (1, 6):
    prototypes.build_prototypes(feats, labels, C, K, "kmeans", it, coll=coll)
torch.cuda.synchronize()
print("phases_us", coll.arena.timing())
coll.close(); dist.destroy_process_group()
