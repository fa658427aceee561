"""The in-step K5 launch of bench.py's default workload in isolation (B=8 latents, fp16 storage, CFG), for the
`ncu --set full` capture that feeds `roofline.traffic`:  L2 is evicted before every launch, like in a real step where
~3.4 GB of UNet traffic separates two K5 launches.
    ncu --set full --clock-control none -k regex:cfg_ddim_fwd --launch-skip 5 -c 1 -o out python tools/instep_k5.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distdiff_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
npred = torch.randn(2 * B, 4, 64, 64, generator=g, device=dev).half()
x = torch.randn(B, 4, 64, 64, generator=g, device=dev).half()
junk = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(10):
    junk.add_(1)                      # 512 MB read + write: evicts the 126 MB L2
    ops.cfg_ddim_step(npred, x, 7.5, 0.3, 0.35)
torch.cuda.synchronize()
print("ok")
