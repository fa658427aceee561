"""K4 at B = 65536: random vs class-sorted targets (is the row gather/scatter pattern the limiter?), plus a torch row gather."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distdiff_b200 import ops
from distdiff_b200.microbench import timeit, hbm_peak_gbs
dev = torch.device("cuda:0"); peak, _ = hbm_peak_gbs()
C, D, B = 100, 2048, 65536
f = torch.randn(B, D, device=dev)
for K in (3, 10):
    g = torch.nn.functional.normalize(torch.randn(C, D, device=dev), dim=-1)
    l = torch.nn.functional.normalize(torch.randn(C, K, D, device=dev), dim=-1)
    yr = torch.randint(0, C, (B,), device=dev); ys = yr.sort().values
    nb = 2 * B * D * 4 + (K + 1) * C * D * 4
    for name, y in (("random", yr), ("sorted", ys)):
        for mode in ("tile_pair", "tile_cta"):
            t = timeit(lambda: ops.energy_fwd_bwd(f, y, g, l, 1.0, 1.0, True, mode=mode), 10)
            print(json.dumps({"K": K, "targets": name, "mode": mode, "ms": round(t * 1e3, 4), "frac": round(nb / t / 1e9 / peak, 3)}), flush=True)
perm = torch.randperm(B, device=dev)
out = torch.empty_like(f)
t = timeit(lambda: torch.index_select(f, 0, perm, out=out), 10)
print(json.dumps({"torch_row_gather_ms": round(t * 1e3, 4), "frac": round(2 * B * D * 4 / t / 1e9 / peak, 3)}))
t = timeit(lambda: out.copy_(f), 10)
print(json.dumps({"torch_copy_ms": round(t * 1e3, 4), "frac": round(2 * B * D * 4 / t / 1e9 / peak, 3)}))
