# round 2: K4 tile kernel, rows-per-batch variants (tools/build_variant.sh) + ncu of the product build.  TAG=... bash tools/r2_k4b.sh
TAG=${TAG:-r2ad}
set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "energy" --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
for v in product $VARIANTS; do
  lib=""; [ "$v" != "product" ] && lib=$PWD/distdiff_b200/_variants/$v.so
  DD_LIB_PATH=$lib KBENCH_KS=${KS:-3,5,10} timeout 600 python tools/kbench.py K4 > gpurun_out/${TAG}_kbench_$v.jsonl 2> gpurun_out/${TAG}_kbench_$v.err
  echo "== $v"; grep -E "B(4096|65536)_norm1_tile|B1024_norm1" gpurun_out/${TAG}_kbench_$v.jsonl | cut -c1-110
done
cat > /tmp/k4one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from distdiff_b200 import ops
K = int(sys.argv[1]); dev = torch.device('cuda:0')
C, D, B = 100, 2048, 65536
g = torch.nn.functional.normalize(torch.randn(C, D, device=dev), dim=-1)
l = torch.nn.functional.normalize(torch.randn(C, K, D, device=dev), dim=-1)
f = torch.randn(B, D, device=dev); y = torch.randint(0, C, (B,), device=dev)
for _ in range(3):
    ops.energy_fwd_bwd(f, y, g, l, 1.0, 1.0, True, mode='tile')
torch.cuda.synchronize()
PY
for K in ${NCU_KS:-10}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_ -s 2 -c 1 -o gpurun_out/${TAG}_k4_K${K} python /tmp/k4one.py $K > gpurun_out/${TAG}_ncu_k4_K${K}.log 2>&1
done
ls -la gpurun_out | tail -3
