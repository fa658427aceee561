# round 2: K4 tile kernel -- parity tests, micro-benchmark (both mappings), one ncu --set full capture.  TAG=... bash tools/r2_k4.sh
TAG=${TAG:-r2a}
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "energy" --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
KBENCH_KS=3,10 timeout 600 python tools/kbench.py K4 > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
cat gpurun_out/${TAG}_kbench.jsonl | cut -c1-160
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
for K in 3 10; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_tile -s 2 -c 1 -o gpurun_out/${TAG}_k4_K${K} python /tmp/k4one.py $K > gpurun_out/${TAG}_ncu_k4_K${K}.log 2>&1
done
ls -la gpurun_out | tail -5
