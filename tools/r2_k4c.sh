TAG=${TAG:-r2ag}
set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_guidance.py -m gpu -x -q -k "energy or guidance" --timeout 300 2>&1 | tail -3
python tools/r2_k4probe.py | tee gpurun_out/${TAG}_probe.jsonl
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
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"energy_|class_sort" -c 12 --csv --log-file gpurun_out/${TAG}_launches.csv python /tmp/k4one.py 10 > /dev/null 2>&1
cat gpurun_out/${TAG}_launches.csv | tail -8 | cut -c1-300
