# Session-5 GPU round at HEAD: full parity suite, bench line (ours + reference arm), kernel micro-benchmarks, ncu captures.
TAG=${TAG:-s5}
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/${TAG}_gpu.txt
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
timeout 600 python tools/kbench.py > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
for K in 5 10; do
KBENCH_KS=$K timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kmeans_(pair|stream|ws)" --launch-skip 3 -c 1 -f -o gpurun_out/${TAG}_k3_K$K python tools/kbench.py K3 --short > gpurun_out/${TAG}_ncu_k3_$K.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bicubic --launch-skip 6 -c 2 -f -o gpurun_out/${TAG}_k8 python tools/kbench.py K8 --short > gpurun_out/${TAG}_ncu_k8.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:image_u8 --launch-skip 6 -c 1 -f -o gpurun_out/${TAG}_k9 python tools/kbench.py K9 --short > gpurun_out/${TAG}_ncu_k9.log 2>&1
ls -la gpurun_out
