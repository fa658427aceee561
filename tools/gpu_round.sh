set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/s2_gpu.txt
cp /root/repo/MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_tests.log
timeout 900 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
timeout 600 python tools/kbench.py > gpurun_out/s2_kbench.jsonl 2> gpurun_out/s2_kbench.err
for K in 3 10; do
KBENCH_KS=$K timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_assign --launch-skip 3 -c 1 -f -o gpurun_out/s2_k3_K$K python tools/kbench.py K3 --short > gpurun_out/s2_ncu_k3_$K.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rownorm_classsum --launch-skip 3 -c 1 -f -o gpurun_out/s2_k1 python tools/kbench.py K1 --short > gpurun_out/s2_ncu_k1.log 2>&1
ls -la gpurun_out
