# One GPU-box round: parity tests, bench line, kernel micro-benchmarks, ncu captures.  Outputs under gpurun_out/$TAG_*.
TAG=${TAG:-s3}
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/${TAG}_gpu.txt
cp /root/repo/MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python tools/kbench.py > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
for K in 3 5 10; do
KBENCH_KS=$K timeout 300 ncu --set full --clock-control none --import-source on -k regex:"kmeans_(pair|stream)" --launch-skip 3 -c 1 -f -o gpurun_out/${TAG}_k3_K$K python tools/kbench.py K3 --short > gpurun_out/${TAG}_ncu_k3_$K.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rownorm_stream --launch-skip 3 -c 1 -f -o gpurun_out/${TAG}_k1 python tools/kbench.py K1 --short > gpurun_out/${TAG}_ncu_k1.log 2>&1
ls -la gpurun_out
