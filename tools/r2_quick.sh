# quick: k-means parity + K3 microbench + sanitizer pass.  TAG=... bash tools/r2_quick.sh
TAG=${TAG:-r2t}
set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "kmeans or guided_step" --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
KBENCH_KS=3,4,5,8,10 timeout 600 python tools/kbench.py K3 > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
grep -v eager gpurun_out/${TAG}_kbench.jsonl | cut -c1-140; tail -3 gpurun_out/${TAG}_kbench.err
TAG=${TAG} bash tools/gpu_sanitize.sh
