TAG=${TAG:-s5q}
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "bicubic or guidance" --timeout 300 --timeout-method thread > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python tools/kbench.py K8 > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
grep -v aten gpurun_out/${TAG}_kbench.jsonl | cut -c1-110; tail -3 gpurun_out/${TAG}_kbench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bicubic_fwd --launch-skip 8 -c 1 -f -o gpurun_out/${TAG}_k8f python tools/kbench.py K8 --short > gpurun_out/${TAG}_ncu_k8f.log 2>&1
