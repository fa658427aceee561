set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "kmeans or proto" --timeout 90 --timeout-method thread > gpurun_out/s3j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3j_tests.log
tail -5 gpurun_out/s3j_tests.log
timeout 300 python tools/kbench.py K3 > gpurun_out/s3j_kbench.jsonl 2> gpurun_out/s3j_kbench.err
cat gpurun_out/s3j_kbench.jsonl | cut -c1-400
for K in 10; do
KBENCH_KS=$K timeout 300 ncu --set full --clock-control none --import-source on -k regex:kmeans_pair --launch-skip 3 -c 1 -f -o gpurun_out/s3j_k3_K$K python tools/kbench.py K3 --short > gpurun_out/s3j_ncu_k3_$K.log 2>&1
done
ls -la gpurun_out
