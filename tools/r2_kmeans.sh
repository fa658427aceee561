# round 2: fused slot reduction + PDL Lloyd loop.  1 GPU: k-means parity tests + sweep; TAG=... bash tools/r2_kmeans.sh
TAG=${TAG:-r2h}
set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "kmeans or peer or prototype or rownorm" --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 600 python tools/proto_sweep.py --ks 3,5,10 > gpurun_out/${TAG}_sweep_n1.jsonl 2> gpurun_out/${TAG}_sweep_n1.err
cut -c1-230 gpurun_out/${TAG}_sweep_n1.jsonl; tail -3 gpurun_out/${TAG}_sweep_n1.err
KBENCH_KS=3,5,10 timeout 600 python tools/kbench.py K3 K1 > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
cut -c1-200 gpurun_out/${TAG}_kbench.jsonl; tail -3 gpurun_out/${TAG}_kbench.err
