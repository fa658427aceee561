# ncu --set full captures of the HBM-bound kernels at micro-benchmark sizes (1 GPU). Usage: bash tools/gpu_ncu.sh <tag>
TAG=${1:-r1}
set -x
for K in 3 10; do
KBENCH_KS=$K timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_stream --launch-skip 3 -c 1 -f -o gpurun_out/${TAG}_k3_K$K python tools/kbench.py K3 --short > gpurun_out/${TAG}_ncu_k3_$K.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rownorm_stream --launch-skip 3 -c 1 -f -o gpurun_out/${TAG}_k1 python tools/kbench.py K1 --short > gpurun_out/${TAG}_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy --launch-skip 40 -c 1 -f -o gpurun_out/${TAG}_k4 python tools/kbench.py K4 --short > gpurun_out/${TAG}_ncu_k4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cfg_ddim_fwd --launch-skip 28 -c 1 -f -o gpurun_out/${TAG}_k5 python tools/kbench.py K5 --short > gpurun_out/${TAG}_ncu_k5.log 2>&1
ls -la gpurun_out
