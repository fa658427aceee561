# K8 ncu captures (fp32 B=128: the 3rd fwd / bwd launch) + bench batch-size sweep
TAG=${TAG:-s5b}
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bicubic_fwd --launch-skip 8 -c 1 -f -o gpurun_out/${TAG}_k8f python tools/kbench.py K8 --short > gpurun_out/${TAG}_ncu_k8f.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bicubic_bwd --launch-skip 8 -c 1 -f -o gpurun_out/${TAG}_k8b python tools/kbench.py K8 --short > gpurun_out/${TAG}_ncu_k8b.log 2>&1
for B in 8 16; do
timeout 600 python bench.py --batch $B --no-kernels --no-cpu-baseline > gpurun_out/${TAG}_bench_B$B.json 2> gpurun_out/${TAG}_bench_B$B.err
tail -c 1500 gpurun_out/${TAG}_bench_B$B.json
done
ls -la gpurun_out | tail -8
