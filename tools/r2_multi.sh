# round 2, N GPUs (gpurun --gpus N): multi-GPU parity + configs[3] sweep + bench.py code paths (tiny) under torchrun.  TAG=... NG=2 bash tools/r2_multi.sh
TAG=${TAG:-r2m}; NG=${NG:-2}
set -x
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/proto_sweep.py --parity --ks 3,5,10 > gpurun_out/${TAG}_sweep_n${NG}.jsonl 2> gpurun_out/${TAG}_sweep_n${NG}.err
cut -c1-260 gpurun_out/${TAG}_sweep_n${NG}.jsonl; tail -5 gpurun_out/${TAG}_sweep_n${NG}.err
if [ "$NG" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 500 --timeout-method thread > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
fi
timeout 600 $TR --master-port 29543 bench.py --gpus $NG --tiny --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_tiny_n${NG}.json 2> gpurun_out/${TAG}_bench_tiny_n${NG}.err
tail -c 1500 gpurun_out/${TAG}_bench_tiny_n${NG}.json; tail -5 gpurun_out/${TAG}_bench_tiny_n${NG}.err
