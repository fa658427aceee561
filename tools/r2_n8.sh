# round 2, 8 GPUs: configs[3] sweep + parity, then the bench under torchrun (short).  TAG=... bash tools/r2_n8.sh
TAG=${TAG:-r2x}; NG=${NG:-8}
set -x
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/proto_sweep.py --parity --ks 3,5,10 > gpurun_out/${TAG}_sweep_n${NG}.jsonl 2> gpurun_out/${TAG}_sweep_n${NG}.err
tail -3 gpurun_out/${TAG}_sweep_n${NG}.err
timeout 900 $TR --master-port 29543 bench.py --gpus $NG --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_n${NG}.json 2> gpurun_out/${TAG}_bench_n${NG}.err
tail -c 400 gpurun_out/${TAG}_bench_n${NG}.json; tail -3 gpurun_out/${TAG}_bench_n${NG}.err
