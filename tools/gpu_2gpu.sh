# 2-GPU check: sharded prototypes over NCCL == single GPU, and the bench under torchrun (image split, weak scaling)
TAG=${TAG:-s5h}
set -x
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 500 --timeout-method thread > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
tail -c 600 gpurun_out/${TAG}_bench_n2.json; tail -3 gpurun_out/${TAG}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err
tail -c 300 gpurun_out/${TAG}_bench_ref_n2.json
