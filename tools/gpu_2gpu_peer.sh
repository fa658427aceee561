TAG=${TAG:-s5l}
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/${TAG}_dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_check.log
grep -E "dist_check|rc=|Error|error" gpurun_out/${TAG}_dist_check.log | tail -12
for X in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/proto_sweep.py --exchange $X --ks 3,5,10 > gpurun_out/${TAG}_sweep_n2_$X.jsonl 2> gpurun_out/${TAG}_sweep_n2_$X.err
cut -c60-260 gpurun_out/${TAG}_sweep_n2_$X.jsonl; grep -iE "error|Traceback" -A3 gpurun_out/${TAG}_sweep_n2_$X.err | tail -8
done
