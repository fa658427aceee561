# usage: NG=2 TAG=s5m bash tools/gpu_ngpu_sweep.sh  -- dist_check + k-means sweep with both exchanges on NG GPUs
TAG=${TAG:-s5m}; NG=${NG:-2}
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/${TAG}_dist_check_n$NG.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_dist_check_n$NG.log
grep -E "dist_check|rc=|Error|error" gpurun_out/${TAG}_dist_check_n$NG.log | tail -8
for X in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tools/proto_sweep.py --exchange $X --ks ${KS:-3,4,5,6,7,8,9,10} > gpurun_out/${TAG}_sweep_n${NG}_$X.jsonl 2> gpurun_out/${TAG}_sweep_n${NG}_$X.err
cut -c75-230 gpurun_out/${TAG}_sweep_n${NG}_$X.jsonl; grep -iE "error|Traceback" -A3 gpurun_out/${TAG}_sweep_n${NG}_$X.err | tail -8
done
