TAG=${TAG:-r2y}; NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/proto_sweep.py --ks 3,10 > gpurun_out/${TAG}_sweep_n${NG}.jsonl 2> gpurun_out/${TAG}_sweep_n${NG}.err
tail -3 gpurun_out/${TAG}_sweep_n${NG}.err
