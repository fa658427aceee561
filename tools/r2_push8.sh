TAG=${TAG:-r2aj}; NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
DD_PEER_TIMEOUT_MS=20000 timeout 300 $TR --master-port 29541 tools/proto_sweep.py --parity --ks 3,5,10 > gpurun_out/${TAG}_sweep_n${NG}.jsonl 2> gpurun_out/${TAG}_sweep_n${NG}.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_sweep_n${NG}.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        if "dist_parity" in d: print("parity", d["dist_parity"], json.dumps(d["detail"])[:300])
        else: print(d["K"], d["exchange"], d["ms_per_iteration"], d.get("exchange_phases_us_rank0"))
PY
tail -3 gpurun_out/${TAG}_sweep_n${NG}.err
