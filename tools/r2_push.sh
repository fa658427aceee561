# round 2: push-based peer exchange -- 1-GPU tests + loopback timing, then N-GPU parity + sweep.  TAG=... NG=2 bash tools/r2_push.sh
TAG=${TAG:-r2ai}; NG=${NG:-2}
timeout 300 python -m pytest tests -m gpu -x -q -k "kmeans or peer or prototype" --timeout 200 2>&1 | tail -2
python tools/r2_exch1.py 3 2>&1 | tail -1
python tools/r2_exch1.py 10 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
DD_PEER_TIMEOUT_MS=20000 timeout 600 $TR --master-port 29541 tools/proto_sweep.py --parity --ks 3,5,10 > gpurun_out/${TAG}_sweep_n${NG}.jsonl 2> gpurun_out/${TAG}_sweep_n${NG}.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_sweep_n${NG}.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        if "dist_parity" in d: print("parity", d["dist_parity"], json.dumps(d["detail"])[:300])
        else: print(d["K"], d["exchange"], d["ms_per_iteration"], d.get("exchange_phases_us_rank0"))
PY
tail -3 gpurun_out/${TAG}_sweep_n${NG}.err
