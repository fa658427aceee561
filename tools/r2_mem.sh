# memory / batch-size runs of the guided step: gradient checkpointing off/on, B = 8 / 16.  TAG=... bash tools/r2_mem.sh
TAG=${TAG:-r2w}
set -x
X="--steps 2 --warmup 1 --no-kernels --no-cpu-baseline --no-ops-compare --no-proto-sweep"
for cfg in "8 " "8 --grad-ckpt" "16 --grad-ckpt" "16 "; do
  set -- $cfg
  name="B$1$( [ -n "$2" ] && echo _ckpt )"
  timeout 600 python bench.py --batch $1 $2 $X > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
ls=[l for l in open('gpurun_out/${TAG}_bench_${name}.json') if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); print('${name}', d['value'], d['ms_per_step'], d['peak_mem_gb'], d['e2e']['value'])
else:
    print('${name}: no line'); print(open('gpurun_out/${TAG}_bench_${name}.err').read()[-600:])
PY
done
