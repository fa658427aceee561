set -x
for n in 100000 50000 25000 12500; do timeout 300 python tools/proto_sweep.py --n $n --ks 3,10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print($n, d['K'], d['ms_per_iteration'])
"; done
