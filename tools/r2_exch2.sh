TAG=${TAG:-r2aa}
set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "kmeans or peer or prototype" --timeout 200 2>&1 | tail -3
python tools/r2_exch1.py 3 2>&1 | tail -1
python tools/r2_exch1.py 10 2>&1 | tail -1
timeout 300 python tools/proto_sweep.py --ks 3,5,10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['K'], d['ms_per_iteration'], d['frac_of_world_x_hbm_peak'])"
