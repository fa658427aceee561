# round 2: tensor-core K3 -- parity tests (k-means), sweep, microbench; ncu of the K=10 kernel.  TAG=... bash tools/r2_k3mma.sh
TAG=${TAG:-r2p}
set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "kmeans or peer or prototype or guided_step" --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
KBENCH_KS=4,5,8,10 timeout 600 python tools/kbench.py K3_ > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
cut -c1-200 gpurun_out/${TAG}_kbench.jsonl; tail -3 gpurun_out/${TAG}_kbench.err
timeout 600 python tools/proto_sweep.py --ks 3,5,10 > gpurun_out/${TAG}_sweep_n1.jsonl 2> gpurun_out/${TAG}_sweep_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/'+__import__('os').environ.get('TAG','r2p')+'_sweep_n1.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['K'], d['ms_per_iteration'], d['frac_of_world_x_hbm_peak'])
PY
cat > /tmp/k3one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from distdiff_b200 import ops
K = int(sys.argv[1]); dev = torch.device('cuda:0'); N, C, D = 100000, 100, 2048
feat = torch.randn(N, D, device=dev); labels = torch.arange(N, device=dev) % C
perm, off = ops.sort_by_class(labels, C)
xs, _, _ = ops.rownorm_classsum(feat, perm, off)
buf = ops.KMeansBuffers(N, D, C, K, dev)
idx = off[:-1, None] + (torch.arange(K, device=dev)[None, :] * (off[1:] - off[:-1])[:, None]) // K
s, c = ops.kmeans_seed(xs, idx.contiguous()); ops.kmeans_update(s, c, buf.centroid, buf.cnorm)
for _ in range(3):
    ops.kmeans_assign_accum(xs, off, buf)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_mma -s 1 -c 1 -o gpurun_out/${TAG}_k3_K10 python /tmp/k3one.py 10 > gpurun_out/${TAG}_ncu_k3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_k3.log
