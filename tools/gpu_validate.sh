# Full single-GPU validation: every GPU parity test, smoke(), the bench line (ours + reference arm).  TAG=... bash tools/gpu_validate.sh
TAG=${TAG:-s5g}
set -x
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 400 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -c 300 gpurun_out/${TAG}_bench_ref.json
