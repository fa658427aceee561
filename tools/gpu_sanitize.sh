# compute-sanitizer over a reduced-size run of every dd_* entry point (tools/sanitize_run.py).  TAG=... bash tools/gpu_sanitize.sh
# memcheck: whole set; racecheck / synccheck: per kernel family (they are ~100x slower).  Summaries land in gpurun_out/${TAG}_san_*.log
TAG=${TAG:-r2s}
CS=/usr/local/cuda/bin/compute-sanitizer
set -x
for tool in memcheck synccheck racecheck; do
  for fam in K5 K4 K8 K1; do
    extra=""; [ "$fam" = "K1" ] && extra="--peer"
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_run.py $fam $extra > gpurun_out/${TAG}_san_${tool}_${fam}.log 2>&1
    echo "$tool $fam rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_san_${tool}_${fam}.log | tail -1)" | tee -a gpurun_out/${TAG}_san_summary.txt
  done
done
# racecheck models happens-before per thread: the elected-lane mbarrier arrive of the K3 pair kernel (one lane arrives for its
# warp after __syncwarp) is reported as a hazard for the other 31 lanes.  Same protocol with every lane arriving:
DD_EXTRA_NVCC_FLAGS="-DDD_ALL_LANES_ARRIVE" python -m distdiff_b200.build > gpurun_out/${TAG}_san_rebuild.log 2>&1
timeout 900 $CS --tool racecheck --print-limit 20 --error-exitcode 9 python tools/sanitize_run.py K1 --peer > gpurun_out/${TAG}_san_racecheck_K1_all_lanes_arrive.log 2>&1
echo "racecheck K1 (all lanes arrive) rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/${TAG}_san_racecheck_K1_all_lanes_arrive.log | tail -1)" | tee -a gpurun_out/${TAG}_san_summary.txt
python -m distdiff_b200.build > /dev/null 2>&1   # back to the product build
