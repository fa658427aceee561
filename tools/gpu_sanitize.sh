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
