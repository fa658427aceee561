TAG=${TAG:-s5j}
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "cfg_ddim or denoise or guidance or graph" --timeout 300 --timeout-method thread > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python tools/kbench.py K5 > gpurun_out/${TAG}_kbench.jsonl 2> gpurun_out/${TAG}_kbench.err
grep fwd gpurun_out/${TAG}_kbench.jsonl | cut -c1-110
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cfg_ddim_fwd --launch-skip 5 -c 1 -f -o gpurun_out/${TAG}_k5_instep python tools/instep_k5.py 8 > gpurun_out/${TAG}_ncu_k5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cfg_ddim_fwd --csv --log-file gpurun_out/${TAG}_k5_instep_B4.csv python tools/instep_k5.py 4 > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_k5_instep_B4.csv
