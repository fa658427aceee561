TAG=${TAG:-s5i}
set -x
timeout 600 python tools/proto_sweep.py > gpurun_out/${TAG}_sweep_n1.jsonl 2> gpurun_out/${TAG}_sweep_n1.err
cut -c1-200 gpurun_out/${TAG}_sweep_n1.jsonl; tail -3 gpurun_out/${TAG}_sweep_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cfg_ddim_fwd --launch-skip 5 -c 1 -f -o gpurun_out/${TAG}_k5_instep python tools/instep_k5.py 8 > gpurun_out/${TAG}_ncu_k5.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_k5.log
