TAG=${TAG:-r2z}
set -x
python tools/r2_exch1.py 3 2>&1 | tail -2
python tools/r2_exch1.py 10 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_exchange -s 4 -c 1 -o gpurun_out/${TAG}_exch_K10 python tools/r2_exch1.py 10 > gpurun_out/${TAG}_ncu_exch.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_exch.log
