set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "image_to_uint8 or png or generate" --timeout 300 --timeout-method thread > gpurun_out/s4e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s4e_tests.log
tail -15 gpurun_out/s4e_tests.log
timeout 300 python tools/kbench.py K9 > gpurun_out/s4e_kbench.jsonl 2> gpurun_out/s4e_kbench.err
cat gpurun_out/s4e_kbench.jsonl | cut -c1-200; tail -3 gpurun_out/s4e_kbench.err
