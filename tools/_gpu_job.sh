mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hensel.py -x -q -k "layouts_agree" > gpurun_out/r2l_pp_tests.log 2>&1; tail -5 gpurun_out/r2l_pp_tests.log
PROBE_LAYOUTS=0,-2,-4 timeout 600 python tools/t1_probe.py 2048 4096 8192 16384 32768 65536 > gpurun_out/r2l_pp_probe.jsonl 2>&1
cat gpurun_out/r2l_pp_probe.jsonl
