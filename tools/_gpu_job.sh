mkdir -p gpurun_out
PROBE_LAYOUTS=0,1 timeout 600 python tools/t1_probe.py 3072 8192 32768 65536 > gpurun_out/r2k_layout_3072.jsonl 2>&1
PROBE_LAYOUTS=0,1 timeout 600 python tools/t1_probe.py 4096 8192 32768 >> gpurun_out/r2k_layout_3072.jsonl 2>&1
PROBE_LAYOUTS=0,1 timeout 600 python tools/t1_probe.py 1024 8192 65536 262144 >> gpurun_out/r2k_layout_3072.jsonl 2>&1
cat gpurun_out/r2k_layout_3072.jsonl
