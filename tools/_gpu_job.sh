set -x
mkdir -p gpurun_out
timeout 300 python tools/t1_probe.py 2048 65536 16384 > gpurun_out/r2b_t1_probe.jsonl 2>&1
cat gpurun_out/r2b_t1_probe.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_gpu_tests.log
tail -5 gpurun_out/r2b_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-strong > gpurun_out/r2b_bench_under_ncu.log 2>&1; echo ncu-list rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_hensel -s 2 -c 1 -f -o /tmp/r2b_hensel_final python tools/hensel_probe.py 2048 65536 > gpurun_out/r2b_ncu_hensel.log 2>&1; echo ncu-full rc=$?
ncu -i /tmp/r2b_hensel_final.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_hensel_raw.csv 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encrypt_hensel -s 2 -c 1 -f -o /tmp/r2b_encrypt_final python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-strong > gpurun_out/r2b_ncu_encrypt.log 2>&1; echo ncu-enc rc=$?
ncu -i /tmp/r2b_encrypt_final.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_encrypt_raw.csv 2>&1
tests/cpp/_build/bench_ipcl 2048 65536 > gpurun_out/r2b_bench_ipcl_resident.jsonl 2>&1
IPCL_B200_DEVICE_RESIDENT=0 tests/cpp/_build/bench_ipcl 16 256 2048 65536 > gpurun_out/r2b_bench_ipcl_hostroundtrip.jsonl 2>&1
du -sh gpurun_out
