set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gpu_tests.log
tail -6 gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo ref rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-strong > gpurun_out/r02_bench_under_ncu.log 2>&1; echo ncu-list rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_hensel -s 2 -c 1 -f -o /tmp/dec python tools/ncu_one.py 2048 65536 > gpurun_out/r02_ncu_decrypt.log 2>&1
ncu -i /tmp/dec.ncu-rep --page raw --csv > gpurun_out/r02_ncu_decrypt_raw.csv 2>&1
python tools/ncu_pick.py gpurun_out/r02_ncu_decrypt_raw.csv > gpurun_out/r02_ncu_decrypt_thread_per_task.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encrypt_hensel -s 2 -c 1 -f -o /tmp/enc python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-strong > gpurun_out/r02_ncu_encrypt.log 2>&1
ncu -i /tmp/enc.ncu-rep --page raw --csv > gpurun_out/r02_ncu_encrypt_raw.csv 2>&1
python tools/ncu_pick.py gpurun_out/r02_ncu_encrypt_raw.csv > gpurun_out/r02_ncu_encrypt_hensel.json
tests/cpp/_build/bench_ipcl > gpurun_out/r02_bench_ipcl_resident.jsonl 2>&1
IPCL_B200_DEVICE_RESIDENT=0 tests/cpp/_build/bench_ipcl > gpurun_out/r02_bench_ipcl_hostroundtrip.jsonl 2>&1
du -sh gpurun_out
