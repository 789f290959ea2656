set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_gpu_tests.log
tail -8 gpurun_out/r2i_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; tail -2 gpurun_out/r2i_smoke.log
timeout 600 python bench.py > gpurun_out/r2i_bench_1gpu.json 2> gpurun_out/r2i_bench_1gpu.err; echo bench rc=$?
python - <<'PY'
import json
for line in open('gpurun_out/r2i_bench_1gpu.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ['value','ms_per_step','encrypt_per_s','decrypt_per_s']}, d['e2e']['value'], d['roofline']['executed_frac'], d['roofline']['launch_ms'])
PY
