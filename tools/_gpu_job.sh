set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo bench8 rc=$?
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > gpurun_out/r02_bench_ref_8gpu.json 2> gpurun_out/r02_bench_ref_8gpu.err; echo ref rc=$?
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 5 --warmup 3 --no-configs > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo bench4 rc=$?
python - <<'PY'
import json
for f in ['gpurun_out/r02_bench_8gpu.json','gpurun_out/r02_bench_4gpu.json','gpurun_out/r02_bench_ref_8gpu.json']:
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, {k:d.get(k) for k in ['value','ms_per_step','n_gpus']}, (d.get('e2e') or {}).get('value'), json.dumps(d.get('strong'))[:900], (d.get('cpu_baseline') or {}).get('cores'))
PY
