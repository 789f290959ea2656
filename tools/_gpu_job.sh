mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r02_sanitizer_$tool.txt
done
tools/microbench/chain_probe > gpurun_out/r02_chain_probe.jsonl; head -30 gpurun_out/r02_chain_probe.jsonl
