# compute-sanitizer passes over the small parity tests (goldens, chain batches, packed layouts):
#   gpurun -- 'bash scripts/sanitize.sh'   -> gpurun_out/sanitizer_<tool>.log
SEL='golden or per_bin_chain_batches or edge_cases or zero_length or api_corner'
for tool in racecheck initcheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_packed_events.py -m gpu -x -q -k "$SEL" \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
