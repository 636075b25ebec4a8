#!/bin/bash
# Round-2 GPU call 1: full GPU test suite (all failures listed), smoke, sanitizer passes on the hand-rolled
# mbarrier / TMEM pipelines (small shapes), and a quick bench line.
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r2_pytest1.log 2>&1
echo "== pytest: $(tail -1 gpurun_out/r2_pytest1.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest1.log | head -40
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke1.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/r2_smoke1.log)"
SAN_TESTS="tests/test_attention_gpu.py::test_tensor_core_dense_global_attention tests/test_attention_gpu.py::test_delta_with_static_input_is_exactly_stationary tests/test_linear_gpu.py"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SAN_TESTS -x -q -p no:cacheprovider -k "not persistent or True" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1) / $(grep -E 'passed|failed' gpurun_out/r2_sanitizer_$tool.log | tail -1)"
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "== bench rc=$?"; head -c 600 gpurun_out/r2_bench1.json
