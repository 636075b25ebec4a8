#!/bin/bash
# Evidence pass on one box: full GPU test suite, smoke, compute-sanitizer memcheck + racecheck on small shapes of the
# tensor-core kernels, the full bench line, the ncu launch list of the same step and a full capture of the dominant kernel.
# Usage: bash profiles/r2_gpu_final.sh [tag]
tag=${1:-r2final}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1
echo "== pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"; grep -E "^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | head -20
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/${tag}_smoke.log)"
SAN_TESTS="tests/test_attention_gpu.py::test_tensor_core_dense_global_attention tests/test_attention_gpu.py::test_global_eventful_attention_sequence tests/test_attention_gpu.py::test_delta_with_static_input_is_stationary tests/test_linear_gpu.py tests/test_variants_gpu.py::test_pool_index_kernel_matches_unique tests/test_modules_gpu.py"
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --print-limit 20 python -m pytest $SAN_TESTS tests/test_variants_gpu.py -k "not vitdet_b and not small_ and not 16384" -x -q -p no:cacheprovider > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
echo "== memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/${tag}_sanitizer_memcheck.log | tail -1) / $(grep -E 'passed|failed' gpurun_out/${tag}_sanitizer_memcheck.log | tail -1)"
timeout 900 compute-sanitizer --tool racecheck --report-api-errors no --print-limit 20 python -m pytest tests/test_attention_gpu.py::test_tensor_core_dense_global_attention "tests/test_attention_gpu.py::test_global_eventful_attention_sequence[768-12-grid0-rel0-0-100-2]" "tests/test_attention_gpu.py::test_window_attention[768-12-grid0-window0-rel0-2]" tests/test_linear_gpu.py::test_linear_scatter_epilogue -x -q -p no:cacheprovider > gpurun_out/${tag}_sanitizer_racecheck.log 2>&1
echo "== racecheck: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/${tag}_sanitizer_racecheck.log | tail -1) / $(grep -E 'passed|failed' gpurun_out/${tag}_sanitizer_racecheck.log | tail -1)"
bash profiles/r2_gpu_bench.sh ${tag} tc_apply_kernel
