#!/bin/bash
# GPU test pass: full suite (all failures listed), smoke, compute-sanitizer memcheck over the tensor-core and the
# general-precision kernels on small shapes.  Usage: bash profiles/r2_gpu_tests.sh [tag]
tag=${1:-r2}
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1
echo "== pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"
grep -E "^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | head -40
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/${tag}_smoke.log)"
if [ -z "$NO_SANITIZER" ]; then
SAN_TESTS="tests/test_attention_gpu.py::test_tensor_core_dense_global_attention tests/test_attention_gpu.py::test_delta_with_static_input_is_exactly_stationary tests/test_linear_gpu.py tests/test_variants_gpu.py::test_pool_index_kernel_matches_unique tests/test_modules_gpu.py"
timeout 600 compute-sanitizer --tool memcheck --report-api-errors no --print-limit 20 python -m pytest $SAN_TESTS tests/test_variants_gpu.py -k "not vitdet_b and not small_" -x -q -p no:cacheprovider > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
echo "== memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/${tag}_sanitizer_memcheck.log | tail -1) / $(grep -E 'passed|failed' gpurun_out/${tag}_sanitizer_memcheck.log | tail -1)"
fi
