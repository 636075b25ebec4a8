#!/bin/bash
# Debug helper for the GPU box: every test file in its own process (a sticky CUDA error cannot cascade),
# blocking launches, stop at the first failure, then compute-sanitizer on that first failing test.
mkdir -p gpurun_out
for f in ${@:-gate linear attention backbone}; do
  CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_${f}_gpu.py ${PYTEST_X:--x} -q --timeout 300 -p no:cacheprovider > gpurun_out/dbg_${f}.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/dbg_${f}.log)"
  first=$(grep -m1 '^FAILED' gpurun_out/dbg_${f}.log | sed 's/^FAILED //; s/ - .*//')
  if [ -n "$first" ]; then
    timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "$first" -x -q -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | grep -E "Invalid|Misaligned|at |by thread|Address|=========     in |Saved host|et_|kernel" | head -40 > gpurun_out/san_${f}.log
    echo "-- sanitizer $first"; head -14 gpurun_out/san_${f}.log
  fi
done
