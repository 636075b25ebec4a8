#!/bin/bash
# A/B pass on one box: new GEMM / gate tests first, then quick bench lines (8 streams and 1 stream) for
#   a) the current build, b) the same with the round-1 gate schedule (gate_gather + dense-A GEMM), c) mbarrier waits without
#   the suspend-time hint.  Usage: bash profiles/r2_gpu_ab.sh [tag]
tag=${1:-r2ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_gate_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest_new.log 2>&1
echo "== new tests: $(tail -1 gpurun_out/${tag}_pytest_new.log)"; grep -E "^FAILED|^ERROR|Error" gpurun_out/${tag}_pytest_new.log | head -10
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], "fps  e2e", d.get("e2e", {}).get("value"), " launches/step", d.get("launches_per_step"), " single", d.get("single_stream", {}).get("value"))
except Exception as e:
    print("unparsed:", e)
PY
}
for variant in cur nofuse nohint; do
  for streams in 8 1; do
    export EVENTFUL_B200_FUSE_GATHER=1; unset EVENTFUL_B200_LIB
    [ $variant = nofuse ] && export EVENTFUL_B200_FUSE_GATHER=0
    [ $variant = nohint ] && export EVENTFUL_B200_LIB=$PWD/eventful-transformer_b200/lib/libeventful_b200_nohint.so
    timeout 300 python bench.py --quick --streams $streams > gpurun_out/${tag}_${variant}_s${streams}.json 2> gpurun_out/${tag}_${variant}_s${streams}.err
    echo "== $variant streams=$streams rc=$? $(line gpurun_out/${tag}_${variant}_s${streams}.json)"
  done
done
unset EVENTFUL_B200_LIB; export EVENTFUL_B200_FUSE_GATHER=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/${tag}_pytest.log 2>&1
echo "== pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"
grep -E "^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | head -40
