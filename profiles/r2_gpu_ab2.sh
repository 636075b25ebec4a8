#!/bin/bash
# A/B pass 2: tc_apply with the selected-key index staged in shared memory (default build) and with 16 state-mover warps
# (libeventful_b200_mv16.so, setmaxnreg-rebalanced); MUFU rate microbenchmark.  Usage: bash profiles/r2_gpu_ab2.sh [tag]
tag=${1:-r2g}
mkdir -p gpurun_out
(cd profiles/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate) > gpurun_out/${tag}_mufu_rate.txt 2>&1; cat gpurun_out/${tag}_mufu_rate.txt
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], "fps  e2e", d.get("e2e", {}).get("value"), " launches/step", d.get("launches_per_step"), " single", d.get("single_stream", {}).get("value"), " apply frac", d.get("roofline", {}).get("frac"))
except Exception as e:
    print("unparsed:", e)
PY
}
for variant in cur mv16; do
    unset EVENTFUL_B200_LIB
    [ $variant = mv16 ] && export EVENTFUL_B200_LIB=$PWD/eventful-transformer_b200/lib/libeventful_b200_mv16.so
    timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_backbone_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_${variant}_pytest_attn.log 2>&1
    echo "== $variant attention+backbone tests: $(tail -1 gpurun_out/${tag}_${variant}_pytest_attn.log)"
    timeout 300 python bench.py --quick > gpurun_out/${tag}_${variant}.json 2> gpurun_out/${tag}_${variant}.err
    echo "== $variant rc=$? $(line gpurun_out/${tag}_${variant}.json)"
done
unset EVENTFUL_B200_LIB
