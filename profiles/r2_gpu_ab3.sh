#!/bin/bash
# Same-box A/B of two builds of the library: quick bench lines, alternating, 8 streams and 1 stream.
# Usage: bash profiles/r2_gpu_ab3.sh [tag] [other .so]
tag=${1:-r2ab3}
other=${2:-$PWD/eventful-transformer_b200/lib/libeventful_b200_prev.so}
mkdir -p gpurun_out
for rep in 1 2; do
  for variant in cur other; do
    unset EVENTFUL_B200_LIB
    [ $variant = other ] && export EVENTFUL_B200_LIB=$other
    for streams in 8 1; do
      timeout 300 python bench.py --quick --streams $streams > gpurun_out/${tag}_${variant}_s${streams}_${rep}.json 2> /dev/null
      python - gpurun_out/${tag}_${variant}_s${streams}_${rep}.json $variant $rep <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], "rep", sys.argv[3], d["config"]["streams_per_gpu"], "streams:", d["value"], "fps  e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("unparsed:", e)
PY
    done
  done
done
