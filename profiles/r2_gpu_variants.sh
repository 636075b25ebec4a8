#!/bin/bash
# Same-box comparison of library build variants: quick bench (8 streams) per variant.  Usage: bash profiles/r2_gpu_variants.sh tag lib1 lib2 ...
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
  export EVENTFUL_B200_LIB=$PWD/eventful-transformer_b200/lib/$lib
  timeout 300 python bench.py --quick --streams 8 > gpurun_out/${tag}_${lib}_${rep}.json 2> /dev/null
  python - gpurun_out/${tag}_${lib}_${rep}.json $lib $rep <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], "rep", sys.argv[3], ":", d["value"], "fps  e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("unparsed:", e)
PY
done
done
