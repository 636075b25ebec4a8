#!/bin/bash
# Quick confirmation pass: attention + backbone GPU tests, then quick bench lines at 8 streams and 1 stream.  Usage: bash profiles/r2_gpu_quick.sh [tag] [pytest targets]
tag=${1:-r2q}
targets=${2:-"tests/test_attention_gpu.py tests/test_backbone_gpu.py"}
mkdir -p gpurun_out
timeout 900 python -m pytest $targets -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1
echo "== tests: $(tail -1 gpurun_out/${tag}_pytest.log)"; grep -E "^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | head
for streams in 8 1; do
timeout 300 python bench.py --quick --streams $streams > gpurun_out/${tag}_s${streams}.json 2> gpurun_out/${tag}_s${streams}.err
python - gpurun_out/${tag}_s${streams}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", d["config"]["streams_per_gpu"], "streams:", d["value"], "fps  e2e", d.get("e2e", {}).get("value"), " launches/step", d.get("launches_per_step"))
except Exception as e:
    print("unparsed:", e)
PY
done
