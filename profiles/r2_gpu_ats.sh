#!/bin/bash
# ATS bring-up: the tests that touch adaptive token sampling, all failures listed.  Usage: bash profiles/r2_gpu_ats.sh [tag]
tag=${1:-r2ats}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_variants_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -k "ats" > gpurun_out/${tag}_pytest.log 2>&1
echo "== ats tests: $(tail -1 gpurun_out/${tag}_pytest.log)"; grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/${tag}_pytest.log | head -40
