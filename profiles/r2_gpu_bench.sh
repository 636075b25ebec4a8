#!/bin/bash
# GPU bench pass: full default bench line, ncu launch list of the same step, ncu --set full capture of the dominant kernel.
# Usage: bash profiles/r2_gpu_bench.sh [tag] [kernel regex for the full capture]
tag=${1:-r2}
pattern=${2:-tc_apply_kernel}
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "== bench rc=$? $(head -c 300 gpurun_out/${tag}_bench.json)"
tail -3 gpurun_out/${tag}_bench.err
# launch list of one graph-free run: dense flush + incremental frames (serialised, cold cache: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-graph > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "== launch list rc=$? $(wc -l < gpurun_out/${tag}_launches.csv) lines"
# the dominant kernel, full set, 2 launches of the steady state
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${pattern} -s 8 -c 2 -o gpurun_out/${tag}_top \
    python bench.py --steps 2 --warmup 3 --quick --no-graph > gpurun_out/${tag}_ncu_full.log 2>&1; echo "== full capture rc=$?"
ls -la gpurun_out/${tag}_top.ncu-rep 2>/dev/null
