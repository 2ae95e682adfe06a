#!/bin/bash
# Quick GPU check: parity tests, the default bench line (+ fast mode) and CTA 0's timeline.  Usage: bash tools/gpu_quick.sh <tag>
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg2_parity.json
timeout 300 python bench.py --steps 20 --warmup 3 --precision fast --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg2_fast.json
for wl in nerf cfg4 cfg1; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_${wl}_parity.json
done
timeout 120 python tools/trace_timeline.py cfg2 > $OUT/timeline_cfg2_parity.txt 2>&1
timeout 120 python tools/trace_timeline.py cfg2 fast > $OUT/timeline_cfg2_fast.txt 2>&1
