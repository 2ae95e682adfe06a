#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list and one full capture of the fused kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench parity"; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_cfg2_parity.json
echo "== bench fast"; timeout 300 python bench.py --steps 20 --warmup 3 --precision fast --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg2_fast.json
for wl in nerf cfg4 cfg1; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_${wl}_parity.json
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nrf_fused -s 3 -c 1 -f -o $OUT/prof_fused \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
