#!/bin/bash
# Round-2 GPU-box visit: bench lines of every workload, ncu launch lists and full captures (fused kernel, training GEMMs),
# trace timeline, compute-sanitizer on the new kernels.  Usage (repo root, under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/nproc.txt
echo "== bench parity"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $OUT/bench_cfg2_parity.json
echo "== bench fast"; timeout 300 python bench.py --steps 20 --warmup 5 --precision fast --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_cfg2_fast.json
for wl in nerf cfg4 cfg1 cfg3 paper; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_${wl}_parity.json
done
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_cfg5_parity_1gpu.json
echo "== train"; timeout 600 python bench.py --workload train --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_train_parity.json
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --precision fast --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_train_fast.json
echo "== reference arms"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > $OUT/bench_reference_cpu.json
timeout 300 python bench.py --impl reference --device cuda --steps 10 --warmup 2 2>&1 | tail -1 > $OUT/bench_reference_torch_cuda.json
if [ -n "$BENCH_ONLY" ]; then ls -la $OUT; exit 0; fi      # BENCH_ONLY=1: just the bench lines (after a change that leaves the kernels alone)
echo "== per-kernel time of a training step"; timeout 300 python tools/train_profile.py 2048 > $OUT/train_step_kernels.txt 2>&1
echo "== trace timeline"; timeout 300 python tools/trace_timeline.py cfg2 > $OUT/timeline_cfg2_parity.txt 2>&1
timeout 300 python tools/trace_timeline.py cfg2 fast > $OUT/timeline_cfg2_fast.txt 2>&1
echo "== ncu launch list (inference bench)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_cfg2_parity.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "== ncu full: fused kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nrf_fused -s 3 -c 1 -f -o $OUT/prof_fused \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "== ncu launch list (durations + DRAM bytes) + full captures: training step"
# tools/train_profile.py runs 8 steps; one step = 50 tile_gemm + 28 dw_gemm launches.  tile_gemm 315 = a fine-pass 256x256 forward layer
# (393,216 samples), 342 = a fine-pass dX launch, dw_gemm 188 = a fine-pass 256x256 dW.
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:'tile_gemm|dw_gemm|head_bwd|colsum|bias_grad|dw_reduce|heads_kernel|encode_planes|composite|smpl_points|rayfeat|ray_bias|split_planes|fine_sampling|absmax|scale_from|ray_feats|points_from' \
   -s 900 -c 420 --csv --log-file $OUT/launches_train.csv python tools/train_profile.py 2048 > $OUT/ncu_launches_train.log 2>&1
for skip in 315 342; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_gemm -s $skip -c 1 -f -o $OUT/prof_tile_gemm_$skip \
   python tools/train_profile.py 2048 > $OUT/ncu_full_tile_$skip.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dw_gemm -s 188 -c 1 -f -o $OUT/prof_dw_gemm \
   python tools/train_profile.py 2048 > $OUT/ncu_full_dw.log 2>&1
echo "== gradient error tables"
for k in nerf append smpl; do timeout 300 python tools/dbg_grad.py $k > $OUT/grad_errors_$k.txt 2>&1; done
echo "== compute-sanitizer (memcheck) on the training tests"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -x -q -k "gemm or forward_matches or fp64_autograd and nerf" > $OUT/sanitizer_train.txt 2>&1; echo "rc=$?" >> $OUT/sanitizer_train.txt
ls -la $OUT; du -sh gpurun_out
