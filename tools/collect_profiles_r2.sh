#!/bin/bash
# Assemble profiles/r2/ (tracked) from gpurun_out/ (scratch): run here, on the CPU box, after tools/gpu_round2.sh + tools/prof_train.sh.
#   bash tools/collect_profiles_r2.sh <sweep dir, e.g. gpurun_out/r2f> [<train capture dir if separate>]
set -e
SW=${1:-gpurun_out/r2f}; TR=${2:-$SW}; OUT=profiles/r2
mkdir -p $OUT
cp $SW/bench_*.json $SW/smi.txt $SW/nproc.txt $OUT/ 2>/dev/null || true
cp $SW/launches_cfg2_parity.csv $OUT/launches_cfg2_parity.csv
cp $SW/sanitizer_train.txt $OUT/ 2>/dev/null || true
cp $SW/train_step_kernels.txt $OUT/ 2>/dev/null || true
cp $SW/grad_errors_*.txt $OUT/ 2>/dev/null || true
rm -f $OUT/ncu_full_prof_tile_gemm_*.txt
for f in timeline_cfg2_parity.txt timeline_cfg2_fast.txt launches_train.csv; do [ -f $TR/$f ] && cp $TR/$f $OUT/$f; done
digest() {   # <ncu-rep> <out txt> <title>
  ncu -i $1 --page raw --csv > /tmp/_raw.csv 2>/dev/null; ncu -i $1 --page source --csv > /tmp/_src.csv 2>/dev/null
  { echo "# $3"; echo "# ncu --set full --clock-control none --import-source on (one launch); digest by tools/ncu_digest.py"; python tools/ncu_digest.py /tmp/_raw.csv /tmp/_src.csv 25; } > $2
}
[ -f $SW/prof_fused.ncu-rep ] && digest $SW/prof_fused.ncu-rep $OUT/ncu_full_fused_cfg2_parity.txt "nrf_fused_kernel, cfg2 (SmplNerfPipeline 128x128, 64+128), parity mode, algebraic fold on"
for r in $TR/prof_tile_gemm*.ncu-rep; do [ -f $r ] && digest $r $OUT/ncu_full_$(basename $r .ncu-rep).txt "tile_gemm_kernel, one launch of a training step (batch 2048 rays; see Grid Size / duration for which layer)"; done
[ -f $TR/prof_dw_gemm.ncu-rep ] && digest $TR/prof_dw_gemm.ncu-rep $OUT/ncu_full_dw_gemm.txt "dw_gemm_kernel, fine pass 256x256 layer of a training step (batch 2048 rays)"
{ echo "# cuobjdump -sass smpl_nerf_b200/csrc/libnrf_b200.so | grep -c <mnemonic>   (evidence of tcgen05 / TMEM / TMA; B200_PROFILING.md table)"
  cuobjdump -sass smpl_nerf_b200/csrc/libnrf_b200.so > /tmp/_sass.txt
  for m in UTCHMMA UTCBAR LDTM STTM UTMALDG UTMASTG UTMAPF UBLKCP "SYNCS" ; do echo "$m $(grep -c "$m" /tmp/_sass.txt)"; done
  echo "legacy HMMA (mma.sync) $(grep -cE '[^C]HMMA' /tmp/_sass.txt)"
  echo "# per kernel"
  awk '/Function :/ {f=$NF} /UTCHMMA/ {a[f]++} /LDTM/ {b[f]++} /UTMALDG/ {c[f]++} /UTMASTG/ {d[f]++} /UTMAPF/ {e[f]++} END {for (k in a) printf "%s UTCHMMA=%d LDTM=%d UTMALDG=%d UTMASTG=%d UTMAPF=%d\n", k, a[k], b[k], c[k], d[k], e[k]}' /tmp/_sass.txt
} > $OUT/sass_counts.txt
ls -la $OUT
