# usage: bash tools/prof_one.sh <skip> [<skip> ...]   one ncu --set full capture per listed tile_gemm launch index of tools/train_profile.py
OUT=gpurun_out/r2c; mkdir -p $OUT
for skip in "$@"; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_gemm -s $skip -c 1 -f -o $OUT/prof_tile_gemm_$skip \
   python tools/train_profile.py 2048 > $OUT/ncu_full_tile_$skip.log 2>&1
done
ls -la $OUT
