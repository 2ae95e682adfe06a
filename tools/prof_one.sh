OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_gemm -s ${1:-364} -c 1 -f -o $OUT/prof_tile_gemm \
   python tools/train_profile.py 2048 > $OUT/ncu_full_tile.log 2>&1
ls -la $OUT
