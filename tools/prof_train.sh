OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 300 python tools/trace_timeline.py cfg2 > $OUT/timeline_cfg2_parity.txt 2>&1
timeout 300 python tools/trace_timeline.py cfg2 fast > $OUT/timeline_cfg2_fast.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tile_gemm|dw_gemm|head_bwd|colsum|bias_grad|dw_reduce|heads_kernel|encode_planes|composite|smpl_points|rayfeat|ray_bias2|split_planes|fine_sampling|absmax|scale_from|ray_feats' \
   -s 1100 -c 170 --csv --log-file $OUT/launches_train.csv python tools/train_profile.py 2048 > $OUT/ncu_launches_train.log 2>&1
# tile_gemm launches 320.. of the run: one training step has 50 (forward coarse 13, fine 13, backward 12 + 12); capture 4 spread over a step
for skip in 322 337 352 364; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_gemm -s $skip -c 1 -f -o $OUT/prof_tile_gemm_$skip \
   python tools/train_profile.py 2048 > $OUT/ncu_full_tile_$skip.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dw_gemm -s 190 -c 2 -f -o $OUT/prof_dw_gemm \
   python tools/train_profile.py 2048 > $OUT/ncu_full_dw.log 2>&1
ls -la $OUT; du -sh gpurun_out
