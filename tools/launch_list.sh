# per-launch durations (+ DRAM bytes) of one training step under ncu; usage: bash tools/launch_list.sh <outdir>
OUT=${1:-gpurun_out/r2d}; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:'tile_gemm|dw_gemm|head_bwd|colsum|bias_grad|dw_reduce|heads_kernel|encode_planes|composite|smpl_points|rayfeat|ray_bias|split_planes|fine_sampling|absmax|scale_from|ray_feats|points_from' \
   -s 1100 -c 175 --csv --log-file $OUT/launches_train.csv python tools/train_profile.py 2048 > $OUT/ncu_launches_train.log 2>&1
tail -3 $OUT/ncu_launches_train.log; wc -l $OUT/launches_train.csv
