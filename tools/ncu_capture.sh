#!/bin/bash
# Round-end ncu evidence (run on the GPU box through gpurun; one GPU).  Full-set captures of the dominant kernels of one
# inference step (B=32) and one training step (B=64); the .ncu-rep files stay on the box (tens of MB), the raw metric
# tables come back as CSV under gpurun_out/ and are summarised into profiles/ by tools/ncu_summarise.py.
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm_kernel|conv_c32_kernel|conv0_tc_kernel" -s 46 -c 23 \
  -o /tmp/prof_infer python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --lanes 1 > gpurun_out/ncu_infer.log 2>&1
ncu -i /tmp/prof_infer.ncu-rep --page raw --csv > gpurun_out/ncu_infer_raw.csv 2>> gpurun_out/ncu_infer.log
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"conv_wgrad_kernel|bn_act_bwd_kernel|bn_act_apply_kernel|bn_stats_kernel|conv0_wgrad_kernel" -s 273 -c 91 \
  -o /tmp/prof_train python bench.py --mode train --no-graph --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/ncu_train_raw.csv 2>> gpurun_out/ncu_train.log
wc -c gpurun_out/ncu_infer_raw.csv gpurun_out/ncu_train_raw.csv
tail -2 gpurun_out/ncu_infer.log gpurun_out/ncu_train.log
