#!/bin/bash
# ncu evidence (run on the GPU box through gpurun; one GPU).  `--set full` captures of ONE step of each workload
# (tools/ncu_targets.py brackets it with cudaProfilerStart/Stop); the .ncu-rep files stay on the box, the raw metric tables come
# back as CSV under gpurun_out/ and are summarised into profiles/ by tools/ncu_summarise.py.
#   tools/ncu_capture.sh [infer] [strict] [train] [mobilenet]
set -u
mkdir -p gpurun_out
for what in "$@"; do
  extra=""
  # one training step has ~300 launches: keep the kernels the step's time is made of
  if [ "$what" = "train" ]; then extra='-k regex:conv_igemm_kernel|conv_wgrad_kernel|conv0_wgrad_kernel|conv0_tc_kernel|conv_c32_kernel|bn_act_bwd_kernel|bn_act_apply_kernel|bn_stats_kernel|region_|unpack_wgrad|pack_weight'; fi
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off $extra \
    -o /tmp/prof_$what -f python tools/ncu_targets.py $what > gpurun_out/ncu_$what.log 2>&1
  ncu -i /tmp/prof_$what.ncu-rep --page raw --csv > gpurun_out/ncu_${what}_raw.csv 2>> gpurun_out/ncu_$what.log
  tail -1 gpurun_out/ncu_$what.log
done
wc -c gpurun_out/ncu_*_raw.csv
