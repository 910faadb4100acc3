#!/bin/bash
# Round-2 evidence, second pass (after the TMA loader and the shortened weight preparation): the launch list of the
# bench and `ncu --set full` of the tensor-core kernels and of the kernels that changed since capture_r02.sh ran.
set -x
mkdir -p gpurun_out /tmp/mpqe_ncu
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_final.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r02_launches_final.out 2>&1
MPQE_NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k 'regex:layer_tc2_kernel|wgrad_tc_kernel|pack_weights2_kernel|transpose_kernel|segment_sum_kernel|rank_tc_kernel' \
  -o /tmp/mpqe_ncu/tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs \
  > gpurun_out/r02_ncu_tc_final.out 2>&1
ncu -i /tmp/mpqe_ncu/tc.ncu-rep --page raw --csv > gpurun_out/r02_ncu_tc_final_raw.csv
ncu -i /tmp/mpqe_ncu/tc.ncu-rep --kernel-name regex:layer_tc2_kernel --page source --csv > gpurun_out/r02_ncu_layer_tc2_source_final.csv 2>/dev/null
ls -la /tmp/mpqe_ncu gpurun_out | tail
du -sh gpurun_out
