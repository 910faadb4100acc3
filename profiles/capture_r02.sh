#!/bin/bash
# Round-2 evidence capture (run under gpurun, 1 GPU):
#   1. launch list of the bench (the graph-replayed steps and the eval), cold-cache device times
#   2. `ncu --set full` of EVERY kernel of one eager training step and one eager ranking batch: bench.py brackets them
#      with cudaProfilerStart/Stop when MPQE_NCU_RANGE is set
#   3. the same for the max readout kernels (MUTAG-shaped MPQE-max)
# The reports stay on the box (a full-set report of ~110 launches exceeds what gpurun copies back); what comes back
# in gpurun_out/ are their raw pages as CSV (every metric of every launch) and the source page of the layer kernel.
# Summaries are produced afterwards with profiles/summarize_ncu.py / parse_launches.py.
set -x
mkdir -p gpurun_out /tmp/mpqe_ncu
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_am_sum.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r02_launches_am_sum.out 2>&1
MPQE_NCU_RANGE=1 timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o /tmp/mpqe_ncu/step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs \
  > gpurun_out/r02_ncu_step.out 2>&1
ncu -i /tmp/mpqe_ncu/step.ncu-rep --page raw --csv > gpurun_out/r02_ncu_step_raw.csv
ncu -i /tmp/mpqe_ncu/step.ncu-rep --kernel-name regex:layer_tc2_kernel --page source --csv > gpurun_out/r02_ncu_layer_tc2_source.csv 2>/dev/null
MPQE_NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k 'regex:max_readout_fwd_kernel|max_readout_bwd_kernel' -o /tmp/mpqe_ncu/max -f \
  python bench.py --config mutag_max --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs --no-eval \
  > gpurun_out/r02_ncu_max.out 2>&1
ncu -i /tmp/mpqe_ncu/max.ncu-rep --page raw --csv > gpurun_out/r02_ncu_max_raw.csv
ls -la /tmp/mpqe_ncu gpurun_out | tail -20
du -sh gpurun_out
