#!/bin/bash
# Round-2 evidence capture (run under gpurun, 1 GPU):
#   1. launch list of the bench (every kernel of the graph-replayed steps and of the eval), cold-cache device times
#   2. `ncu --set full` of one launch of every kernel family of the training step and of the ranking eval
# Outputs land in gpurun_out/; summaries are produced afterwards with profiles/summarize_ncu.py / parse_launches.py.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_am_sum.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r02_launches_am_sum.out 2>&1
K='regex:layer_tc2_kernel|wgrad_tc_kernel|wgrad_reduce_kernel|rank_tc_kernel|pack_rows_kernel|gather_fwd_multi_kernel|gather_bwd_multi_kernel|gather_ids_multi_kernel|margin_bwd_multi_kernel|margin_mean_multi_kernel|segment_sum_kernel|radix_hist_kernel|radix_scatter_kernel|radix_row_scan_kernel|colsum_partial_multi_kernel|colsum_finish_multi_kernel|layer_row_kernel|wgrad_row_kernel|pack_weights2_kernel|matrix_sum_multi_kernel|transpose_kernel|prep_queries_kernel|row_inv_norm_kernel|cosine_scores_kernel'
# eager steps (no graph) so that every kernel is a separate launch; skip the warm-up launches, take one step + the eval
timeout 1200 ncu --set full --clock-control none --import-source on -k "$K" -s 400 -c 90 -o gpurun_out/r02_ncu_step \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extra-configs > gpurun_out/r02_ncu_step.out 2>&1
timeout 600 ncu --set full --clock-control none -k 'regex:max_readout_fwd_kernel|max_readout_bwd_kernel' -s 6 -c 4 -o gpurun_out/r02_ncu_max \
  python bench.py --config mutag_max --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extra-configs --no-eval > gpurun_out/r02_ncu_max.out 2>&1
ls -la gpurun_out/r02_ncu_*.ncu-rep gpurun_out/r02_launches_am_sum.csv
tail -2 gpurun_out/r02_ncu_step.out
