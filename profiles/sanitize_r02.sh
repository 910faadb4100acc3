#!/bin/bash
# compute-sanitizer passes over the library (run under gpurun, 1 GPU).  memcheck: out-of-bounds / misaligned accesses of
# every kernel family at small shapes; racecheck: shared-memory hazards of the kernels that stage through shared memory.
# Logs: gpurun_out/r02_sanitizer_*.log; the summary lines are copied to profiles/r02_sanitizer.txt.
set -x
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $S --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" \
  > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
timeout 2400 $S --tool memcheck --error-exitcode 7 python -m pytest -q -x -m gpu tests/test_gpu_kernels.py \
  -k "not 4096 and not 1000" > gpurun_out/r02_sanitizer_memcheck_kernels.log 2>&1; echo "memcheck kernels rc=$?"
timeout 1500 $S --tool memcheck --error-exitcode 7 python -m pytest -q -x -m gpu tests/test_gpu_zz_peer_rows.py tests/test_gpu_optim.py \
  > gpurun_out/r02_sanitizer_memcheck_peer_optim.log 2>&1; echo "memcheck peer+optim rc=$?"
timeout 1500 $S --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" \
  > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
for f in gpurun_out/r02_sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error" $f | tail -5; done
