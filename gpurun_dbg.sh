timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layer" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_v4.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_v4.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["kernels"], d["gpu_launches"], d["e2e"]["value"], d["clocks"])
PY
