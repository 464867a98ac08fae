#!/bin/bash
# round 2, first GPU call: parity of everything on one GPU, then a bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2a_gpu.txt
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 --extras none > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2a_bench.json"))
    print("value %.1fM ms %.4f e2e %.2fM" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), d["roofline"]["kernel_ms"], d.get("parity"))
except Exception as e:
    print("bench parse failed", e)
PY
