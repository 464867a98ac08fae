#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 -k "not fullsize" > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu15.log
timeout 900 python tools/bench_churn.py > gpurun_out/churn_r01.json 2> gpurun_out/churn_r01.err; cat gpurun_out/churn_r01.json | cut -c1-900
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f kernel_ms=%s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
}
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/t.json 2> gpurun_out/t.err; show gpurun_out/t.json after_lora
