#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f kernel_ms=%s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
}
run() { env "$1" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e $2 > gpurun_out/ab.json 2> gpurun_out/ab.err; show gpurun_out/ab.json "$*" || tail -5 gpurun_out/ab.err; }
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 2>&1 | tail -n 3
run A=1 ""
run A=2 ""
run A=1 "--cfg 2"
