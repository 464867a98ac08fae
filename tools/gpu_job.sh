#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 900 -k "not fullsize and not multi" > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu12.log
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f kernel_ms=%s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
}
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/t.json 2> gpurun_out/t.err; show gpurun_out/t.json match_trim
