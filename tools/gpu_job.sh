#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 900 -k "hash or ragged" > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu7.log
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f kernel_ms=%s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
}
FI_EPP_TRACE=12 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/ov.json 2> gpurun_out/ov.err; show gpurun_out/ov.json overlap; grep "fi_epp trace" gpurun_out/ov.err
