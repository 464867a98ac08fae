#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f kernel_ms=%s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
}
run() { env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err; show gpurun_out/ab.json "$*"; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "hash or sliced or ragged" 2>&1 | tail -n 3
run FI_EPP_HASH=ldg
run FI_EPP_HASH=tma2
run FI_EPP_HASH=tma3
run FI_EPP_HASH=ldg
run FI_EPP_HASH=tma2
