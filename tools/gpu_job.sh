#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f kernel_ms=%s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
}
run() { env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err; show gpurun_out/ab.json "$*"; }
run FI_BENCH_PRESORT=1
run FI_EPP_ORDER=1
run FI_EPP_SLICES=1
run FI_EPP_SLICES=2
run FI_EPP_SLICES=3
run FI_EPP_SLICES=4
run FI_EPP_SLICES=2 FI_EPP_ORDER=1
run FI_EPP_SLICES=3 FI_EPP_ORDER=1
run FI_EPP_SLICES=1
FI_EPP_SLICES=2 FI_EPP_TRACE=12 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 >/dev/null | grep "fi_epp trace"
FI_EPP_SLICES=3 FI_EPP_ORDER=1 FI_EPP_TRACE=12 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 >/dev/null | grep "fi_epp trace"
FI_EPP_SLICES=3 FI_EPP_ORDER=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 2>&1 | tail -n 3
