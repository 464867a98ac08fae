#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"]))
PY
}
run() { env $1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pipeline > gpurun_out/ab.json 2> gpurun_out/ab.err; show gpurun_out/ab.json "$1" || tail -5 gpurun_out/ab.err; }
run "A=1"
run "FI_EPP_PIPE_GATE=1"
run "FI_EPP_PIPE_GATE=1 FI_EPP_PIPE_MATCH_CTAS=2"
run "FI_EPP_PIPE_MATCH_CTAS=2"
run "FI_EPP_PIPE_GATE=1 FI_EPP_PIPE_MATCH_CTAS=1"
FI_EPP_PIPE_GATE=1 FI_EPP_PIPE_MATCH_CTAS=2 FI_EPP_TRACE=20 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pipeline 2>&1 >/dev/null | grep -a "fi_epp trace" | head -8
