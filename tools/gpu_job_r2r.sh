#!/bin/bash
# r2r: SM partition sweep (walker SMs, walker shape); two repetitions each to see the run-to-run spread
mkdir -p gpurun_out
for cfg in "40 0" "40 1" "48 0" "56 0" "32 0" "24 0" "40 0" "48 0" "0 0"; do
set -- $cfg
FI_EPP_PIPE_PARTITION=$1 FI_EPP_WALK_COMPACT=$2 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2r_bench_$1_$2.json 2> gpurun_out/r2r_bench_$1_$2.err
python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/r2r_bench_{sys.argv[1]}_{sys.argv[2]}.json")); print("partition", sys.argv[1], "compact", sys.argv[2], "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["config"]["pipeline"][:60])
except Exception as e:
    print("failed", sys.argv[1:], e)
PY
done
