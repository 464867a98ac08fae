#!/bin/bash
# r2x: the two co-running kernels of the big partition: CTAs per SM of hash_blocks (0 = one CTA per request) / match_pick (0 = 3)
mkdir -p gpurun_out
for cfg in "0 0" "2 0" "4 0" "6 0" "0 2" "4 2"; do
set -- $cfg
FI_EPP_PIPE_HASH_CTAS=$1 FI_EPP_PIPE_MATCH_CTAS=$2 timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2x_$1_$2.json 2> gpurun_out/r2x_$1_$2.err
python - $1 $2 <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2x_{sys.argv[1]}_{sys.argv[2]}.json")); print("hash_ctas", sys.argv[1], "match_ctas", sys.argv[2], "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]))
PY
done
