#!/bin/bash
# r2y: which of the two kernels that share the big partition should the block scheduler prefer?
mkdir -p gpurun_out
for cfg in "match 0" "hash 0" "none 0" "match 4" "hash 4"; do
set -- $cfg
FI_EPP_PIPE_PRIO=$1 FI_EPP_PIPE_HASH_CTAS=$2 timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2y_$1_$2.json 2> gpurun_out/r2y_$1_$2.err
python - $1 $2 <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2y_{sys.argv[1]}_{sys.argv[2]}.json")); print("prio", sys.argv[1], "hash_ctas", sys.argv[2], "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]))
PY
done
