#!/bin/bash
# r2z: final defaults (partition 40 SMs, 4 persistent hashing CTAs per SM): parity of the pipelined paths + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "pipelin or lru or hash" 2>&1 | tail -3
timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2z.json 2> gpurun_out/r2z.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2z.json")); print("default value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["stream_ordered"]["ms_per_step"])
PY
FI_EPP_PIPE_HASH_CTAS=1000 timeout 200 python bench.py --steps 300 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2z_uncapped.json 2> gpurun_out/r2z_uncapped.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2z_uncapped.json")); print("uncapped value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]))
PY
