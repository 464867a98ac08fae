#!/bin/bash
# r2q: SM-partitioned pipeline (green contexts): walker on 16 / 24 SMs vs the unpartitioned two-stream pipeline
mkdir -p gpurun_out
FI_EPP_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipelin" 2>&1 | tail -6
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
for part in 16 24 8 0; do
FI_EPP_PIPE_PARTITION=$part timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2q_bench_$part.json 2> gpurun_out/r2q_bench_$part.err
python - "$part" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/r2q_bench_{sys.argv[1]}.json")); print("partition", sys.argv[1], "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"], d["roofline"].get("stream_ordered",{}).get("ms_per_step"))
except Exception as e:
    print("partition", sys.argv[1], "failed", e); import subprocess; print(subprocess.run(["tail","-5",f"gpurun_out/r2q_bench_{sys.argv[1]}.err"],capture_output=True,text=True).stdout)
PY
done
