#!/bin/bash
# round 2, second GPU call (1 GPU): all GPU tests, the default bench line, pipeline SM-split sweep, churn, latency
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2> gpurun_out/r2b_bench.time
tail -2 gpurun_out/r2b_bench.time
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("%-40s value=%.1fM ms=%.4f %s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k: round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
show gpurun_out/r2b_bench.json default
for combo in "0 0" "3 2" "2 2" "4 2" "6 2" "3 1" "2 3" "8 3"; do
  set -- $combo
  FI_EPP_PIPE_HASH_CTAS=$1 FI_EPP_PIPE_MATCH_CTAS=$2 timeout 300 python bench.py --pipeline --steps 60 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2b_pipe_$1_$2.json 2> gpurun_out/r2b_pipe_$1_$2.err
  show gpurun_out/r2b_pipe_$1_$2.json "pipeline hash_ctas=$1 match_ctas=$2"
done
timeout 600 python tools/bench_churn.py --steps 8 --oracle-steps 1 > gpurun_out/r2b_churn.json 2> gpurun_out/r2b_churn.err; tail -c 1500 gpurun_out/r2b_churn.json; tail -2 gpurun_out/r2b_churn.err
timeout 600 python tools/bench_latency.py --calls 500 > gpurun_out/r2b_latency.json 2> gpurun_out/r2b_latency.err; tail -c 1200 gpurun_out/r2b_latency.json; tail -2 gpurun_out/r2b_latency.err
