#!/bin/bash
# round 2: chain-latency microbenchmark, churn with phase timings and a worker sweep, tie test
mkdir -p gpurun_out
nproc > gpurun_out/r2e_host.txt; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" >> gpurun_out/r2e_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2e_host.txt 2>&1
./tools/microbench/chainlat > gpurun_out/r2e_chainlat.txt 2>&1; cat gpurun_out/r2e_chainlat.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tie or add_chains or restatement or label" 2>&1 | tail -3
for t in 0 32 128; do
  FI_EPP_VERBOSE=1 timeout 600 python tools/bench_churn.py --steps 6 --no-oracle --threads $t > gpurun_out/r2e_churn_$t.json 2> gpurun_out/r2e_churn_$t.err
  python - $t <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/r2e_churn_{t}.json"))
    print("threads", t, "dec/s %.0f" % d["decisions_per_s"], "pick", d["pick_ms"], "add", d["add_ms"], "idx kernels ms/step", d["index_kernels_ms_per_step"], d["index_kernel_launches_per_step"])
except Exception as e:
    print("churn", t, "failed", e)
PY
  grep "add_chains" gpurun_out/r2e_churn_$t.err | tail -3
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench.json")); print("bench value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"])
PY
