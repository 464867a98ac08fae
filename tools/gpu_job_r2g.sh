#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2g_pytest.log; tail -3 gpurun_out/r2g_pytest.log
for pl in "" "--no-pipeline"; do
timeout 300 python bench.py $pl --steps 100 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2g_bench$pl.json 2> gpurun_out/r2g_bench$pl.err
python - "$pl" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2g_bench{sys.argv[1]}.json")); print(sys.argv[1] or "pipeline", "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"], "frac", round(d["roofline"]["frac"],3), d["roofline"].get("stream_ordered"))
PY
done
timeout 300 python bench.py --index-order shuffled --steps 50 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2g_shuf.json 2> gpurun_out/r2g_shuf.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g_shuf.json")); print("shuffled value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"])
PY
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_timing.so FI_EPP_VERBOSE=1 timeout 300 python bench.py --no-pipeline --steps 20 --warmup 5 --no-cpu --no-e2e --extras none 2>&1 >/dev/null | grep "match_pick phases" | tail -1
FI_EPP_VERBOSE=1 timeout 600 python tools/bench_churn.py --steps 6 --no-oracle > gpurun_out/r2g_churn.json 2> gpurun_out/r2g_churn.err; grep add_chains gpurun_out/r2g_churn.err | tail -2; python -c "
import json; d=json.load(open('gpurun_out/r2g_churn.json')); print(d['decisions_per_s'], d['pick_ms'], d['add_ms'], d['lru_threads'])"
