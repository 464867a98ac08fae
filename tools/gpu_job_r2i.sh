#!/bin/bash
# r2i: device-resident LRU — parity tests, churn benchmark (device vs host LRU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lru or add_chains or churn" 2>&1 | tail -15 > gpurun_out/r2i_pytest_lru.log; tail -5 gpurun_out/r2i_pytest_lru.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2i_pytest.log; tail -3 gpurun_out/r2i_pytest.log
FI_EPP_VERBOSE=1 timeout 900 python tools/bench_churn.py --steps 8 --oracle-steps 1 > gpurun_out/r2i_churn_device.json 2> gpurun_out/r2i_churn_device.err; echo "churn device rc=$?"; grep "device LRU" gpurun_out/r2i_churn_device.err | tail -2
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2i_churn_device.json')); print("device", d['decisions_per_s'], d['pick_ms'], d['add_ms'], d['index_kernels_ms_per_step'], d['index_kernel_launches_per_step'], d['index'], d['oracle'])
except Exception as e: print("no json", e)
PY
tail -3 gpurun_out/r2i_churn_device.err
timeout 600 python tools/bench_churn.py --steps 5 --no-oracle --lru host > gpurun_out/r2i_churn_host.json 2> gpurun_out/r2i_churn_host.err; echo "churn host rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2i_churn_host.json')); print("host", d['decisions_per_s'], d['pick_ms'], d['add_ms'])
except Exception as e: print("no json", e)
PY
