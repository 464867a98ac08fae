#!/bin/bash
# r2m: match_pick in runs (confirm lookup overlapped by the row reads), two tickets ahead; BATCH 16 (default) vs 8; churn with the oracle
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2m_pytest.log; tail -4 gpurun_out/r2m_pytest.log
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_b8.so timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pick" 2>&1 | tail -3
for v in "" "_b8"; do
for pl in "" "--no-pipeline"; do
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp$v.so timeout 300 python bench.py $pl --steps 100 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2m_bench$v$pl.json 2> gpurun_out/r2m_bench$v$pl.err
python - "$v$pl" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2m_bench{sys.argv[1]}.json")); print(sys.argv[1] or "default pipeline", "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"], "frac", round(d["roofline"]["frac"],3))
PY
done
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp$v.so timeout 300 python bench.py --index-order shuffled --steps 50 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2m_shuf$v.json 2> gpurun_out/r2m_shuf$v.err
python - "$v" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2m_shuf{sys.argv[1]}.json")); print("shuffled", sys.argv[1], "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"])
PY
done
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_timing.so FI_EPP_VERBOSE=1 timeout 300 python bench.py --no-pipeline --steps 20 --warmup 5 --no-cpu --no-e2e --extras none 2>&1 >/dev/null | grep "match_pick phases" | tail -1
FI_EPP_VERBOSE=1 timeout 900 python tools/bench_churn.py --steps 6 --oracle-steps 1 > gpurun_out/r2m_churn.json 2> gpurun_out/r2m_churn.err; python -c "
import json; d=json.load(open('gpurun_out/r2m_churn.json')); print('churn', d['decisions_per_s'], d['pick_ms'], d['add_ms'], d['index_kernels_ms_per_step'])"
