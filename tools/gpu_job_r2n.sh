#!/bin/bash
# r2n: probe pass (chain_probe_kernel: run of consecutive nodes per request, cost bins) + match_pick on the records, longest first
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2n_pytest.log; tail -4 gpurun_out/r2n_pytest.log
for pl in "" "--no-pipeline"; do
timeout 300 python bench.py $pl --steps 100 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2n_bench$pl.json 2> gpurun_out/r2n_bench$pl.err
python - "$pl" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2n_bench{sys.argv[1]}.json")); print(sys.argv[1] or "pipeline", "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"], "frac", round(d["roofline"]["frac"],3))
PY
done
timeout 300 python bench.py --index-order shuffled --steps 50 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2n_shuf.json 2> gpurun_out/r2n_shuf.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2n_shuf.json")); print("shuffled value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"])
PY
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_timing.so FI_EPP_VERBOSE=1 timeout 300 python bench.py --no-pipeline --steps 20 --warmup 5 --no-cpu --no-e2e --extras none 2>&1 >/dev/null | grep "match_pick phases" | tail -1
CMD="python bench.py --no-pipeline --steps 2 --warmup 3 --no-cpu --no-e2e --extras none --batches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hash_blocks|chain_|match_pick' --launch-skip 12 --launch-count 8 --csv --log-file gpurun_out/r2n_launches.csv $CMD > /dev/null 2>&1; grep -a "gpu__time_duration" gpurun_out/r2n_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
