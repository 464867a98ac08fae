#!/bin/bash
# r2t: match_pick with the run-ending lookup under the row reads (default build) vs the committed kernel (libfi_epp_base.so)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
for v in "" "_base" "" "_base"; do
for pl in "" "--no-pipeline"; do
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp$v.so timeout 300 python bench.py $pl --steps 200 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2t_bench$v$pl.json 2> gpurun_out/r2t_bench$v$pl.err
python - "$v$pl" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2t_bench{sys.argv[1]}.json")); print(sys.argv[1] or "new pipeline", "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"])
PY
done
done
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp.so timeout 300 python bench.py --index-order shuffled --steps 50 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2t_shuf.json 2> gpurun_out/r2t_shuf.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2t_shuf.json")); print("shuffled value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"])
PY
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_timing.so FI_EPP_VERBOSE=1 timeout 300 python bench.py --no-pipeline --steps 20 --warmup 5 --no-cpu --no-e2e --extras none 2>&1 >/dev/null | grep "match_pick phases" | tail -1
