#!/bin/bash
mkdir -p gpurun_out
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_timing.so FI_EPP_VERBOSE=1 timeout 300 python bench.py --no-pipeline --steps 20 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2f_timing.json 2> gpurun_out/r2f_timing.err
grep "match_pick phases" gpurun_out/r2f_timing.err | tail -2
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_timing.so FI_EPP_VERBOSE=1 FI_EPP_MATCH_VEC=4 timeout 300 python bench.py --no-pipeline --steps 20 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2f_timing4.json 2> gpurun_out/r2f_timing4.err
grep "match_pick phases" gpurun_out/r2f_timing4.err | tail -1
python - <<'PY'
import json
for f in ("r2f_timing","r2f_timing4"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"], d["roofline"]["n_probe_per_decision"])
PY
