#!/bin/bash
mkdir -p gpurun_out
for V in "FI_EPP_COPY_STREAMS=1" "FI_EPP_COPY_STREAMS=2" "FI_EPP_COPY_STREAMS=2 FI_EPP_FEED_SLICES=16" "FI_EPP_COPY_STREAMS=1"; do
env $V timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/c.json 2> gpurun_out/c.err
python - gpurun_out/c.json "$V" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[2], "value %.1fM" % (d["value"]/1e6), "e2e %.3fM/s  %.3f ms/step" % (d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]))
PY
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 2>&1 | tail -n 2
