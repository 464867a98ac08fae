#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -n 4
for S in 8 1 4 16 8; do
FI_EPP_FEED_SLICES=$S timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/feed_$S.json 2> gpurun_out/feed_$S.err
python - gpurun_out/feed_$S.json $S <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("feed_slices", sys.argv[2], "value %.1fM" % (d["value"]/1e6), "e2e %.3fM/s  %.3f ms/step" % (d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), "parity", d.get("parity"))
PY
done
