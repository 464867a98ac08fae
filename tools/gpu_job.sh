#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("%s value=%.1fM ms=%.4f kernel_ms=%s" % (sys.argv[2], d["value"]/1e6, d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
}
for v in "" _spre _sboth; do
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp$v.so timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/s$v.json 2> gpurun_out/s$v.err; show gpurun_out/s$v.json "stream$v"
done
