#!/bin/bash
# r2k: match_pick with the request's nodes resolved up front; per-kernel times of the device LRU under churn
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log
for pl in "" "--no-pipeline"; do
timeout 300 python bench.py $pl --steps 100 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2k_bench$pl.json 2> gpurun_out/r2k_bench$pl.err
python - "$pl" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2k_bench{sys.argv[1]}.json")); print(sys.argv[1] or "pipeline", "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"], "frac", round(d["roofline"]["frac"],3), d["roofline"].get("stream_ordered"))
PY
done
timeout 300 python bench.py --index-order shuffled --steps 50 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2k_shuf.json 2> gpurun_out/r2k_shuf.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2k_shuf.json")); print("shuffled value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"])
PY
FI_EPP_LIB=$PWD/fusioninfer_b200/lib/libfi_epp_timing.so FI_EPP_VERBOSE=1 timeout 300 python bench.py --no-pipeline --steps 20 --warmup 5 --no-cpu --no-e2e --extras none 2>&1 >/dev/null | grep "match_pick phases" | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lru_|index_' --csv --log-file gpurun_out/r2k_lru_launches.csv python tools/bench_churn.py --steps 3 --no-oracle > gpurun_out/r2k_churn_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2k_lru_launches.csv", errors="ignore")))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            k = d["Kernel Name"].split("(")[0][-40:]
            v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
            v = v / 1e3 if u in ("ns", "nsecond") else v
            a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
for k, (n, t, mx) in agg.items(): print(f"{k:42s} n={n:5d} total={t/1e3:9.3f} ms max={mx:9.1f} us")
PY
