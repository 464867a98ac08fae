#!/bin/bash
# r2s: bench lines of BASELINE.json's configs 2 and 5 on one GPU; churn long enough to include an index rebuild
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r02.json')); r=d['roofline']; print('value %.1fM' % (d['value']/1e6), d['ms_per_step'], {k:(round(v['decisions_per_s']/1e6,1)) for k,v in r['index_order'].items() if isinstance(v,dict)})"
for c in 2 5; do
timeout 600 python bench.py --cfg $c --no-cpu --steps 200 --warmup 5 > gpurun_out/bench_r02_cfg$c.json 2> gpurun_out/bench_r02_cfg$c.err; echo "cfg$c rc=$?"; python - $c <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_r02_cfg{sys.argv[1]}.json")); print("cfg", sys.argv[1], "value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["kernel_ms"], "e2e %.2fM" % (d["e2e"]["value"]/1e6), d["roofline"].get("stream_ordered",{}).get("ms_per_step"))
PY
done
timeout 900 python tools/bench_churn.py --steps 14 --oracle-steps 1 > gpurun_out/churn_r02_rebuild.json 2> gpurun_out/churn_r02_rebuild.err
python -c "
import json; d=json.load(open('gpurun_out/churn_r02_rebuild.json')); print('churn14', d['decisions_per_s'], d['pick_ms'], d['add_ms'], d['index']['rebuilds'], d['index']['growth'][-3:], d['oracle']['bit_exact'])"
