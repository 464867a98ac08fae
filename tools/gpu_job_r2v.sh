#!/bin/bash
# r2v: host cost of a pipelined submit (13 runtime calls), 1 and 4 ranks
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "pipelin or lru" 2>&1 | tail -3
timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2v_bench1.json 2> gpurun_out/r2v_bench1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2v_bench1.json")); print("1 gpu value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["config"]["pipeline"][-60:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 400 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2v_bench$N.json 2> gpurun_out/r2v_bench$N.err
python - $N <<'PY'
import json,sys
d=json.loads([l for l in open(f"gpurun_out/r2v_bench{sys.argv[1]}.json") if l.startswith("{")][-1]); print(sys.argv[1], "gpus value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["config"]["pipeline"][-60:])
PY
