#!/bin/bash
# r2w: replicas at N ranks, pipelined: does every rank get its SM partition?  (FI_EPP_VERBOSE prints one line per rank)
N=${1:-8}
mkdir -p gpurun_out
FI_EPP_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 400 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2w_bench$N.json 2> gpurun_out/r2w_bench$N.err
grep -a "SM partition" gpurun_out/r2w_bench$N.err | sort | uniq -c
python - $N <<'PY'
import json,sys
d=json.loads([l for l in open(f"gpurun_out/r2w_bench{sys.argv[1]}.json") if l.startswith("{")][-1]); print(sys.argv[1], "gpus value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["config"]["pipeline"][-60:], d["roofline"]["stream_ordered"]["ms_per_step"])
PY
FI_EPP_PIPE_PARTITION=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 400 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2w_bench${N}_nopart.json 2> gpurun_out/r2w_bench${N}_nopart.err
python - $N <<'PY'
import json,sys
d=json.loads([l for l in open(f"gpurun_out/r2w_bench{sys.argv[1]}_nopart.json") if l.startswith("{")][-1]); print(sys.argv[1], "gpus unpartitioned value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]))
PY
