#!/bin/bash
# r2u: device LRU on sharded pools (gossip round per sub-batch): multi-GPU tests with short timeouts; single-GPU parity + bench
mkdir -p gpurun_out
FI_EPP_VERBOSE=1 timeout 700 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 200 2>&1 | tail -30 > gpurun_out/r2u_pytest_multi.log; tail -5 gpurun_out/r2u_pytest_multi.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 300 2>&1 | tail -8 > gpurun_out/r2u_pytest.log; tail -4 gpurun_out/r2u_pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u_bench.json")); print("value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]), d["roofline"]["stream_ordered"]["ms_per_step"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --no-e2e --extras none > gpurun_out/r2u_bench2.json 2> gpurun_out/r2u_bench2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2u_bench2.json") if l.startswith("{")][-1]); print("2 gpus value %.1fM ms %.4f" % (d["value"]/1e6, d["ms_per_step"]))
PY
