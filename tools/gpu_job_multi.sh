#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
FI_EPP_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 8 gpurun_out/pytest_multi.log
N=$(nvidia-smi -L | wc -l)
for X in peer nccl; do
FI_EPP_VERBOSE=1 FI_EPP_EXCHANGE=$X timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --cfg 4 --mode sharded --scale 0.25 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_sharded_${N}_$X.json 2> gpurun_out/bench_sharded_${N}_$X.err; echo "sharded $X rc=$?"; grep -v "^W0\|^\*\*" gpurun_out/bench_sharded_${N}_$X.err | tail -n 4; cut -c1-500 gpurun_out/bench_sharded_${N}_$X.json
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_replicas_$N.json 2> gpurun_out/bench_replicas_$N.err; echo "replicas rc=$?"; tail -n 2 gpurun_out/bench_replicas_$N.err; cut -c1-400 gpurun_out/bench_replicas_$N.json
