#!/bin/bash
mkdir -p gpurun_out
N=2
for X in peer nccl; do
FI_EPP_TRACE=14 FI_EPP_EXCHANGE=$X timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --cfg 4 --mode sharded --scale 0.25 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/trace_$X.json 2> gpurun_out/trace_$X.err; echo "sharded $X rc=$?"; grep "fi_epp trace" gpurun_out/trace_$X.err
done
