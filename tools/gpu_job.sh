#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --batches 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'match_pick_kernel' --launch-skip 3 --launch-count 1 -o gpurun_out/prof_match_nodes -f $CMD > gpurun_out/ncu_match_nodes.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/prof_match_nodes.ncu-rep
