#!/bin/bash
# round-1 evidence run: full gpu suite, bench (both arms), ncu launch list + full capture of the hot kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_r01.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_r01.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; echo "ref rc=$?"
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
cat gpurun_out/bench_ref_r01.json gpurun_out/bench_r01.json | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --batches 1 > gpurun_out/ncu_launches_r01.log 2>&1; echo "launchlist rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'hash_blocks_kernel|chain_finalize_kernel|match_pick_kernel' -s 6 -c 3 -o gpurun_out/prof_r01 -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --batches 1 > gpurun_out/ncu_full_r01.log 2>&1; echo "full rc=$?"
