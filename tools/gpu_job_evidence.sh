#!/bin/bash
# Round evidence on ONE B200: full GPU test suite, both bench arms, churn bench, ncu launch list and one
# --set full capture of each hot kernel.  Outputs under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_r01.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu_r01.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_r01.json
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_r01.json
timeout 600 python bench.py --pipeline --no-cpu --no-e2e > gpurun_out/bench_r01_pipeline.json 2> gpurun_out/bench_r01_pipeline.err; echo "pipeline rc=$?"; cut -c1-200 gpurun_out/bench_r01_pipeline.json
timeout 600 python tools/bench_churn.py > gpurun_out/churn_r01.json 2> gpurun_out/churn_r01.err; cut -c1-400 gpurun_out/churn_r01.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --batches 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches_r01.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'hash_blocks_kernel|chain_finalize_kernel|match_pick_kernel' --launch-skip 9 --launch-count 3 -o gpurun_out/prof_r01 -f $CMD > gpurun_out/ncu_full_r01.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/prof_r01.ncu-rep
