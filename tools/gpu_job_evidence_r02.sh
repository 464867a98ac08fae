#!/bin/bash
# Round-2 evidence on ONE B200: full GPU test suite, both bench arms, churn (device and host LRU), latency table,
# ncu launch list and one --set full capture of each hot kernel.  Outputs under gpurun_out/ (copied into profiles/).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_r02.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu_r02.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_r02.json
timeout 1200 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_r02.json
timeout 600 python bench.py --no-pipeline --no-cpu --no-e2e --extras none > gpurun_out/bench_r02_stream_ordered.json 2> gpurun_out/bench_r02_stream_ordered.err; echo "stream-ordered rc=$?"; cut -c1-200 gpurun_out/bench_r02_stream_ordered.json
timeout 900 python tools/bench_churn.py --steps 8 --oracle-steps 2 > gpurun_out/churn_r02.json 2> gpurun_out/churn_r02.err; cut -c1-600 gpurun_out/churn_r02.json
timeout 900 python tools/bench_churn.py --steps 5 --no-oracle --lru host > gpurun_out/churn_r02_host_lru.json 2> gpurun_out/churn_r02_host_lru.err; cut -c1-300 gpurun_out/churn_r02_host_lru.json
timeout 600 python tools/bench_latency.py > gpurun_out/latency_r02.json 2> gpurun_out/latency_r02.err; cut -c1-600 gpurun_out/latency_r02.json
CMD="python bench.py --no-pipeline --steps 2 --warmup 3 --no-cpu --no-e2e --extras none --batches 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches_r02.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'hash_blocks_kernel|chain_finalize_kernel|match_pick_kernel' --launch-skip 9 --launch-count 3 -o gpurun_out/prof_r02 -f $CMD > gpurun_out/ncu_full_r02.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/prof_r02.ncu-rep
