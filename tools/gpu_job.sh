#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -k "lru or fullsize or multi or limits or device_resident" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu2.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
tail -n 30 gpurun_out/pytest_gpu2.log; tail -n 20 gpurun_out/bench1.err; cat gpurun_out/bench1.json; tail -n 5 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
