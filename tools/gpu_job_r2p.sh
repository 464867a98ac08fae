#!/bin/bash
# round 2, multi-GPU evidence: `gpurun --gpus N -- bash tools/gpu_job_r2p.sh N [full]`
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2p_topo_$N.txt 2>&1
if [ "$2" = "full" ]; then SEL=""; else SEL="-k not(full_size)"; fi
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q $SEL 2>&1 | tail -40 > gpurun_out/r2p_pytest_$N.log
tail -4 gpurun_out/r2p_pytest_$N.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2p_bench_$N.json 2> gpurun_out/r2p_bench_$N.err ) 2> gpurun_out/r2p_bench_$N.time
tail -3 gpurun_out/r2p_bench_$N.time
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/r2p_bench_{N}.json") if l.startswith("{")][-1])
    print("replicas value=%.1fM ms=%.4f e2e=%.2fM" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), d["e2e"]["roofline"])
    for k,v in (d["roofline"].get("sharded") or {}).items():
        print(k, "%.1fM ms=%.4f" % (v["decisions_per_s"]/1e6, v["ms_per_step"]), {a: round(b,4) for a,b in v["kernel_ms"].items()}, v.get("parity"), v["exchange"])
        for n2,vv in v.get("variants",{}).items():
            print("   ", n2, ("%.1fM ms=%.4f" % (vv["decisions_per_s"]/1e6, vv["ms_per_step"])) if "ms_per_step" in vv else vv)
except Exception as e:
    print("parse failed", e)
    import subprocess; print(subprocess.run(["tail","-30",f"gpurun_out/r2p_bench_{N}.err"],capture_output=True,text=True).stdout)
PY
