#!/bin/bash
# 2-GPU validation of the panel ring: bit-exact check, reduced bench at N=2, a small config-4-shaped run
mkdir -p gpurun_out
P=29517
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P tools/check_multi_gpu.py > gpurun_out/r2_mgcheck2.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_mgcheck2.log
tail -3 gpurun_out/r2_mgcheck2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --genomes 400 --steps 2 --warmup 2 > gpurun_out/r2_bench_n2_400.json 2> gpurun_out/r2_bench_n2_400.err; echo "bench2 rc=$?"
timeout 300 python bench.py --genomes 400 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_n1_400.json 2> gpurun_out/r2_bench_n1_400.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P tools/run_c4.py --genomes 1200 --panel 100 --check 64 > gpurun_out/r2_c4_small_n2.json 2> gpurun_out/r2_c4_small_n2.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n1_400","r2_bench_n2_400"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, "frac", round(d["roofline"]["frac"],3), d["stages"]["wall_ms"])
    except Exception as e:
        print(f, "failed", e)
try:
    print(open("gpurun_out/r2_c4_small_n2.json").read())
except Exception as e:
    print(e)
PY
tail -5 gpurun_out/r2_c4_small_n2.err gpurun_out/r2_bench_n2_400.err
