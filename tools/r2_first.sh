#!/bin/bash
# first round-2 GPU pass: parity, smoke, a reduced-size bench per kernel-4 configuration
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke1.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke1.log
for c in 0 1 2 3 4 5; do
  GKD_ISECT_CFG=$c timeout 300 python bench.py --genomes 300 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_cfg$c.json 2> gpurun_out/r2_cfg$c.err
done
tail -5 gpurun_out/r2_tests1.log; cat gpurun_out/r2_smoke1.log | tail -3
for c in 0 1 2 3 4 5; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_cfg$c.json"))
    print("cfg$c", round(d["value"]), "pairs/s  isect GB/s", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"],3), d["stages"])
except Exception as e:
    print("cfg$c failed", e)
PY
done
