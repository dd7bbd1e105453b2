#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests2.log
for c in 3 6 7 8 9; do
  GKD_ISECT_CFG=$c timeout 300 python bench.py --genomes 300 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2b_cfg$c.json 2> gpurun_out/r2b_cfg$c.err
done
tail -8 gpurun_out/r2_tests2.log
for c in 3 6 7 8 9; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_cfg$c.json"))
    print("cfg$c", round(d["value"]), "pairs/s  isect GB/s", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"],3), "isect_ms", round(d["stages"]["intersect_ms"],1))
except Exception as e:
    print("cfg$c failed", e)
PY
done
export GKD_ISECT_CFG=3
timeout 300 python bench.py --genomes 100 --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r2_plain100.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_bucket -s 1 -c 1 -o gpurun_out/r2_prof_isect_cfg3 python bench.py --genomes 100 --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r2_ncu1.log 2>&1
echo "ncu rc=$?"
