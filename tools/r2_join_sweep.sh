#!/bin/bash
# block-join geometry sweep: GKD_JOIN_CFG x {300, 1000 genomes}; parity first
mkdir -p gpurun_out
T=${1:-j2}
timeout 900 python -m pytest tests/test_gpu_join.py -x -q > gpurun_out/${T}_tests.log 2>&1; echo "join tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
for n in 300 1000; do
  for c in ${CFGS:-0 1 2 3 4 5 6 7}; do
    GKD_JOIN_CFG=$c timeout 300 python bench.py --genomes $n --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_b${n}_cfg$c.json 2> gpurun_out/${T}_b${n}_cfg$c.err; echo "n=$n cfg$c rc=$?"
    grep '^{' gpurun_out/${T}_b${n}_cfg$c.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ', round(d['value']), 'pairs/s; step ms', round(d['ms_per_step'],1), 'k4 ms', round(d['stages']['intersect_ms'],2), d['roofline']['kernel'][:12])"
  done
done
