#!/bin/bash
# first GPU session of the block-join kernel: parity (join vs merge vs oracle), then a geometry sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_join.py -x -q > gpurun_out/j1_tests.log 2>&1; echo "join tests rc=$?"; tail -15 gpurun_out/j1_tests.log
B="python bench.py --genomes 300 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
GKD_ISECT_ALGO=merge timeout 300 $B > gpurun_out/j1_b300_merge.json 2> gpurun_out/j1_b300_merge.err; echo "merge rc=$?"
for c in 0 1 2 3 4; do
  GKD_JOIN_CFG=$c timeout 300 $B > gpurun_out/j1_b300_cfg$c.json 2> gpurun_out/j1_b300_cfg$c.err; echo "cfg$c rc=$?"
done
for f in 20 40; do
  GKD_JOIN_FILL=$f timeout 300 $B > gpurun_out/j1_b300_fill$f.json 2> gpurun_out/j1_b300_fill$f.err; echo "fill$f rc=$?"
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j1_b1000.json 2> gpurun_out/j1_b1000.err; echo "b1000 rc=$?"
for f in gpurun_out/j1_b*.json; do
  grep '^{' $f | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$f', round(d['value']), 'ms', round(d['ms_per_step'],1), 'k4 ms', round(d['stages']['intersect_ms'],2), d['roofline']['kernel'][:30], 'e2e', d.get('e2e',{}).get('value'))" 2>/dev/null || echo "$f: no line"
done
