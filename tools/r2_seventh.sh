#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "kernel4 or c1_full or skewed or config3 or many_small or fuzz" > gpurun_out/r2_tests7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests7.log
tail -4 gpurun_out/r2_tests7.log
timeout 300 python bench.py --genomes 300 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_b300u.json 2> gpurun_out/r2_b300u.err; echo "bench rc=$?"
grep '^{' gpurun_out/r2_b300u.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), 'isect ms', d['stages']['intersect_ms'], 'frac', d['roofline']['frac'])"
