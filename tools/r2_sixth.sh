#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests6.log
tail -15 gpurun_out/r2_tests6.log
timeout 300 python bench.py --genomes 300 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_b300.json 2> gpurun_out/r2_b300.err; echo "bench rc=$?"
grep '^{' gpurun_out/r2_b300.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), d['stages'])"
