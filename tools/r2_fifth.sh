#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "kernel3_paths or dna_sets_bit_exact or protein_sets or fuzz" > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests5.log
tail -30 gpurun_out/r2_tests5.log
timeout 300 python bench.py --genomes 300 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_msd_300.json 2> gpurun_out/r2_msd_300.err; echo "bench rc=$?"
grep '^{' gpurun_out/r2_msd_300.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), d['stages'])"
tail -3 gpurun_out/r2_msd_300.err
