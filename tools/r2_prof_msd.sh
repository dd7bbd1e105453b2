#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --genomes 100 --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/r2_plain_msd.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_msd_binsort|k_encode_scatter" -s 2 -c 2 -o gpurun_out/r2_prof_msd2 $CMD > gpurun_out/r2_ncu_msd.log 2>&1
echo "ncu rc=$?"
