#!/bin/bash
# Round-1 profiling recipe (B200_PROFILING.md): launch list of one small bench step, then one
# `--set full` capture of the dominant kernel.  Run under gpurun; outputs land in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --genomes ${GENOMES:-64} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_intersect} -c 1 -o gpurun_out/prof_${KERNEL:-k_intersect} $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
tail -c 1500 gpurun_out/plain.log
