#!/bin/bash
# Round-2 profiling recipe (B200_PROFILING.md), one B200 under gpurun; everything lands in gpurun_out/.
#   1. full GPU test suite
#   2. the default bench (config 2, value + e2e + cpu_baseline) -- plain run, never under ncu
#   3. ncu launch list of one small bench step (shares of the step, cold-cache and serialised)
#   4. ncu --set full of kernel 4 and of the two kernel-3 fast-path kernels
#   5. configs 1, 3, 5 at full size and the gene-sized-records run
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2f_tests.log
tail -3 gpurun_out/r2f_tests.log
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
CMD="python bench.py --genomes 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r2f_plain100.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $CMD > gpurun_out/r2f_plain100b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect_bucket|k_msd_binsort|k_encode_scatter" -s 3 -c 3 -o gpurun_out/r2f_prof_top $CMD > gpurun_out/r2f_ncu_full.log 2>&1
echo "ncu full rc=$?"
for c in c1 c3 c5; do
  timeout 900 python tools/run_config.py $c > gpurun_out/r2f_$c.json 2> gpurun_out/r2f_$c.err; echo "$c rc=$?"
done
timeout 600 python tools/bench_small.py > gpurun_out/r2f_small.json 2> gpurun_out/r2f_small.err; echo "small rc=$?"
grep '^{' gpurun_out/r2f_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d['stages'])"
