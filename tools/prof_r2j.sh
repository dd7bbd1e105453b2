#!/bin/bash
# Final round-2 profiling recipe (B200_PROFILING.md), one B200 under gpurun; everything lands in gpurun_out/.
#   1. full GPU test suite   2. the default bench (config 2: value + e2e + cpu_baseline), plain, never under ncu
#   3. ncu launch list of one config-2 step   4. ncu --set full of k_join on the same step
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/r2j_tests.log
timeout 900 python bench.py > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"
if [ "${REF:-0}" = "1" ]; then timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; echo "ref rc=$?"; fi
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r2j_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2j_launches.csv $CMD > gpurun_out/r2j_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $CMD > gpurun_out/r2j_plainb.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_join -s 1 -c 1 -o gpurun_out/r2j_join $CMD > gpurun_out/r2j_ncu_full.log 2>&1
echo "ncu full rc=$?"
grep '^{' gpurun_out/r2j_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d['stages'])"
