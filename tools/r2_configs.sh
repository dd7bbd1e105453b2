#!/bin/bash
mkdir -p gpurun_out
for c in c1 c3 c5; do
  timeout 900 python tools/run_config.py $c > gpurun_out/r2_$c.json 2> gpurun_out/r2_$c.err; echo "$c rc=$?"; tail -c 1200 gpurun_out/r2_$c.json
done
timeout 600 python tools/bench_small.py > gpurun_out/r2_small.json 2> gpurun_out/r2_small.err; echo "small rc=$?"; tail -c 600 gpurun_out/r2_small.json
