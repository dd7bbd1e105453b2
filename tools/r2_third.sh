#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
tail -15 gpurun_out/r2_tests3.log
timeout 600 python tools/bench_small.py > gpurun_out/r2_small2.json 2> gpurun_out/r2_small2.err; echo "small rc=$?"; tail -c 600 gpurun_out/r2_small2.json
