#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "group or every_visible or fastadist_report" > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests4.log
tail -25 gpurun_out/r2_tests4.log
