#!/bin/bash
# 8-GPU: ring parity check + config 2 at N=8
mkdir -p gpurun_out
P=29561
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P "${@:2}"; }
timeout 300 bash -c "$(declare -f run); P=$P; run 8 tools/check_multi_gpu.py" > gpurun_out/jn8_check.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/jn8_check.log
timeout 600 bash -c "$(declare -f run); P=$P; run 8 bench.py --gpus 8 --steps 3 --warmup 3" > gpurun_out/jn8_bench.json 2> gpurun_out/jn8_bench.err; echo "bench8 rc=$?"
grep '^{' gpurun_out/jn8_bench.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=8 C2', round(d['value']), 'e2e', round(d['e2e']['value']), 'step ms', round(d['ms_per_step'],1), d['stages']['intersect_ms'], d['stages']['wall_ms'])"
