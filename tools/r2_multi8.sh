#!/bin/bash
# 8-GPU session: bit-exact ring check, config-2 strong-scaling bench at N=8, then the full config-4 run
mkdir -p gpurun_out
P=29531
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P "$@"; }
timeout 600 bash -c "$(declare -f run); P=$P; run tools/check_multi_gpu.py" > gpurun_out/r2_mgcheck8.log 2>&1; echo "check rc=$?" | tee -a gpurun_out/r2_mgcheck8.log
tail -2 gpurun_out/r2_mgcheck8.log
timeout 900 bash -c "$(declare -f run); P=$P; run bench.py --gpus 8 --steps 3 --warmup 3" > gpurun_out/r2_bench_c2_n8.json 2> gpurun_out/r2_bench_c2_n8.err; echo "bench8 rc=$?"
grep '^{' gpurun_out/r2_bench_c2_n8.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=8 C2', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d['stages']['wall_ms'])"
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 2000 > gpurun_out/r2_c4_clocks.csv &
SMI=$!
timeout 1500 bash -c "$(declare -f run); P=$P; run tools/run_c4.py --genomes 20000 --panel 250 --check 1024 --oracle 2" > gpurun_out/r2_c4_full_8gpu.jsonl 2> gpurun_out/r2_c4_full_8gpu.err; echo "c4 rc=$?"
kill $SMI
grep '^{' gpurun_out/r2_c4_full_8gpu.jsonl | tail -1
tail -n 5 gpurun_out/r2_c4_full_8gpu.err
