#!/bin/bash
# 8-GPU session: bit-exact ring check, config-2 strong scaling at N=2/4/8 (torchrun), the single-process group
# (gkd_group_*, peer copies) on 8 GPUs, then the full config-4 run
mkdir -p gpurun_out
P=29531
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P "${@:2}"; }
timeout 600 bash -c "$(declare -f run); P=$P; run 8 tools/check_multi_gpu.py" > gpurun_out/r2f_mgcheck8.log 2>&1; echo "check rc=$?" | tee -a gpurun_out/r2f_mgcheck8.log
tail -2 gpurun_out/r2f_mgcheck8.log
for n in 2 4 8; do
  timeout 900 bash -c "$(declare -f run); P=$P; run $n bench.py --gpus $n --steps 3 --warmup 3" > gpurun_out/r2f_bench_c2_n$n.json 2> gpurun_out/r2f_bench_c2_n$n.err; echo "bench$n rc=$?"
  grep '^{' gpurun_out/r2f_bench_c2_n$n.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$n C2', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d['stages']['wall_ms'])"
done
timeout 900 python tools/bench_group.py --gpus 8 --genomes 1000 --verify > gpurun_out/r2f_group8.json 2> gpurun_out/r2f_group8.err; echo "group rc=$?"; tail -c 700 gpurun_out/r2f_group8.json
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 2000 > gpurun_out/r2f_c4_clocks.csv &
SMI=$!
timeout 1500 bash -c "$(declare -f run); P=$P; run 8 tools/run_c4.py --genomes 20000 --panel 250 --check 1200 --oracle 2" > gpurun_out/r2f_c4_full_8gpu.jsonl 2> gpurun_out/r2f_c4_full_8gpu.err; echo "c4 rc=$?"
kill $SMI
grep '^{' gpurun_out/r2f_c4_full_8gpu.jsonl | tail -1
tail -n 5 gpurun_out/r2f_c4_full_8gpu.err
