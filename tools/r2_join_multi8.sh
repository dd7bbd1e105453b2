#!/bin/bash
# 8-GPU session with the block join: bit-exact ring check (join in the ring vs merge on one GPU), config 2 at N=8,
# then config 4 at full size (20,000 x 5 Mbp, 199,990,000 pairs) with sampled pairs recomputed by the merge kernel
mkdir -p gpurun_out
P=29551
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P "${@:2}"; }
timeout 300 bash -c "$(declare -f run); P=$P; run 8 tools/check_multi_gpu.py" > gpurun_out/j8g_check.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/j8g_check.log
timeout 600 bash -c "$(declare -f run); P=$P; run 8 bench.py --gpus 8 --steps 3 --warmup 3" > gpurun_out/j8g_bench_c2_n8.json 2> gpurun_out/j8g_bench_c2_n8.err; echo "bench8 rc=$?"
grep '^{' gpurun_out/j8g_bench_c2_n8.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=8 C2', round(d['value']), 'e2e', round(d['e2e']['value']), 'step ms', round(d['ms_per_step'],1), d['stages']['intersect_ms'], d['stages']['wall_ms'])"
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 2000 > gpurun_out/j8g_c4_clocks.csv &
SMI=$!
timeout 900 bash -c "$(declare -f run); P=$P; run 8 tools/run_c4.py --genomes 20000 --panel 250 --check 1200 --oracle 2" > gpurun_out/j8g_c4_full_8gpu.jsonl 2> gpurun_out/j8g_c4_full_8gpu.err; echo "c4 rc=$?"
kill $SMI
grep '^{' gpurun_out/j8g_c4_full_8gpu.jsonl | tail -1 | cut -c1-1500
tail -n 3 gpurun_out/j8g_c4_full_8gpu.err
