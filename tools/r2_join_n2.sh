#!/bin/bash
# 2-GPU session: ring with the block join bit-exact against a single-GPU merge run, then config 2 at 400 genomes, N=1 and N=2
mkdir -p gpurun_out
P=29541
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P "${@:2}"; }
timeout 300 bash -c "$(declare -f run); P=$P; run 2 tools/check_multi_gpu.py" > gpurun_out/jn2_check.log 2>&1; echo "check rc=$?"; tail -2 gpurun_out/jn2_check.log
timeout 300 python bench.py --genomes ${G:-600} --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/jn2_b_n1.json 2> gpurun_out/jn2_b_n1.err; echo "n1 rc=$?"
timeout 300 bash -c "$(declare -f run); P=$P; run 2 bench.py --gpus 2 --genomes ${G:-600} --steps 3 --warmup 3" > gpurun_out/jn2_b_n2.json 2> gpurun_out/jn2_b_n2.err; echo "n2 rc=$?"
for f in gpurun_out/jn2_b_n1.json gpurun_out/jn2_b_n2.json; do
grep '^{' $f | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), 'e2e', round(d['e2e']['value']), 'step ms', round(d['ms_per_step'],1), d['stages']['intersect_ms'], d['stages']['wall_ms'], d['roofline']['kernel'][:10])"
done
