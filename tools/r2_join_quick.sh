#!/bin/bash
# quick check of a block-join change: parity, then config 2 at 300 and 1000 genomes (kernel-only lines)
mkdir -p gpurun_out
T=${1:-jq}
timeout 900 python -m pytest tests/test_gpu_join.py -x -q > gpurun_out/${T}_tests.log 2>&1; echo "join tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
run() { # name, genomes, env...
  local name=$1 n=$2; shift 2
  env "$@" timeout 300 python bench.py --genomes $n --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; echo "$name rc=$?"
  grep '^{' gpurun_out/${T}_$name.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ', round(d['value']), 'pairs/s; step ms', round(d['ms_per_step'],1), 'k4 ms', round(d['stages']['intersect_ms'],2), d['roofline']['kernel'][:12])"
}
run b300 300 X=1
run b1000 1000 X=1
for extra in "$@"; do
  [ "$extra" = "$T" ] && continue
  run "b1000_${extra//=/_}" 1000 $extra
done
