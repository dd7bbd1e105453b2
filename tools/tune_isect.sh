#!/bin/bash
# parity tests, then a bench per intersect configuration (GKD_ISECT_CFG)
for c in ${CFGS:-0}; do
  echo "== GKD_ISECT_CFG=$c"
  if [ -z "$NOTEST" ]; then
    GKD_ISECT_CFG=$c timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
  fi
  GKD_ISECT_CFG=$c timeout 300 python bench.py --genomes ${GENOMES:-200} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']; print('pairs/s %.0f  isect_ms %.1f  achieved %.0f GB/s  frac %.3f' % (j['value'], r['ms_per_launch'], r['achieved'], r['frac']))"
done
