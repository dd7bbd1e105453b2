#!/bin/bash
# bench of experimental warp-cooperative intersect builds (GKD_ISECT_ALGO=warp, GKD_LIB=variant .so)
for lib in ${LIBS:-libgkd.so}; do
  export GKD_LIB=$PWD/genome/distance_b200/$lib
  if [ -z "$NOTEST" ]; then GKD_ISECT_ALGO=warp timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1; fi
  GKD_ISECT_ALGO=warp timeout 300 python bench.py --genomes ${GENOMES:-200} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']; print('$lib warp pairs/s %.0f  isect_ms %.1f  achieved %.0f GB/s  frac %.3f' % (j['value'], r['ms_per_launch'], r['achieved'], r['frac']))"
done
