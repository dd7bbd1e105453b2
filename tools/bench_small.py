#!/usr/bin/env python
"""fastaDist on many gene-sized records: GPU (warp-per-pair kernel) vs the CPU port, same inputs."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genome.distance_b200 as gkd
from oracle import oracle as orc

n, length = int(sys.argv[1]) if len(sys.argv) > 1 else 6000, 1500
seqs = []
for g in range(n):
    a = np.empty(length, dtype=np.uint8)
    gkd.synth(a, 77, g % 50, g // 50, 0.0 if g < 50 else 0.03)
    seqs.append(a.tobytes())
with gkd.Engine(k=21) as e:
    for rep in range(2):
        e.reset()
        t0 = time.perf_counter()
        for s in seqs:
            e.add(s)
        e.build()
        t1 = time.perf_counter()
        gi, gd = e.all_vs_all()
        t2 = time.perf_counter()
    m = e.metrics()
pairs = n * (n - 1) // 2
ns = min(n, 1500)
t3 = time.perf_counter()
oi, od = orc.fasta_dist(seqs[:ns], 21, batch=20, threads=0, mode=0)
t4 = time.perf_counter()
cpu_pairs = ns * (ns - 1) // 2
ok = bool(np.array_equal(gd[: ns - 1], od[: ns - 1]))
print(json.dumps({"records": n, "bp": length, "pairs": pairs, "gpu_add_build_s": t1 - t0, "gpu_distance_s": t2 - t1,
                  "gpu_pairs_per_s": pairs / (t2 - t1), "gpu_e2e_pairs_per_s": pairs / (t2 - t0), "intersect_ms": m["intersect_ms"],
                  "cpu_port_pairs_per_s": cpu_pairs / (t4 - t3), "cpu_threads": orc.max_threads(), "cpu_sample_records": ns,
                  "first_row_matches_cpu": ok}))
