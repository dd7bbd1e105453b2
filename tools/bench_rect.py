#!/usr/bin/env python
"""Times gkd_query_vs_ref on a rows x cols block of 5 Mbp genomes (the shape one rank of the 8-GPU ring computes),
for a few block-join geometries (GKD_JOIN_CFG is read when a context is created)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genome.distance_b200 as gkd


def main():
    rows, cols, length = int(sys.argv[1]), int(sys.argv[2]), 5_000_000
    cfgs = sys.argv[3:] or ["auto"]
    n = rows + cols
    buf = torch.empty((100, length), dtype=torch.uint8, device="cuda")
    for cfg in cfgs:
        if cfg == "auto":
            os.environ.pop("GKD_JOIN_CFG", None)
        else:
            os.environ["GKD_JOIN_CFG"] = cfg
        with gkd.Engine(k=21) as e:
            for g in range(n):
                gkd.synth(buf[g % 100], 0x5EED0000, g // 100, g % 100, [0.001, 0.01, 0.05, 0.2][g % 4] if g % 100 else 0.0)
                e.add(buf[g % 100])
                if g % 100 == 99:
                    e.build()
            e.build()
            q, r = np.arange(rows, dtype=np.uint32), np.arange(rows, n, dtype=np.uint32)
            e.query_vs_ref(q, r)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            inter, d = e.query_vs_ref(q, r)
            dt = time.perf_counter() - t0
            m = e.metrics()
            print(json.dumps({"rows": rows, "cols": cols, "cfg": cfg, "wall_ms": 1e3 * dt, "intersect_ms": m["intersect_ms"],
                              "pairs_per_s": rows * cols / dt, "kernel": m["intersect_kernel"], "inter_sum": int(inter.sum())}))


if __name__ == "__main__":
    main()
