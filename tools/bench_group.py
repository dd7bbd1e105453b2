#!/usr/bin/env python
"""Config-2 shape through gkd_group_* (several GPUs in ONE process, no torchrun): N synthetic genomes are
block-distributed over the devices, every member builds its slice, the pair matrix is computed with peer
copies of set arenas.  Prints one JSON line; with --verify the result is compared bit for bit with a
single-context run of the same genomes."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genome.distance_b200 as gkd

SEED = 0x5EED0000
RATES = [0.001, 0.01, 0.05, 0.2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--families", type=int, default=10)
    ap.add_argument("--panel", type=int, default=128)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    n, R = a.genomes, a.gpus
    per_fam = max(1, (n + a.families - 1) // a.families)
    host = torch.empty((n, a.length), dtype=torch.uint8, pin_memory=True)
    tmp = torch.empty(a.length, dtype=torch.uint8, device="cuda:0")
    for g in range(n):
        fam, mem = g // per_fam, g % per_fam
        gkd.synth(tmp, SEED, fam, mem, 0.0 if mem == 0 else RATES[mem % 4], device=0)
        host[g].copy_(tmp)
    torch.cuda.synchronize()
    base, extra = divmod(n, R)
    owner = [m for m in range(R) for _ in range(base + (1 if m < extra else 0))]
    best = None
    for rep in range(a.reps):
        with gkd.Group(list(range(R)), k=21, workspace_bytes=min(a.panel, base + 1) * a.length * 16 + (1 << 24),
                       panel=a.panel) as grp:
            t0 = time.perf_counter()
            for g in range(n):
                grp.add(owner[g], host[g].numpy())
            grp.build()
            t1 = time.perf_counter()
            inter, dist = grp.all_vs_all()
            t2 = time.perf_counter()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    ok = None
    if a.verify:
        with gkd.Engine(k=21, device=0) as e:
            for g in range(n):
                e.add(host[g].numpy())
            e.build()
            wi, wd = e.all_vs_all()
        ok = bool(np.array_equal(wi, inter) and np.array_equal(wd, dist))
    pairs = n * (n - 1) // 2
    print(json.dumps({"config": f"gkd_group: {n} x {a.length / 1e6:g} Mbp all-vs-all on {R} GPUs in one process",
                      "pairs": pairs, "total_s": best[0], "add_build_s": best[1], "distance_s": best[2],
                      "pairs_per_s_e2e_from_host_text": pairs / best[0], "pairs_per_s_distance": pairs / best[2],
                      "bit_exact_vs_single_context": ok}), flush=True)
    sys.exit(0 if ok is not False else 1)


if __name__ == "__main__":
    main()
