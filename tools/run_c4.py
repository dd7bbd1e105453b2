#!/usr/bin/env python
"""BASELINE config 4 shape (N synthetic 5 Mbp genomes all-vs-all, tile-sharded over the GPUs of one
box) through the streamed column-panel path: no rank ever holds more than its own slice of the sets
plus one sub-panel, so the total set volume may exceed one GPU's HBM.

  torchrun --nproc-per-node 8 tools/run_c4.py --genomes 8000 --panel 250

Prints one JSON line on rank 0 (pairs/s aggregate, per-phase wall times, a parity spot check)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genome.distance_b200 as gkd
from genome.distance_b200 import sharding

SEED = 0x5EED0000
RATES = [0.001, 0.01, 0.05, 0.2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=8000)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--families", type=int, default=100)
    ap.add_argument("--panel", type=int, default=250)
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = a.genomes
    per = max(1, (n + a.families - 1) // a.families)
    mine = sharding.genome_slice(n, world, rank)
    buf = torch.empty(a.length, dtype=torch.uint8, device=dev)
    eng = gkd.Engine(k=21, device=local)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for g in mine:
        fam, mem = g // per, g % per
        gkd.synth(buf, SEED, fam, mem, 0.0 if mem == 0 else RATES[mem % 4], device=local)
        eng.add(buf)
    eng.build()
    torch.cuda.synchronize(); dist.barrier()
    t1 = time.perf_counter()
    gi, gj, inter, d = sharding.streamed_all_vs_all(eng, n, world, rank, dev, panel_genomes=a.panel)
    torch.cuda.synchronize()
    t_mine = time.perf_counter() - t1
    dist.barrier()
    t2 = time.perf_counter()
    # parity spot check: recompute a few of this rank's pairs that involve only its own genomes
    ok = True
    if len(mine) >= 2:
        i2, u2, d2 = eng.pair(0, 1)
        sel = np.where((gi == mine[0]) & (gj == mine[1]))[0]
        ok = len(sel) == 1 and int(inter[sel[0]]) == i2 and d[sel[0]] == d2
    stats = torch.tensor([float(len(gi)), t_mine, 1.0 if ok else 0.0, float((d < 1.0).sum())], dtype=torch.float64, device=dev)
    tmax = stats.clone()
    dist.all_reduce(stats)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_pairs = n * (n - 1) // 2
    if rank == 0:
        print(json.dumps({"config": f"c4-shape: {n} x {a.length / 1e6:g} Mbp all-vs-all, streamed panels of {a.panel}",
                          "n_gpus": world, "pairs": total_pairs, "pairs_computed": int(stats[0]),
                          "set_bytes_total": int(n * 8 * a.length), "build_s": t1 - t0, "distance_s": t2 - t1,
                          "slowest_rank_distance_s": float(tmax[1]), "pairs_per_s": total_pairs / (t2 - t1),
                          "related_pairs": int(stats[3]), "spot_check_ok": bool(stats[2] == world)}), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
