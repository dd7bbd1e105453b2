#!/usr/bin/env python
"""torchrun check of the N>1 path on real GPUs: every rank builds its genome slice, the panel ring
(sharding.ring_all_vs_all) moves the peers' sets over NCCL and adopts them in place, each rank computes
its rank blocks; the union of the results must equal a single-GPU all-vs-all of the same genomes
computed on every rank independently (bit-exact, every pair exactly once)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genome.distance_b200 as gkd
from genome.distance_b200 import sharding


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, length, k = 23, 400_000, 21
    seqs = []
    for g in range(n):
        t = torch.empty(length if g % 5 else length // 7, dtype=torch.uint8, device=dev)
        gkd.synth(t, 17, g % 3, g // 3, 0.02 if g // 3 else 0.0, device=local)
        seqs.append(t)
    # the single-GPU answer comes from the bucket-merge kernel, the ring runs the block join on every block
    # (GKD_ISECT_ALGO is read when a context is created): the two kernel-4 forms check each other
    os.environ["GKD_ISECT_ALGO"] = "merge"
    with gkd.Engine(k=k, device=local) as ref:
        for s in seqs:
            ref.add(s)
        ref.build()
        want_i, want_d = ref.all_vs_all()
    total = n * (n - 1) // 2
    mine = sharding.genome_slice(n, world, rank)
    ok = True
    # panel ring: no rank ever holds more than its slice plus two panels; tiny workspace -> several arenas
    os.environ["GKD_ISECT_ALGO"] = os.environ.get("GKD_CHECK_RING_ALGO", "join")
    with gkd.Engine(k=k, device=local, workspace_bytes=24 << 20) as eng:
        for g in mine:
            eng.add(seqs[g])
        eng.build()
        stats = {}
        si, sj, s_inter, s_dist = sharding.ring_all_vs_all(eng, n, world, rank, dev, panel_genomes=3, stats=stats)
        ok = ok and len(eng) == len(mine)
    lin = np.array([sharding.row_start(int(a), n) + int(b) - int(a) - 1 for a, b in zip(si, sj)], dtype=np.int64)
    ok = ok and np.array_equal(s_inter, want_i[lin]) and np.array_equal(s_dist, want_d[lin])
    n_stream = torch.tensor([len(lin)], device=dev)
    dist.all_reduce(n_stream)
    ok = ok and int(n_stream) == total
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-gpu parity", "OK" if int(flag) else "MISMATCH", "world", world, "pairs", total)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
