#!/usr/bin/env python
"""torchrun check of the N>1 path on real GPUs: every rank builds its genome slice, sets are exchanged
over NCCL, each rank computes its pair slice; the gathered result must equal a single-GPU
all-vs-all of the same genomes computed on every rank independently (bit-exact)."""
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
    with gkd.Engine(k=k, device=local) as ref:
        for s in seqs:
            ref.add(s)
        ref.build()
        want_i, want_d = ref.all_vs_all()
    total = n * (n - 1) // 2
    mine = sharding.genome_slice(n, world, rank)
    first, count = sharding.pair_slice(total, world, rank)
    with gkd.Engine(k=k, device=local) as eng:
        for g in mine:
            eng.add(seqs[g])
        eng.build()
        id_map = sharding.exchange_sets(eng, n, world, rank, dev)
        ia, ib = sharding.local_pair_ids(id_map, n, first, count)
        gi, gd = eng.pairs(ia, ib)
    ok = np.array_equal(gi, want_i[first:first + count]) and np.array_equal(gd, want_d[first:first + count])
    # streamed column panels (config-4 path): same numbers without any rank holding every set
    with gkd.Engine(k=k, device=local) as eng:
        for g in mine:
            eng.add(seqs[g])
        eng.build()
        si, sj, s_inter, s_dist = sharding.streamed_all_vs_all(eng, n, world, rank, dev, panel_genomes=3)
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
