#!/usr/bin/env python
"""BASELINE config 4 (N synthetic 5 Mbp genomes all-vs-all, tile-sharded over the GPUs of one box)
through the overlapped panel ring (sharding.ring_all_vs_all): no rank ever holds more than its own
slice of the sets plus two panels, so the total set volume may exceed one GPU's HBM.

  torchrun --nproc-per-node 8 tools/run_c4.py --genomes 20000 --panel 250

Prints one JSON line on rank 0: pairs/s aggregate and per GPU, kernel-4 roofline fraction of the slowest
rank, exposed exchange time, and the parity evidence:
  * >= --check pairs sampled uniformly from every rank's results are recomputed from the generator
    seeds in a fresh single-GPU context (two genomes, gkd_pair) and must match bit for bit;
  * --oracle of those pairs are also recomputed by the CPU oracle (integer mode) on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genome.distance_b200 as gkd
from genome.distance_b200 import sharding

SEED = 0x5EED0000
RATES = [0.001, 0.01, 0.05, 0.2]


def params(g, per):
    fam, mem = g // per, g % per
    return fam, mem, (0.0 if mem == 0 else RATES[mem % 4])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=20000)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--families", type=int, default=100)
    ap.add_argument("--panel", type=int, default=250)
    ap.add_argument("--check", type=int, default=1024, help="pairs recomputed in a fresh context (all ranks together)")
    ap.add_argument("--oracle", type=int, default=2, help="pairs also recomputed by the CPU oracle on rank 0")
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = a.genomes
    per = max(1, (n + a.families - 1) // a.families)
    mine = sharding.genome_slice(n, world, rank)
    # one build batch (= one set arena = one exchange panel) per `panel` genomes: the text buffers of a batch
    # are read asynchronously, so there is one per genome of the batch
    nbuf = max(2, min(a.panel, len(mine)))
    buf = torch.empty((nbuf, a.length), dtype=torch.uint8, device=dev)
    eng = gkd.Engine(k=21, device=local, workspace_bytes=nbuf * a.length * 16 + (1 << 24))
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for i, g in enumerate(mine):
        if i % nbuf == 0 and i:
            eng.build()  # consumes the text buffers before they are overwritten
        fam, mem, rate = params(g, per)
        gkd.synth(buf[i % nbuf], SEED, fam, mem, rate, device=local)
        eng.add(buf[i % nbuf])
    eng.build()
    torch.cuda.synchronize(); dist.barrier()
    t1 = time.perf_counter()

    # results: checksums over everything, and a uniform sample of this rank's pairs for the parity check
    per_rank_check = (a.check + world - 1) // world
    rng = np.random.default_rng(1234 + rank)
    agg = {"pairs": 0, "inter_sum": 0, "dist_sum": 0.0, "related": 0}
    sample = []  # (gi, gj, inter, dist)
    my_pairs_expected = n * (n - 1) // 2 / world

    def sink(gi, gj, inter, d):
        inter, d = np.asarray(inter).reshape(-1), np.asarray(d).reshape(-1)
        agg["pairs"] += inter.size
        agg["inter_sum"] += int(inter.sum(dtype=np.uint64))
        agg["dist_sum"] += float(d.sum())
        agg["related"] += int((d < 1.0).sum())
        # binomial thinning keeps ~2x per_rank_check pairs overall, uniformly over this rank's pairs
        k = rng.binomial(inter.size, min(1.0, 2.0 * per_rank_check / my_pairs_expected))
        if k:
            idx = rng.choice(inter.size, size=k, replace=False)
            gi, gj = np.asarray(gi).reshape(-1), np.asarray(gj).reshape(-1)
            for t in idx:
                sample.append((int(min(gi[t], gj[t])), int(max(gi[t], gj[t])), int(inter[t]), float(d[t])))

    stats = {}
    sharding.ring_all_vs_all(eng, n, world, rank, dev, panel_genomes=a.panel, sink=sink, stats=stats)
    rng.shuffle(sample)
    torch.cuda.synchronize()
    t_mine = time.perf_counter() - t1
    m = eng.metrics()
    dist.barrier()
    t2 = time.perf_counter()
    peak_mem = torch.cuda.max_memory_allocated(dev)
    free_b, total_b = torch.cuda.mem_get_info(dev)
    eng.close()

    # parity: recompute sampled pairs from the seeds in a fresh context (different build batch, two-set
    # all-vs-all instead of a rectangular block) and compare bit for bit
    chosen = sample[:per_rank_check]  # already a uniform random sample of this rank's pairs
    bad = 0
    t3 = time.perf_counter()
    with gkd.Engine(k=21, device=local) as chk:
        ta, tb = buf[0], buf[1]
        for gi, gj, inter, d in chosen:
            chk.reset()
            fa, ma, ra = params(gi, per)
            fb, mb, rb = params(gj, per)
            gkd.synth(ta, SEED, fa, ma, ra, device=local)
            gkd.synth(tb, SEED, fb, mb, rb, device=local)
            x, y = chk.add(ta), chk.add(tb)
            chk.build()
            i2, _, d2 = chk.pair(x, y)
            bad += int(i2 != inter or d2 != d)
    oracle_ok = None
    if rank == 0 and a.oracle:
        from oracle import oracle as orc
        from oracle import synth as osynth

        oracle_ok = True
        for gi, gj, inter, d in chosen[: a.oracle]:
            sa = orc.IntSet(osynth.synth(a.length, SEED, *params(gi, per)).tobytes(), 21)
            sb = orc.IntSet(osynth.synth(a.length, SEED, *params(gj, per)).tobytes(), 21)
            oracle_ok = oracle_ok and sa.similarity(sb) == inter and sa.distance(sb) == d
    t4 = time.perf_counter()

    isect_gbs = m["total_intersect_bytes"] / max(m["total_intersect_ms"] * 1e-3, 1e-9) / 1e9
    vec = torch.tensor([float(agg["pairs"]), float(len(chosen)), float(bad), float(agg["related"]), agg["dist_sum"],
                        float(agg["inter_sum"] % (1 << 52))], dtype=torch.float64, device=dev)
    vmax = torch.tensor([t_mine, m["total_intersect_ms"], stats.get("exposed_wait_s", 0.0), float(peak_mem),
                         -isect_gbs, float(total_b - free_b)], dtype=torch.float64, device=dev)
    dist.all_reduce(vec)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    total_pairs = n * (n - 1) // 2
    if rank == 0:
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
        slow_gbs = -float(vmax[4])
        print(json.dumps({
            "config": f"c4: {n} x {a.length / 1e6:g} Mbp all-vs-all DNA K=21, overlapped panel ring, panels of {a.panel} sets",
            "n_gpus": world, "pairs": total_pairs, "pairs_computed": int(vec[0]),
            "stored_set_bytes_total": int(m["keys_unique"] / max(len(mine), 1) * n * 4.25),
            "build_s": t1 - t0, "distance_s": t2 - t1, "slowest_rank_distance_s": float(vmax[0]),
            "pairs_per_s": total_pairs / (t2 - t1), "pairs_per_s_per_gpu": total_pairs / (t2 - t1) / world,
            "kernel4_ms_slowest_rank": float(vmax[1]), "kernel4_algorithmic_GBps_slowest_rank": slow_gbs,
            "kernel4_frac_of_measured_hbm_slowest_rank": slow_gbs / peak,
            "exposed_exchange_wait_s_max": float(vmax[2]),
            "torch_peak_allocated_bytes_max": int(vmax[3]), "device_bytes_in_use_max": int(vmax[5]),
            "related_pairs": int(vec[3]), "dist_sum": float(vec[4]),
            "parity": {"pairs_rechecked_in_fresh_context": int(vec[1]), "mismatches": int(vec[2]),
                       "oracle_pairs_rank0": a.oracle, "oracle_ok": oracle_ok, "recheck_s": t4 - t3}}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(vec[2]) == 0 and int(vec[0]) == total_pairs and oracle_ok is not False else 1)


if __name__ == "__main__":
    main()
