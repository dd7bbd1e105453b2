"""Multi-GPU sharding of the all-vs-all pair matrix (one process per GPU, torch.distributed).

The path shards naturally: set construction is per genome, and the pair matrix has no
cross-pair dependence and no reduction.  Rank r builds the sets of a contiguous slice of the genomes;
ONE exchange step makes every set available on every rank (one broadcast per set from its owner,
NCCL over NVLink on GPUs, gloo in the CPU tests); then rank r intersects a contiguous slice of the
row-major strict-upper-triangle pair enumeration.  Results return to the host per rank.

Everything here is host logic; it is exercised on CPU with gloo (tests/test_sharding_cpu.py) through
the same functions the GPU bench uses.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def genome_slice(n_genomes: int, world: int, rank: int) -> List[int]:
    """Contiguous block distribution of genome indices (sizes differ by at most one)."""
    base, extra = divmod(n_genomes, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def owner_of(g: int, n_genomes: int, world: int) -> int:
    base, extra = divmod(n_genomes, world)
    cut = extra * (base + 1)
    return g // (base + 1) if g < cut else extra + (g - cut) // max(base, 1)


def pair_slice(total_pairs: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of this rank's contiguous slice of the pair enumeration."""
    base, extra = divmod(total_pairs, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def row_start(i: int, n: int) -> int:
    """linear index of pair (i, i+1) in the row-major strict upper triangle of an n x n matrix"""
    return i * (2 * n - i - 1) // 2


def pair_lists(n: int, first: int, count: int) -> Tuple[np.ndarray, np.ndarray]:
    """Global (i, j) of pairs [first, first+count) -- the same order gkd_all_vs_all_range uses
    (FastaDistanceProcessor.java:177: for jdx in idx+1 .. nSeqs-1)."""
    total = n * (n - 1) // 2 if n > 1 else 0
    if first < 0 or count < 0 or first + count > total:
        raise ValueError(f"pair range [{first}, +{count}) outside 0..{total}")
    a = np.empty(count, dtype=np.uint32)
    b = np.empty(count, dtype=np.uint32)
    if count == 0:
        return a, b
    # find the row of `first`
    lo, hi = 0, n - 2
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if row_start(mid, n) <= first:
            lo = mid
        else:
            hi = mid - 1
    i, j, w = lo, first - row_start(lo, n) + lo + 1, 0
    while w < count:
        take = min(count - w, n - j)
        a[w:w + take] = i
        b[w:w + take] = np.arange(j, j + take, dtype=np.uint32)
        w += take
        i += 1
        j = i + 1
    return a, b


def exchange_sets(eng, n_genomes: int, world: int, rank: int, device) -> Dict[int, int]:
    """Make every genome's key set resident in `eng` on every rank.

    `eng` holds this rank's sets as ids 0..len(slice)-1 (built, in slice order) and must provide
    `set_tensor(id) -> 1-D int64 torch tensor on `device`` and `import_set(tensor) -> id`.
    Returns the map global genome index -> engine set id.
    """
    import torch
    import torch.distributed as dist

    mine = genome_slice(n_genomes, world, rank)
    id_map = {g: i for i, g in enumerate(mine)}
    # set sizes of every genome (one all_gather of a padded vector)
    per = (n_genomes + world - 1) // world
    sizes = torch.zeros(per, dtype=torch.int64, device=device)
    for i in range(len(mine)):
        sizes[i] = eng.set_tensor(i).numel()
    all_sizes = [torch.zeros(per, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [t.cpu().tolist() for t in all_sizes]
    if hasattr(eng, "import_sets") and hasattr(dist, "all_gather_into_tensor"):
        # ONE collective: every rank packs its sets back to back into a send buffer padded to the
        # largest per-rank total, all_gather_into_tensor moves them at full NVLink bandwidth, and each
        # peer's segment is adopted with one batched import.
        totals = [sum(int(x) for x in all_sizes[o][: len(genome_slice(n_genomes, world, o))]) for o in range(world)]
        cap = max(max(totals), 1)
        send = torch.empty(cap, dtype=torch.int64, device=device)
        off = 0
        for i in range(len(mine)):
            t = eng.set_tensor(i)
            send[off:off + t.numel()].copy_(t)
            off += t.numel()
        gathered = torch.empty(world * cap, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(gathered, send)
        del send
        for owner in range(world):
            if owner == rank:
                continue
            theirs = genome_slice(n_genomes, world, owner)
            offs = np.zeros(len(theirs) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum(np.asarray([int(all_sizes[owner][li]) for li in range(len(theirs))], dtype=np.uint64))
            first = eng.import_sets(gathered[owner * cap: owner * cap + int(offs[-1])], offs)
            for li, g in enumerate(theirs):
                id_map[g] = first + li
        del gathered
        return id_map
    # generic path (any engine with import_set; used by the CPU/gloo tests): one broadcast per set
    for owner in range(world):
        theirs = genome_slice(n_genomes, world, owner)
        sizes_o = [int(all_sizes[owner][li]) for li in range(len(theirs))]
        if owner == rank:
            for li in range(len(theirs)):
                if sizes_o[li]:
                    dist.broadcast(eng.set_tensor(li), src=owner)
            continue
        recv = torch.empty(max(max(sizes_o) if sizes_o else 0, 1), dtype=torch.int64, device=device)
        for li, g in enumerate(theirs):
            buf = recv[: sizes_o[li]]
            if sizes_o[li]:
                dist.broadcast(buf, src=owner)
            id_map[g] = eng.import_set(buf)
        del recv
    return id_map


def local_pair_ids(id_map: Dict[int, int], n_genomes: int, first: int, count: int) -> Tuple[np.ndarray, np.ndarray]:
    """engine set ids of this rank's pair slice"""
    a, b = pair_lists(n_genomes, first, count)
    lut = np.empty(n_genomes, dtype=np.uint32)
    for g, i in id_map.items():
        lut[g] = i
    return lut[a], lut[b]
