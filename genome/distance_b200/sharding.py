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


# ------------------------------------------------------------------------------------------------
# Streamed column panels: all-vs-all when the sets do not fit one GPU (BASELINE config 4)
# ------------------------------------------------------------------------------------------------
def ring_partners(world: int, rank: int):
    """Steps s = 1..world//2 of the block ring: rank x owns block (x, (x+s) % world) and therefore
    receives the panel of src = (x+s) % world and sends its own to dst = (x-s) % world.  For even
    worlds the half-way step pairs the same two ranks in both directions; that block is split between
    them (see half_step_ranges) so every rank does the same amount of work.
    Returns [(src, dst, half)]."""
    return [((rank + s) % world, (rank - s) % world, world % 2 == 0 and s == world // 2)
            for s in range(1, world // 2 + 1)]


def half_step_ranges(m_mine: int, m_partner: int, i_am_lower: bool):
    """Split of the half-way block X x Y between its two ranks (X = lower rank's genomes, Y = higher's):
    the lower rank computes X[:h] x Y, the higher rank computes Y x X[h:], h = ceil(|X| / 2).
    Returns (rows, recv, send) as (start, stop) ranges of local set indices: my rows to use, the
    partner's sets I receive, my sets I send."""
    if i_am_lower:
        h = (m_mine + 1) // 2
        return (0, h), (0, m_partner), (h, m_mine)
    h = (m_partner + 1) // 2
    return (0, m_mine), (h, m_partner), (0, m_mine)


def streamed_all_vs_all(eng, n_genomes: int, world: int, rank: int, device, panel_genomes: int = 256):
    """All-vs-all without ever holding more than (own slice + one sub-panel) of sets per rank.

    Every rank keeps its genome slice resident (built, ids 0..m-1), computes its diagonal block,
    then walks the ring: the owner of a column panel sends it in sub-panels of `panel_genomes` sets
    (NCCL send/recv over NVLink), the receiver adopts a sub-panel (`import_sets`), intersects
    its rows against it (`query_vs_ref`), and drops it (`truncate`).  Every unordered pair of
    genomes is computed exactly once somewhere; there is no reduction.

    Returns this rank's results as (gi, gj, inter, dist) arrays with global ids gi < gj.
    """
    import torch
    import torch.distributed as dist

    mine = genome_slice(n_genomes, world, rank)
    m = len(mine)
    per = (n_genomes + world - 1) // world
    sizes = torch.zeros(per, dtype=torch.int64, device=device)
    for i in range(m):
        sizes[i] = eng.set_tensor(i).numel()
    all_sizes = [torch.zeros(per, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [t.cpu().numpy() for t in all_sizes]

    out_i, out_j, out_inter, out_dist = [], [], [], []

    def emit(rows, cols, inter, d):
        rows, cols = np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)
        out_i.append(np.minimum(rows, cols))
        out_j.append(np.maximum(rows, cols))
        out_inter.append(np.asarray(inter).reshape(-1))
        out_dist.append(np.asarray(d).reshape(-1))

    # diagonal block: my genomes against each other
    if m >= 2:
        cnt = m * (m - 1) // 2
        inter, d = eng.all_vs_all_range(m, 0, cnt)
        a, b = pair_lists(m, 0, cnt)
        base = mine[0]
        emit(a.astype(np.int64) + base, b.astype(np.int64) + base, inter, d)

    for src, dst, half in ring_partners(world, rank):
        theirs = genome_slice(n_genomes, world, src)
        if half:  # src == dst: the block is split between the two ranks
            rows, recv_rng, send_rng = half_step_ranges(m, len(theirs), rank < src)
        else:
            rows, recv_rng, send_rng = (0, m), (0, len(theirs)), (0, m)
        my_rows = np.arange(rows[0], rows[1], dtype=np.uint32)
        n_recv = (recv_rng[1] - recv_rng[0] + panel_genomes - 1) // panel_genomes
        n_send = (send_rng[1] - send_rng[0] + panel_genomes - 1) // panel_genomes
        for k in range(max(n_recv, n_send)):
            ops, send, recv, offs, chunk = [], None, None, None, None
            if k < n_send:
                lo = send_rng[0] + k * panel_genomes
                hi = min(send_rng[1], lo + panel_genomes)
                parts = [eng.set_tensor(i) for i in range(lo, hi)]
                send = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64, device=device)
                if send.numel() == 0:
                    send = torch.zeros(1, dtype=torch.int64, device=device)
                ops.append(dist.P2POp(dist.isend, send, dst))
            if k < n_recv:
                li0 = recv_rng[0] + k * panel_genomes
                li1 = min(recv_rng[1], li0 + panel_genomes)
                chunk = theirs[li0:li1]
                sz = all_sizes[src][li0:li1].astype(np.uint64)
                offs = np.zeros(len(chunk) + 1, dtype=np.uint64)
                offs[1:] = np.cumsum(sz)
                recv = torch.empty(max(int(offs[-1]), 1), dtype=torch.int64, device=device)
                ops.append(dist.P2POp(dist.irecv, recv, src))
            # one NCCL group per sub-panel: every rank's send and receive are posted together, so the
            # ring cannot deadlock on unmatched point-to-point calls
            for r in (dist.batch_isend_irecv(ops) if ops else []):
                r.wait()
            if recv is not None and len(my_rows) > 0:
                first = eng.import_sets(recv[: int(offs[-1])], offs)
                cols = np.arange(first, first + len(chunk), dtype=np.uint32)
                inter, d = eng.query_vs_ref(my_rows, cols)
                rows_g = np.repeat(np.asarray(mine[rows[0]:rows[1]], dtype=np.int64), len(chunk))
                cols_g = np.tile(np.asarray(chunk, dtype=np.int64), len(my_rows))
                emit(rows_g, cols_g, inter, d)
                eng.truncate(m)
            del send, recv
    if not out_i:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.float64)
    return np.concatenate(out_i), np.concatenate(out_j), np.concatenate(out_inter), np.concatenate(out_dist)
