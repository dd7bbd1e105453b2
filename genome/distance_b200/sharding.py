"""Multi-GPU sharding of the all-vs-all pair matrix (one process per GPU, torch.distributed).

The path shards naturally: set construction is per genome, and the pair matrix has no
cross-pair dependence and no reduction.  Rank r builds the sets of a contiguous slice of the genomes and
owns a set of rank blocks of the pair matrix; the sets of the peers it needs arrive as panels over NCCL
send/recv (NVLink on GPUs, gloo in the CPU tests) while earlier blocks are being intersected, and are
adopted in place (ring_all_vs_all below).  Results return to the host per rank.

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


# ------------------------------------------------------------------------------------------------
# Panel ring: all-vs-all across ranks with the set exchange hidden behind the intersect kernel
# ------------------------------------------------------------------------------------------------
# Every rank keeps only its genome slice resident.  The pair matrix is cut into rank blocks; rank x owns
# its diagonal block and the blocks (x, (x+s) mod R) for s = 1..R/2 (the half-way block of an even ring is
# split between its two ranks), so it needs the sets of at most R/2 peers, one peer at a time.  A peer's
# slice arrives as PANELS: byte ranges of the owner's set arenas (bucket tables + key low words exactly as
# kernel 4 reads them), moved with NCCL send/recv over NVLink and adopted in place by the receiver
# (gkd_adopt_sets: no unpack, no second copy).  The transfer of panel k+1 is posted before the kernels of
# panel k are launched, so the only exposed transfer is the first one -- and that one runs under the
# rank's diagonal block, which needs no remote data.  Every unordered pair of genomes is computed exactly
# once; there is no reduction.  The same code serves config 2 (sets fit one GPU) and config 4 (they do not).
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


def plan_panels(meta, lo: int, hi: int, max_sets: int):
    """Panels that carry the local sets lo..hi-1 of a rank whose arenas are described by `meta`
    (Engine.arena_meta(): [(first_id, n_sets, bytes, table)], table offsets relative to the arena base).
    A panel never spans two arenas and holds at most max_sets sets.  Pure function of its arguments, so
    the sender and the receiver derive the same plan.  Returns dicts: arena (index), first (local id),
    count, begin/end (byte range inside the arena) and table (offsets relative to `begin`)."""
    out = []
    for ai, (first, n, nbytes, table) in enumerate(meta):
        a, b = max(lo, first), min(hi, first + n)
        while a < b:
            m = min(b - a, max(1, max_sets))
            t = np.array(table[a - first:a - first + m], copy=True)
            begin = int(t["offs_off"][0])
            end = int(nbytes) if a + m == first + n else int(table["offs_off"][a - first + m])
            for f in ("offs_off", "lows_off"):
                t[f] -= np.uint64(begin)
            for f in ("pal_offs_off", "pal_lows_off"):
                nz = t[f] != 0
                t[f][nz] -= np.uint64(begin)
            out.append({"arena": ai, "first": a, "count": m, "begin": begin, "end": end, "table": t})
            a += m
    return out


def ring_all_vs_all(eng, n_genomes: int, world: int, rank: int, device, panel_genomes: int = 128, sink=None,
                    stats=None, group_sets: int = 1024):
    """All-vs-all across `world` ranks (see the section comment).  `eng` holds this rank's genome slice
    as built sets 0..m-1 and provides arena_meta / arena_view / adopt_sets / all_vs_all_range /
    query_vs_ref / truncate.  `sink(gi, gj, inter, dist)` receives every block (global ids, row-major);
    by default the blocks are collected and returned as (gi, gj, inter, dist) with gi < gj.

    Panels are the unit of TRANSFER; the unit of COMPUTE is a group of consecutive panels that meet the same
    rows, up to `group_sets` sets (they may come from several peers): kernel 4's block join builds one table
    per (64 rows, key range) and probes every column of the call with it, so a call with many columns
    amortises the tables.  The transfers of group g+1 are in flight while group g is computed, so at most
    two groups of received sets are resident beside the rank's own slice."""
    import time

    import torch
    import torch.distributed as dist

    on_gpu = torch.device(device).type == "cuda"
    mine = genome_slice(n_genomes, world, rank)
    m = len(mine)
    out_i, out_j, out_inter, out_dist = [], [], [], []

    def collect(rows, cols, inter, d):
        rows, cols = np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)
        out_i.append(np.minimum(rows, cols))
        out_j.append(np.maximum(rows, cols))
        out_inter.append(np.asarray(inter).reshape(-1))
        out_dist.append(np.asarray(d).reshape(-1))

    emit = sink or collect
    t_wait = 0.0

    # everybody learns everybody's arena layout (a few KB per rank), so panel plans need no handshake
    metas = [None] * world
    if world > 1:
        dist.all_gather_object(metas, eng.arena_meta())
    else:
        metas[0] = eng.arena_meta()

    # flat list of transfer slots: (send panel | None, dst, recv panel | None, src, my rows, their slice)
    slots = []
    for src, dst, half in ring_partners(world, rank):
        theirs = genome_slice(n_genomes, world, src)
        if half:  # src == dst: the block is split between the two ranks
            rows, recv_rng, send_rng = half_step_ranges(m, len(theirs), rank < src)
        else:
            rows, recv_rng, send_rng = (0, m), (0, len(theirs)), (0, m)
        recv_list = plan_panels(metas[src], recv_rng[0], recv_rng[1], panel_genomes)
        send_list = plan_panels(metas[rank], send_rng[0], send_rng[1], panel_genomes)
        for k in range(max(len(recv_list), len(send_list))):
            slots.append((send_list[k] if k < len(send_list) else None, dst,
                          recv_list[k] if k < len(recv_list) else None, src, rows, theirs))

    def post(slot):
        """start the transfer of one slot (asynchronous on GPUs); one group per slot, so every rank's send
        and receive are posted together and the ring cannot deadlock on unmatched point-to-point calls"""
        send_p, dst, recv_p, src, rows, theirs = slot
        ops, keep, buf = [], [], None
        if send_p is not None:
            view = eng.arena_view(send_p["arena"], send_p["begin"], send_p["end"])
            keep.append(view)
            ops.append(dist.P2POp(dist.isend, view, dst))
        if recv_p is not None:
            buf = torch.empty(max(recv_p["end"] - recv_p["begin"], 16), dtype=torch.uint8, device=device)
            ops.append(dist.P2POp(dist.irecv, buf, src))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        return reqs, buf, keep

    def finish(reqs):
        nonlocal t_wait
        t0 = time.perf_counter()
        for r in reqs:
            r.wait()
        if on_gpu:
            torch.cuda.current_stream().synchronize()  # only what this stream waits for: the slot's transfer
        t_wait += time.perf_counter() - t0

    # compute groups: consecutive slots with the same rows, at most group_sets received sets together
    groups, cur_g, cur_cols = [], [], 0
    for idx, slot in enumerate(slots):
        n_recv = slot[2]["count"] if slot[2] is not None else 0
        if cur_g and (slots[cur_g[0]][4] != slot[4] or cur_cols + n_recv > max(1, group_sets)):
            groups.append(cur_g)
            cur_g, cur_cols = [], 0
        cur_g.append(idx)
        cur_cols += n_recv
    if cur_g:
        groups.append(cur_g)

    posted = {}

    def post_group(g):
        for idx in g:
            posted[idx] = post(slots[idx])

    if groups:
        post_group(groups[0])

    # diagonal block: my genomes against each other (runs while the first group is in flight)
    if m >= 2:
        cnt = m * (m - 1) // 2
        inter, d = eng.all_vs_all_range(m, 0, cnt)
        a, b = pair_lists(m, 0, cnt)
        base = mine[0]
        emit(a.astype(np.int64) + base, b.astype(np.int64) + base, inter, d)

    for gi_, g in enumerate(groups):
        if gi_ + 1 < len(groups):
            post_group(groups[gi_ + 1])  # in flight while this group is intersected
        rows = slots[g[0]][4]
        parts, bufs, cols = [], [], []
        for idx in g:
            reqs, buf, keep = posted.pop(idx)
            finish(reqs)
            send_p, dst, recv_p, src, _, theirs = slots[idx]
            bufs.append((buf, keep))
            if recv_p is not None and rows[1] > rows[0]:
                first = eng.adopt_sets(buf, recv_p["table"])
                cols.append(np.arange(first, first + recv_p["count"], dtype=np.uint32))
                parts.append((recv_p["count"], theirs[recv_p["first"]:recv_p["first"] + recv_p["count"]]))
        if cols:
            my_rows = np.arange(rows[0], rows[1], dtype=np.uint32)
            inter, d = eng.query_vs_ref(my_rows, np.concatenate(cols))
            inter = np.asarray(inter).reshape(len(my_rows), -1)
            d = np.asarray(d).reshape(len(my_rows), -1)
            off = 0
            for count, chunk in parts:
                rows_g = np.repeat(np.asarray(mine[rows[0]:rows[1]], dtype=np.int64), len(chunk))
                cols_g = np.tile(np.asarray(chunk, dtype=np.int64), len(my_rows))
                emit(rows_g, cols_g, np.ascontiguousarray(inter[:, off:off + count]).reshape(-1),
                     np.ascontiguousarray(d[:, off:off + count]).reshape(-1))
                off += count
            eng.truncate(m)
        del bufs
    if stats is not None:
        stats["exposed_wait_s"] = t_wait
        stats["slots"] = len(slots)
        stats["groups"] = len(groups)
    if sink is not None:
        return None
    if not out_i:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.float64)
    return np.concatenate(out_i), np.concatenate(out_j), np.concatenate(out_inter), np.concatenate(out_dist)
