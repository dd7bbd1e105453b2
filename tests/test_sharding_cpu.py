"""Host-side multi-GPU logic on CPU: partitions, pair enumeration, panel plans, and the overlapped panel
ring over gloo at world sizes 2-4 through the same functions the GPU bench uses (the compute stand-in is
the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genome.distance_b200 import sharding


def test_slices_partition_everything():
    for n in (0, 1, 2, 7, 10, 1000, 1001):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s = sharding.genome_slice(n, world, r)
                assert all(sharding.owner_of(g, n, world) == r for g in s)
                seen += s
            assert seen == list(range(n))
            total = n * (n - 1) // 2
            cover = 0
            for r in range(world):
                first, count = sharding.pair_slice(total, world, r)
                assert first == cover
                cover += count
            assert cover == total


def test_pair_lists_match_row_major_order():
    for n in (2, 3, 5, 12, 33):
        full = [(i, j) for i in range(n) for j in range(i + 1, n)]
        total = len(full)
        a, b = sharding.pair_lists(n, 0, total)
        assert list(zip(a.tolist(), b.tolist())) == full
        for first, count in ((0, 0), (min(1, total - 1), 1), (total // 3, total // 2), (total - 1, 1), (total, 0)):
            a, b = sharding.pair_lists(n, first, count)
            assert list(zip(a.tolist(), b.tolist())) == full[first:first + count]
        for i in range(n - 1):
            assert full[sharding.row_start(i, n)] == (i, i + 1)
        with pytest.raises(ValueError):
            sharding.pair_lists(n, total, 1)


def _seqs(n, length):
    import genome.distance_b200 as gkd

    out = []
    for g in range(n):
        a = np.empty(length, dtype=np.uint8)
        gkd.synth(a, 99, g % 2, g // 2, 0.03 if g // 2 else 0.0)
        out.append(a.tobytes())
    return out


def test_ring_partners_cover_every_rank_pair_once():
    for world in (1, 2, 3, 4, 5, 8):
        blocks = {}
        for r in range(world):
            for src, dst, half in sharding.ring_partners(world, r):
                assert (not half) or src == dst
                # my receive is matched by the partner's send in the same step
                assert any(p[1] == r for p in sharding.ring_partners(world, src))
                blocks.setdefault(frozenset((r, src)), []).append((r, half))
        assert len(blocks) == world * (world - 1) // 2
        for key, owners in blocks.items():
            # a normal block has one owner; the half-way block of an even ring is shared by its two ranks
            assert len(owners) == (2 if owners[0][1] else 1), (world, key, owners)
    for m_lo, m_hi in ((5, 4), (4, 4), (1, 1), (0, 3)):
        rows_l, recv_l, send_l = sharding.half_step_ranges(m_lo, m_hi, True)
        rows_h, recv_h, send_h = sharding.half_step_ranges(m_hi, m_lo, False)
        assert recv_l == send_h and recv_h == send_l
        # lower covers X[:h] x Y, higher covers Y x X[h:]: together the whole block, no overlap
        assert (rows_l[1] - rows_l[0]) + (recv_h[1] - recv_h[0]) == m_lo and rows_h == (0, m_hi)


class _CpuPanelEngine:
    """CPU stand-in for gkd.Engine with the calls ring_all_vs_all needs.  Sets are the oracle's key arrays;
    an "arena" is a uint8 tensor that holds the keys of `per_arena` consecutive sets back to back, each
    block padded to 16 bytes, described with the engine's packed-set table (numpy stand-ins for kernels 4-5)."""

    def __init__(self, orc, seqs, k, per_arena=2):
        from genome.distance_b200.engine import PACKED_DTYPE

        self.orc = orc
        self.sets = [orc.IntSet(s, k).keys().astype(np.int64) for s in seqs]
        self.own = []  # (first, n, uint8 tensor, table)
        for first in range(0, len(self.sets), per_arena):
            chunk = self.sets[first:first + per_arena]
            table = np.zeros(len(chunk), dtype=PACKED_DTYPE)
            blobs, off = [], 0
            for i, keys in enumerate(chunk):
                raw = keys.tobytes() + b"\0" * ((-keys.nbytes) % 16 + 16)
                table[i]["offs_off"] = table[i]["lows_off"] = off
                table[i]["n"] = keys.size
                blobs.append(raw)
                off += len(raw)
            self.own.append((first, len(chunk), torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8), table))
        self.adopted_buffers = 0

    def arena_meta(self):
        return [(first, n, t.numel(), table) for first, n, t, table in self.own]

    def arena_view(self, arena, begin, end):
        return self.own[arena][2][begin:max(end, begin + 1)]

    def adopt_sets(self, buf, table):
        first = len(self.sets)
        raw = buf.numpy().tobytes()
        for row in table:
            lo = int(row["lows_off"])
            self.sets.append(np.frombuffer(raw[lo:lo + 8 * int(row["n"])], dtype=np.int64))
        self.adopted_buffers += 1
        return first

    def _pair(self, i, j):
        x, y = self.sets[i], self.sets[j]
        inter = 2 * np.intersect1d(x, y, assume_unique=True).size  # odd K: both-strand count
        return inter, self.orc.distance(inter, 2 * x.size, 2 * y.size)

    def truncate(self, n):
        del self.sets[n:]

    def all_vs_all_range(self, n, first, count):
        a, b = sharding.pair_lists(n, first, count)
        r = [self._pair(int(i), int(j)) for i, j in zip(a, b)]
        return np.array([x[0] for x in r], dtype=np.uint64), np.array([x[1] for x in r], dtype=np.float64)

    def query_vs_ref(self, q, r):
        res = [[self._pair(int(i), int(j)) for j in r] for i in q]
        return (np.array([[x[0] for x in row] for row in res], dtype=np.uint64).reshape(len(q), len(r)),
                np.array([[x[1] for x in row] for row in res], dtype=np.float64).reshape(len(q), len(r)))


def test_plan_panels_is_a_partition():
    from genome.distance_b200.engine import PACKED_DTYPE

    meta = []
    first = 0
    for n in (3, 1, 4):
        table = np.zeros(n, dtype=PACKED_DTYPE)
        table["offs_off"] = np.arange(n) * 256
        table["lows_off"] = table["offs_off"] + 32
        meta.append((first, n, n * 256, table))
        first += n
    for lo, hi, mx in ((0, 8, 2), (1, 7, 3), (2, 3, 1), (5, 5, 4), (0, 8, 100)):
        plan = sharding.plan_panels(meta, lo, hi, mx)
        ids = [i for p in plan for i in range(p["first"], p["first"] + p["count"])]
        assert ids == list(range(lo, hi))
        for p in plan:
            assert p["count"] <= mx and p["table"]["offs_off"][0] == 0 and p["end"] > p["begin"]
            assert (p["table"]["lows_off"] - p["table"]["offs_off"] == 32).all()


def _ring_worker(rank, world, port, n, length, k, panel, group, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    seqs = _seqs(n, length)
    mine = sharding.genome_slice(n, world, rank)
    eng = _CpuPanelEngine(orc, [seqs[g] for g in mine], k)
    stats = {}
    gi, gj, inter, d = sharding.ring_all_vs_all(eng, n, world, rank, torch.device("cpu"), panel_genomes=panel, stats=stats,
                                                 group_sets=group)
    assert stats["groups"] <= stats["slots"]
    assert len(eng.sets) == len(mine)  # every received panel was dropped again
    ret.put((rank, gi.tolist(), gj.tolist(), [int(x) for x in inter], d.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,panel,group", [(2, 2, 1024), (3, 1, 2), (4, 3, 1), (4, 1, 1024)])
def test_ring_all_vs_all_over_gloo(orc, world, panel, group):
    """world_size 2-4 on CPU: panels planned from the owners' arena layouts, moved with send/recv, adopted,
    intersected in groups of up to `group` sets (several panels, possibly of several peers, per call) and
    dropped; every pair computed exactly once and equal to the oracle"""
    n, length, k = 9, 12000, 15
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, n, length, k, panel, group, ret), daemon=True)
             for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, gi, gj, inter, d = ret.get(timeout=180)
        for a, b, i, x in zip(gi, gj, inter, d):
            assert (a, b) not in got, "pair computed twice"
            got[(a, b)] = (i, x)
    for p in procs:
        p.join(timeout=60)
    sets = [orc.IntSet(x, k) for x in _seqs(n, length)]
    want = {(i, j): (sets[i].similarity(sets[j]), sets[i].distance(sets[j])) for i in range(n) for j in range(i + 1, n)}
    assert got == want
