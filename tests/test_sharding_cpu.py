"""Host-side multi-GPU logic on CPU: partitions, pair enumeration, and the world_size-2 set exchange
over gloo through the same functions the GPU bench uses (the compute stand-in is the oracle)."""
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


class _CpuEngine:
    """stand-in with the two methods exchange_sets needs; sets are the oracle's key arrays"""

    def __init__(self, orc, seqs, k):
        self.orc = orc
        self.sets = [torch.from_numpy(orc.IntSet(s, k).keys().astype(np.int64)) for s in seqs]

    def set_tensor(self, i):
        return self.sets[i]

    def import_set(self, t):
        self.sets.append(t.clone())
        return len(self.sets) - 1


def _seqs(n, length):
    import genome.distance_b200 as gkd

    out = []
    for g in range(n):
        a = np.empty(length, dtype=np.uint8)
        gkd.synth(a, 99, g % 2, g // 2, 0.03 if g // 2 else 0.0)
        out.append(a.tobytes())
    return out


def _worker(rank, world, port, n, length, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    seqs = _seqs(n, length)
    mine = sharding.genome_slice(n, world, rank)
    eng = _CpuEngine(orc, [seqs[g] for g in mine], k)
    id_map = sharding.exchange_sets(eng, n, world, rank, torch.device("cpu"))
    assert sorted(id_map) == list(range(n))
    total = n * (n - 1) // 2
    first, count = sharding.pair_slice(total, world, rank)
    ia, ib = sharding.local_pair_ids(id_map, n, first, count)
    out = torch.zeros(total, dtype=torch.int64)
    for t in range(count):
        x, y = eng.sets[ia[t]].numpy(), eng.sets[ib[t]].numpy()
        out[first + t] = np.intersect1d(x, y, assume_unique=True).size
    dist.all_reduce(out)  # disjoint slices: the sum is the concatenation (test-only gather)
    if rank == 0:
        ret.put(out.tolist())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_and_pair_slices_over_gloo(orc, world):
    n, length, k = 7, 20000, 15
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, length, k, ret), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seqs = _seqs(n, length)
    sets = [orc.IntSet(x, k) for x in seqs]
    want = [sets[i].intersect(sets[j])[0] for i in range(n) for j in range(i + 1, n)]
    assert got == want


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


class _CpuPanelEngine(_CpuEngine):
    """adds the batched calls streamed_all_vs_all needs (numpy stand-ins for kernels 4-5)"""

    def _pair(self, i, j):
        x, y = self.sets[i].numpy(), self.sets[j].numpy()
        inter = 2 * np.intersect1d(x, y, assume_unique=True).size  # odd K: both-strand count
        return inter, self.orc.distance(inter, 2 * x.size, 2 * y.size)

    def import_sets(self, t, offs):
        first = len(self.sets)
        for a, b in zip(offs[:-1], offs[1:]):
            self.sets.append(t[int(a):int(b)].clone())
        return first

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


def _stream_worker(rank, world, port, n, length, k, panel, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    seqs = _seqs(n, length)
    mine = sharding.genome_slice(n, world, rank)
    eng = _CpuPanelEngine(orc, [seqs[g] for g in mine], k)
    gi, gj, inter, d = sharding.streamed_all_vs_all(eng, n, world, rank, torch.device("cpu"), panel_genomes=panel)
    assert len(eng.sets) == len(mine)  # every streamed panel was dropped again
    ret.put((rank, gi.tolist(), gj.tolist(), [int(x) for x in inter], d.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,panel", [(2, 2), (3, 1), (4, 3)])
def test_streamed_panels_over_gloo(orc, world, panel):
    n, length, k = 9, 12000, 15
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_stream_worker, args=(r, world, port, n, length, k, panel, ret), daemon=True)
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
