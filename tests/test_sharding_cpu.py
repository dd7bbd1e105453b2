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
