"""Kernel 4, block-join form (csrc/join.cu), against the bucket-merge kernel and the CPU oracle.

GKD_ISECT_ALGO (read when a context is created) pins the kernel-4 variant: "merge" = bucket merge only,
"join" = block join for every call it structurally serves (upper-triangle ranges and query x reference
rectangles; 32-bit low words and 64-bit keys have their own table layouts), default = chosen per call.  Both must give the same exact counts, and the same
doubles, as the oracle's sorted-integer restatement of SequenceKmers.similarity / distance.
"""
import numpy as np
import pytest

import genome.distance_b200 as gkd

pytestmark = pytest.mark.gpu


def _seqs(seed, lens, protein=False, families=3):
    out = []
    for g, n in enumerate(lens):
        a = np.empty(n, dtype=np.uint8)
        mem = g // families
        gkd.synth(a, seed, g % families, mem, [0.0, 0.002, 0.03, 0.15][mem % 4] if mem else 0.0, protein=protein)
        out.append(a.tobytes())
    return out


def _run(monkeypatch, algo, k, alphabet, seqs, calls, jcfg=None, fill=None):
    """build the sets under one kernel-4 variant and run `calls` = list of (kind, args)"""
    monkeypatch.setenv("GKD_ISECT_ALGO", algo)
    if jcfg is not None:
        monkeypatch.setenv("GKD_JOIN_CFG", str(jcfg))
    if fill is not None:
        monkeypatch.setenv("GKD_JOIN_FILL", str(fill))
    res, kernels = [], []
    with gkd.Engine(k=k, alphabet=alphabet) as e:
        for s in seqs:
            e.add(s)
        e.build()
        for kind, args in calls:
            if kind == "all":
                gi, gd = e.all_vs_all()
            elif kind == "range":
                gi, gd = e.all_vs_all_range(len(seqs), *args)
            else:
                gi, gd = e.query_vs_ref(*args)
            res.append((np.asarray(gi).copy(), np.asarray(gd).copy()))
            kernels.append(e.metrics()["intersect_kernel"])
    return res, kernels


LENS = ([300_000] * 20 + [310_000, 150_000, 290_000, 25, 600_000, 40_000, 1_200_000, 0, 300_000, 9_000]
        + [280_000] * 23 + [2_100_000, 20, 75_000])


@pytest.mark.parametrize("k,alphabet", [(21, gkd.DNA), (16, gkd.DNA), (12, gkd.DNA), (5, gkd.PROT), (25, gkd.DNA), (32, gkd.DNA),
                                        (8, gkd.PROT)])
def test_block_join_matches_merge_and_oracle(orc, k, alphabet, monkeypatch):
    """56 sets of mixed sizes (25 bp ... 2.1 Mbp, one empty, related families and unrelated ones): the whole
    triangle, ranges that start and end inside rows, and rectangles with few and many query rows (the kernel
    puts the longer side on the table when there are fewer than 32 queries)."""
    prot = alphabet == gkd.PROT
    lens = [max(n // 4, 8) if n else 0 for n in LENS] if prot else LENS
    seqs = _seqs(77, lens, protein=prot)
    n = len(seqs)
    total = n * (n - 1) // 2
    calls = [("all", ()), ("range", (37, 911)), ("range", (total - 400, 400)), ("range", (0, 60)),
             ("rect", (list(range(0, 40)), list(range(10, n)))),
             ("rect", ([5, 50, 7], list(range(0, n)))),
             ("rect", (list(range(n - 1, -1, -1)), [3, 3, 54, 0]))]
    merge, km = _run(monkeypatch, "merge", k, alphabet, seqs, calls)
    join, kj = _run(monkeypatch, "join", k, alphabet, seqs, calls)
    assert all(x in (3, 4) for x in km)
    assert all(x == 5 for x in kj), kj  # every one of these calls is served by the join when it is pinned
    for (mi, md), (ji, jd), call in zip(merge, join, calls):
        assert np.array_equal(mi, ji), call
        assert np.array_equal(md, jd), call
    if not prot:
        # and the triangle against the oracle (sorted canonical integer sets)
        osets = [orc.IntSet(s, k) for s in seqs]
        gi, gd = join[0]
        t = 0
        for i in range(n):
            for j in range(i + 1, n):
                if (i * 7 + j) % 5 == 0:
                    assert int(gi[t]) == osets[i].similarity(osets[j]), (i, j)
                    assert gd[t] == osets[i].distance(osets[j]), (i, j)
                t += 1


@pytest.mark.parametrize("k,jcfg,fill", [(21, 0, None), (21, 1, None), (21, 2, None), (21, 3, None), (21, 4, None), (21, 5, None),
                                         (21, 6, None), (21, 7, None), (21, 1, 45), (21, 5, 5),
                                         (25, 8, None), (25, 9, None), (25, 10, None), (25, 11, None), (25, 9, 45)])
def test_every_join_configuration_is_exact(k, jcfg, fill, monkeypatch):
    """GKD_JOIN_CFG pins the table geometry (0-7: 32-bit low words, 8-11: 64-bit keys), GKD_JOIN_FILL the target
    load: every geometry gives the merge kernel's counts, for sets whose own bucket tables are finer and coarser
    than the key ranges."""
    seqs = _seqs(5, [200_000] * 34 + [1_500_000, 30_000, 2_000, 800_000], families=2)
    calls = [("all", ()), ("rect", (list(range(2, 38)), list(range(0, 36))))]
    merge, _ = _run(monkeypatch, "merge", k, gkd.DNA, seqs, calls)
    join, kj = _run(monkeypatch, "join", k, gkd.DNA, seqs, calls, jcfg=jcfg, fill=fill)
    assert kj == [5, 5]
    for (mi, md), (ji, jd) in zip(merge, join):
        assert np.array_equal(mi, ji) and np.array_equal(md, jd)


def test_default_choice_uses_the_join_for_matrices_and_the_merge_for_small_calls(monkeypatch):
    """The per-call choice: a 70-genome triangle of 400 kbp genomes goes to the block join, a three-row rectangle
    and a triangle of gene-sized records stay on the merge kernel; results agree either way."""
    monkeypatch.delenv("GKD_ISECT_ALGO", raising=False)
    seqs = _seqs(9, [400_000] * 70)
    with gkd.Engine(k=21) as e:
        for s in seqs:
            e.add(s)
        e.build()
        gi, gd = e.all_vs_all()
        assert e.metrics()["intersect_kernel"] == 5
        qi, qd = e.query_vs_ref([0, 1, 2], [3, 4, 5, 6])
        assert e.metrics()["intersect_kernel"] == 3
    monkeypatch.setenv("GKD_ISECT_ALGO", "merge")
    with gkd.Engine(k=21) as e:
        for s in seqs:
            e.add(s)
        e.build()
        mi, md = e.all_vs_all()
        assert e.metrics()["intersect_kernel"] == 3
    assert np.array_equal(gi, mi) and np.array_equal(gd, md)
    monkeypatch.delenv("GKD_ISECT_ALGO", raising=False)
    small = _seqs(10, [1500] * 80)
    with gkd.Engine(k=21) as e:
        for s in small:
            e.add(s)
        e.build()
        e.all_vs_all()
        assert e.metrics()["intersect_kernel"] == 3
