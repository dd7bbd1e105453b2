"""GPU parity tests proper: libgkd.so (through the C ABI) against the CPU oracle on the same inputs.

Bar: bit-exact set sizes, keys, intersection counts and distances (the distance is computed with the
same double formula, so equality is exact, not a tolerance).  PARITY UNPINNED by the reference itself
(no golden vectors exist); the oracle is pinned to the SURVEY 8(c) known answers in test_oracle.py.
"""
import json
import os
import random

import numpy as np
import pytest

import genome.distance_b200 as gkd

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "kat_survey_8c.json")


def _rand_dna(rng, n, alphabet="acgt"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def _mutate(rng, s, rate, alphabet="acgt"):
    out = list(s)
    for i in range(len(out)):
        if rng.random() < rate:
            out[i] = rng.choice(alphabet)
    return "".join(out)


def _np_family(seed, n_genomes, length, rates, protein=False):
    """family-structured synthetic sequences from the engine's own generator (host path)"""
    seqs = []
    for g in range(n_genomes):
        a = np.empty(length, dtype=np.uint8)
        gkd.synth(a, seed, g % 3, g // 3, rates[g % len(rates)] if g // 3 else 0.0, protein=protein)
        seqs.append(a.tobytes())
    return seqs


@pytest.mark.parametrize("kat", json.load(open(GOLD))["pairs"], ids=lambda k: k["name"])
def test_kat_through_engine(orc, kat):
    alpha = gkd.PROT if kat["alphabet"] == "PROT" else gkd.DNA
    with gkd.Engine(k=kat["k"], alphabet=alpha) as e:
        a, b = e.add(kat["a"]), e.add(kat["b"])
        e.build()
        inter, uni, dist = e.pair(a, b)
        assert [e.set_size(a)[0], e.set_size(b)[0], inter] == kat["both"]
        assert gkd.format_double(dist) == kat["distance"]
        if "canonical" in kat:
            assert [e.set_size(a)[1], e.set_size(b)[1]] == kat["canonical"][:2]
        if "palindromes" in kat:
            assert [e.set_size(a)[2], e.set_size(b)[2]] == kat["palindromes"][:2]
        assert uni == kat["both"][0] + kat["both"][1] - inter


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 11, 15, 16, 20, 21, 22, 27, 31, 32])
def test_dna_sets_bit_exact(orc, k):
    rng = random.Random(100 + k)
    genomes = [
        [_rand_dna(rng, 5000)],
        [_rand_dna(rng, 3000, "ACGTacgtNn"), _rand_dna(rng, 700), "", _rand_dna(rng, k - 1 if k > 1 else 0), _rand_dna(rng, k)],
        ["A" * 400 + "ACGT" * 100 + "T" * 300],                  # low complexity, tandem repeats, palindromes
        [_rand_dna(rng, 9000, "acgtRYKMSWn-*")],                 # IUPAC and junk characters are skipped
        [""],
        [_rand_dna(rng, 33), _rand_dna(rng, 64), _rand_dna(rng, 31), _rand_dna(rng, 4097)],
        [_rand_dna(rng, 70000)],
    ]
    with gkd.Engine(k=k) as e:
        ids = [e.add(g) for g in genomes]
        e.build()
        osets = [orc.IntSet(g, k) for g in genomes]
        for i, o in zip(ids, osets):
            both, canon, pal = e.set_size(i)
            assert (both, canon, pal) == (len(o), o.count, o.palindromes)
            assert np.array_equal(e.export_set(i), o.keys())
        inter, dist = e.all_vs_all()
        t = 0
        for i in range(len(genomes)):
            for j in range(i + 1, len(genomes)):
                assert int(inter[t]) == osets[i].similarity(osets[j]), (i, j)
                assert dist[t] == osets[i].distance(osets[j])
                t += 1
    # string-set restatement agrees on a couple of the genomes too (both-strand sizes)
    s0, s1 = orc.StrSet(genomes[0], k), orc.StrSet(genomes[1], k)
    assert len(s0) == len(osets[0]) and len(s1) == len(osets[1])
    assert s0.similarity(s1) == osets[0].similarity(osets[1])


def test_rna_reads_u_as_t(orc):
    rng = random.Random(5)
    s = _rand_dna(rng, 4000, "acgu")
    with gkd.Engine(k=9, alphabet=gkd.RNA) as e:
        a = e.add(s)
        b = e.add(s.replace("u", "t").upper())
        e.build()
        assert e.pair(a, b)[2] == 0.0
        o = orc.IntSet(s, 9, orc.RNA)
        assert np.array_equal(e.export_set(a), o.keys())


@pytest.mark.parametrize("k", [1, 3, 7, 8])
def test_protein_sets_bit_exact(orc, k):
    rng = random.Random(200 + k)
    aa = "ACDEFGHIKLMNPQRSTVWY"
    proteomes = []
    for g in range(5):
        base = [_rand_dna(rng, rng.randint(0, 400), aa) for _ in range(30)]
        proteomes.append(base)
    proteomes.append([_mutate(rng, p, 0.1, aa) for p in proteomes[0]])
    proteomes.append(["MKV", "", "X*xa" * 20])  # shorter than K, empty, odd characters stay literal
    with gkd.Engine(k=k, alphabet=gkd.PROT) as e:
        ids = [e.add(p) for p in proteomes]
        e.build()
        osets = [orc.IntSet(p, k, orc.PROT) for p in proteomes]
        ssets = [orc.StrSet(p, k, orc.PROT) for p in proteomes]
        for i, o, s in zip(ids, osets, ssets):
            assert e.set_size(i) == (len(s), o.count, 0)
            assert np.array_equal(e.export_set(i), o.keys())
        qi, qd = e.query_vs_ref(ids[:4], ids[4:])
        for a in range(4):
            for b in range(len(ids) - 4):
                assert int(qi[a, b]) == ssets[a].similarity(ssets[4 + b])
                assert qd[a, b] == ssets[a].distance(ssets[4 + b])


def test_all_vs_all_matches_command_oracle(orc):
    seqs = _np_family(42, 12, 150000, [0.001, 0.01, 0.05, 0.2])
    oi, od = orc.fasta_dist(seqs, 21, batch=20, threads=0, mode=1)
    si, sd = orc.fasta_dist(seqs[:5], 21, batch=2, threads=0, mode=0)  # HashSet<String> analogue, rebuild path
    with gkd.Engine(k=21) as e:
        for s in seqs:
            e.add(s)
        e.build()
        gi, gd = e.all_vs_all()
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
        assert (gd == 1.0).any() and (gd < 0.1).any()
        # order-insensitive command semantics: any pair by explicit list gives the same answer
        li, ld = e.pairs([0, 3, 4], [3, 0, 11])
        assert li[0] == li[1] and ld[0] == ld[1]
    with gkd.Engine(k=21) as e:
        for s in seqs[:5]:
            e.add(s)
        e.build()
        gi, gd = e.all_vs_all()
        assert np.array_equal(gi, si) and np.array_equal(gd, sd)


@pytest.mark.parametrize("k", [20, 21])
def test_segmented_merge_path_is_exact(orc, k):
    """Small work items (many (pair, group-run) items per pair) must give the same counts as whole-pair
    items and as the oracle -- including skewed sizes, where the sets meet at the larger one's level."""
    lens = [1200000, 1200000, 90000, 2500000, 40]
    seqs = []
    for g, n in enumerate(lens):
        a = np.empty(n, dtype=np.uint8)
        gkd.synth(a, 9, 0, g, 0.02 if g else 0.0)
        seqs.append(a.tobytes())
    osets = [orc.IntSet(s, k) for s in seqs]
    want_i = [osets[i].similarity(osets[j]) for i in range(len(seqs)) for j in range(i + 1, len(seqs))]
    want_d = [osets[i].distance(osets[j]) for i in range(len(seqs)) for j in range(i + 1, len(seqs))]
    for seg in (0, 2048, 5000, 65536, 1 << 22):
        with gkd.Engine(k=k, segment_keys=seg) as e:
            for s in seqs:
                e.add(s)
            e.build()
            gi, gd = e.all_vs_all()
            assert gi.tolist() == want_i, seg
            assert gd.tolist() == want_d, seg


def test_c1_full_size_pair_properties(orc):
    """BASELINE config 1 at full size: 2 x 5 Mbp, K=21, against the oracle and the domain's invariants."""
    import torch

    n = 5_000_000
    dev = torch.empty((3, n), dtype=torch.uint8, device="cuda")
    gkd.synth(dev[0], 0x5EED0000, 1, 0, 0.0)
    gkd.synth(dev[1], 0x5EED0000, 1, 1, 0.01)
    host = dev[:2].cpu().numpy()
    comp = np.zeros(256, dtype=np.uint8)
    comp[list(b"acgt")] = list(b"tgca")
    rc0 = comp[host[0][::-1]].copy()
    with gkd.Engine(k=21) as e:
        a, b = e.add(dev[0]), e.add(dev[1])       # device-resident inputs
        c = e.add(rc0)                             # host input: reverse complement of genome 0
        e.build()
        oa, ob = orc.IntSet(host[0].tobytes(), 21), orc.IntSet(host[1].tobytes(), 21)
        assert e.set_size(a) == (len(oa), oa.count, 0) and e.set_size(b) == (len(ob), ob.count, 0)
        inter, uni, dist = e.pair(a, b)
        assert inter == oa.similarity(ob) and dist == oa.distance(ob)
        assert e.pair(b, a) == (inter, uni, dist)                    # symmetry
        assert e.pair(a, a)[2] == 0.0                                # identity
        assert e.pair(a, c)[2] == 0.0                                # reverse-complement invariance
        assert np.array_equal(e.export_set(a), oa.keys())
        assert 0.3 < dist < 0.4                                      # ~ 1 - 0.81/(2-0.81)


def test_fasta_file_ingest(orc, tmp_path):
    rng = random.Random(77)
    recs = [("s1", "first record", _rand_dna(rng, 1000, "ACGT")), ("s2", "", _rand_dna(rng, 1300, "acgtn")),
            ("s3", "two  words", ""), ("s4", "x", _rand_dna(rng, 777))]
    text = ""
    for label, comment, seq in recs:
        text += ">" + label + (" " + comment if comment else "") + "\r\n"
        for i in range(0, len(seq), 60):
            text += seq[i:i + 60] + "\n"
    p = tmp_path / "in.fa"
    p.write_text(text)
    parsed = orc.parse_fasta(text)
    assert [(l, c, s) for l, c, s in parsed] == recs
    with gkd.Engine(k=11) as e:
        first, n = e.add_fasta(str(p))
        assert (first, n) == (0, 4)
        assert [(e.label(i), e.comment(i)) for i in range(4)] == [(l, c) for l, c, _ in recs]
        e.build()
        osets = [orc.IntSet(s, 11) for _, _, s in recs]
        for i, o in enumerate(osets):
            assert np.array_equal(e.export_set(i), o.keys())
    with gkd.Engine(k=11) as e:  # whole file as one multi-contig genome
        assert e.add_fasta(str(p), per_record=False) == (0, 1)
        e.build()
        o = orc.IntSet([s for _, _, s in recs], 11)
        assert np.array_equal(e.export_set(0), o.keys())
    with gkd.Engine(k=11) as e:
        with pytest.raises(gkd.GkdError) as err:
            e.add_fasta(str(tmp_path / "missing.fa"))
        assert err.value.code == -2 and "not found or unreadable" in err.value.msg


def test_import_export_and_ranges(orc):
    seqs = _np_family(5, 9, 60000, [0.01, 0.05])
    with gkd.Engine(k=21) as e, gkd.Engine(k=21) as f:
        for s in seqs:
            e.add(s)
        e.build()
        gi, gd = e.all_vs_all()
        # ship the keys to a second context and compare (shuffled and with duplicates: import sorts)
        for i in range(len(seqs)):
            keys = e.export_set(i)
            assert (np.diff(keys.astype(np.int64)) > 0).all()
            f.import_set(np.concatenate([keys[::-1], keys[:7]]))
        fi, fd = f.all_vs_all()
        assert np.array_equal(fi, gi) and np.array_equal(fd, gd)
        # any contiguous slice of the pair enumeration reproduces the same numbers
        total = len(gi)
        for first, count in ((0, total), (0, 1), (5, 17), (total - 3, 3), (total, 0), (8, 1)):
            ri, rd = f.all_vs_all_range(len(seqs), first, count)
            assert np.array_equal(ri, gi[first:first + count]) and np.array_equal(rd, gd[first:first + count])
        with pytest.raises(gkd.GkdError):
            f.all_vs_all_range(len(seqs), total - 1, 5)


def test_state_errors():
    with gkd.Engine(k=21) as e:
        e.add("ACGT" * 100)
        with pytest.raises(gkd.GkdError) as err:
            e.all_vs_all()
        assert err.value.code == -5
        e.build()
        e.add("ACGT" * 50)
        e.add("")
        e.build()  # incremental build
        inter, dist = e.all_vs_all()
        assert dist[-1] == 1.0 and len(inter) == 3
        e.reset()
        assert len(e) == 0


def test_synth_device_matches_host():
    import torch

    for protein in (False, True):
        d = torch.empty(100003, dtype=torch.uint8, device="cuda")
        h = np.empty(100003, dtype=np.uint8)
        gkd.synth(d, 3, 4, 2, 0.07, protein=protein)
        gkd.synth(h, 3, 4, 2, 0.07, protein=protein)
        assert np.array_equal(d.cpu().numpy(), h)


def test_config3_shape_protein_query_vs_reference(orc):
    """BASELINE configs[2] shape at reduced size: per-genome protein 8-mer sets (one piece per protein,
    k-mers never span proteins), queries x references."""
    rng = random.Random(303)

    def proteome(family, member, rate, n_prot=300):
        prots = []
        for p in range(n_prot):
            n = max(50, min(1500, int(rng.lognormvariate(5.5, 0.5))))
            a = np.empty(n, dtype=np.uint8)
            gkd.synth(a, 1000 + p, family, member, rate, protein=True)
            prots.append(a.tobytes())
        return prots

    rng_state = rng.getstate()
    refs, queries = [], []
    for f in range(2):
        rng.setstate(rng_state)  # same protein lengths in every member of a family
        refs.append(proteome(f, 0, 0.0))
    for q in range(5):
        rng.setstate(rng_state)
        queries.append(proteome(q % 2, 1 + q, 0.05 + 0.05 * q))
    with gkd.Engine(k=8, alphabet=gkd.PROT) as e:
        rid = [e.add(p) for p in refs]
        qid = [e.add(p) for p in queries]
        e.build()
        gi, gd = e.query_vs_ref(qid, rid)
    oi, od = np.zeros_like(gi), np.zeros_like(gd)
    rs = [orc.IntSet(p, 8, orc.PROT) for p in refs]
    for a, q in enumerate(queries):
        qs = orc.IntSet(q, 8, orc.PROT)
        for b in range(len(refs)):
            oi[a, b], od[a, b] = qs.similarity(rs[b]), qs.distance(rs[b])
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    assert (gd < 1.0).any() and (gd == 1.0).any()
    s0, s1 = orc.StrSet(queries[0], 8, orc.PROT), orc.StrSet(refs[0], 8, orc.PROT)
    assert gi[0, 0] == s0.similarity(s1) and gd[0, 0] == s0.distance(s1)


def test_config5_shape_skewed_sizes(orc):
    """BASELINE configs[4] shape: plasmid-scale to 12 Mbp genomes in one all-vs-all (uneven set sizes
    exercise the (pair, segment) work split)."""
    import torch

    lens = [100_000, 12_000_000, 350_000, 1_300_000, 5_000_000, 100_000, 2_400_000]
    seqs = []
    for g, n in enumerate(lens):
        t = torch.empty(n, dtype=torch.uint8, device="cuda")
        gkd.synth(t, 55, g % 2, g // 2, 0.01 * (g // 2))
        seqs.append(t)
    with gkd.Engine(k=21) as e:
        for t in seqs:
            e.add(t)
        e.build()
        gi, gd = e.all_vs_all()
    osets = [orc.IntSet(t.cpu().numpy().tobytes(), 21) for t in seqs]
    want_i = [osets[i].similarity(osets[j]) for i in range(len(lens)) for j in range(i + 1, len(lens))]
    want_d = [osets[i].distance(osets[j]) for i in range(len(lens)) for j in range(i + 1, len(lens))]
    assert gi.tolist() == want_i and gd.tolist() == want_d


@pytest.mark.parametrize("k,alpha", [(21, "DNA"), (12, "DNA"), (8, "PROT")])
def test_many_small_sequences_all_vs_all(orc, k, alpha):
    """fastaDist's own use case: hundreds of gene/protein-sized records (tiny sets: every pair is one or a
    few 32-bucket groups at the lowest table level)."""
    rng = random.Random(900 + k)
    letters = "ACDEFGHIKLMNPQRSTVWY" if alpha == "PROT" else "acgt"
    al = gkd.PROT if alpha == "PROT" else gkd.DNA
    oal = orc.PROT if alpha == "PROT" else orc.DNA
    bases = [_rand_dna(rng, rng.randint(300, 3000), letters) for _ in range(12)]
    seqs = []
    for i in range(150):
        s = _mutate(rng, bases[i % 12], 0.01 * (i // 12), letters)
        seqs.append(s[: rng.randint(len(s) // 2, len(s))] if i % 7 == 0 else s)
    seqs += ["", letters[:3], _rand_dna(rng, k, letters)]
    with gkd.Engine(k=k, alphabet=al) as e:
        for s in seqs:
            e.add(s)
        e.build()
        gi, gd = e.all_vs_all()
    oi, od = orc.fasta_dist(seqs, k, alphabet=oal, batch=20, threads=0, mode=1)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    assert (gd < 0.5).any() and (gd == 1.0).any()


def test_kset_cache_and_truncate(orc, tmp_path):
    """persisted sorted-set cache (.kset) round trip, and dropping a streamed panel with truncate"""
    seqs = _np_family(31, 6, 40000, [0.01, 0.05])
    path = str(tmp_path / "panel.kset")
    with gkd.Engine(k=20) as e:  # even K: palindrome lists must survive the round trip
        for s in seqs:
            e.add(s)
        e.build()
        want_i, want_d = e.all_vs_all()
        sizes = [e.set_size(i) for i in range(len(seqs))]
        e.save_sets(path)
    with gkd.Engine(k=20) as f:
        assert f.load_sets(path) == (0, len(seqs))
        assert [f.set_size(i) for i in range(len(seqs))] == sizes
        gi, gd = f.all_vs_all()
        assert np.array_equal(gi, want_i) and np.array_equal(gd, want_d)
        # append the panel again, use it, drop it: ids below the cut stay valid
        first, n = f.load_sets(path)
        assert (first, n) == (len(seqs), len(seqs))
        qi, qd = f.query_vs_ref(list(range(len(seqs))), list(range(first, first + n)))
        assert all(qd[i, i] == 0.0 for i in range(len(seqs)))
        f.truncate(len(seqs))
        assert len(f) == len(seqs)
        gi2, gd2 = f.all_vs_all()
        assert np.array_equal(gi2, want_i) and np.array_equal(gd2, want_d)
    with gkd.Engine(k=21) as g:
        with pytest.raises(gkd.GkdError) as err:
            g.load_sets(path)
        assert err.value.code == -1


def test_fuzz_against_oracle(orc):
    """Randomised differential test: random alphabets, K, contig structures and sizes through the
    C ABI versus the literal Python-set restatement (the most direct reading of HashSet<String>)."""
    rng = random.Random(20261018)
    for trial in range(40):
        prot = trial % 4 == 3
        k = rng.randint(1, 8) if prot else rng.choice([2, 3, 4, 5, 6, 7, 9, 12, 13, 16, 17, 21, 24, 31, 32])
        letters = "ACDEFGHIKLMNPQRSTVWYX" if prot else rng.choice(["acgt", "ACGTacgt", "acgtn", "ac", "acgtRYN"])
        n_genomes = rng.randint(2, 6)
        base = [_rand_dna(rng, rng.randint(0, 400), letters) for _ in range(rng.randint(1, 4))]
        genomes = []
        for g in range(n_genomes):
            if rng.random() < 0.6:
                genomes.append([_mutate(rng, c, rng.choice([0.0, 0.02, 0.2]), letters) for c in base])
            else:
                genomes.append([_rand_dna(rng, rng.randint(0, 300), letters) for _ in range(rng.randint(0, 3))])
        al, oal = (gkd.PROT, orc.PROT) if prot else (gkd.DNA, orc.DNA)
        with gkd.Engine(k=k, alphabet=al) as e:
            for g in genomes:
                e.add(g)
            e.build()
            gi, gd = e.all_vs_all()
            sizes = [e.set_size(i)[0] for i in range(n_genomes)]
        psets = [orc.py_kmer_set(g, k, oal) for g in genomes]
        assert sizes == [len(s) for s in psets], (trial, k, letters)
        t = 0
        for i in range(n_genomes):
            for j in range(i + 1, n_genomes):
                inter, dist = orc.py_distance(psets[i], psets[j])
                assert (int(gi[t]), gd[t]) == (inter, dist), (trial, k, letters, i, j)
                t += 1


@pytest.mark.parametrize("cfg,tmax,table_tmax", [("0", None, None), ("1", None, None), ("2", None, None),
                                                 ("3", None, None), ("4", None, None), ("5", None, None),
                                                 ("0", "64", "16"), ("0", "3", "2"), ("2", "16", "4")])
def test_every_kernel4_configuration_is_exact(orc, cfg, tmax, table_tmax, monkeypatch):
    """GKD_ISECT_CFG pins the stage geometry of kernel 4, GKD_ISECT_TMAX / GKD_TABLE_TMAX the lane target and
    the table resolution (read when a context is created).  Every configuration must be exact on balanced
    AND skewed pairs, whole and split into small work items -- including lane targets that overflow the
    shared-memory stage (tmax 64: the groups are merged straight from global memory), tables finer than
    the walk level, and walk levels capped by a coarse table."""
    monkeypatch.setenv("GKD_ISECT_CFG", cfg)
    if tmax:
        monkeypatch.setenv("GKD_ISECT_TMAX", tmax)
        monkeypatch.setenv("GKD_TABLE_TMAX", table_tmax)
    lens = [600_000, 610_000, 590_000, 30_000, 1_400_000, 25]
    seqs = []
    for g, n in enumerate(lens):
        a = np.empty(n, dtype=np.uint8)
        gkd.synth(a, 123, g % 2, g // 2, 0.03 if g // 2 else 0.0)
        seqs.append(a.tobytes())
    for k in (21, 16, 25):
        osets = [orc.IntSet(s, k) for s in seqs]
        want_i = [osets[i].similarity(osets[j]) for i in range(len(seqs)) for j in range(i + 1, len(seqs))]
        want_d = [osets[i].distance(osets[j]) for i in range(len(seqs)) for j in range(i + 1, len(seqs))]
        for seg in (0, 3000, 100_000):
            with gkd.Engine(k=k, segment_keys=seg) as e:
                for s in seqs:
                    e.add(s)
                e.build()
                gi, gd = e.all_vs_all()
            assert gi.tolist() == want_i and gd.tolist() == want_d, (cfg, tmax, k, seg)


def test_containment_outputs(orc):
    seqs = _np_family(77, 6, 50000, [0.01, 0.1])
    for k in (21, 12):
        with gkd.Engine(k=k) as e:
            for s in seqs:
                e.add(s)
            e.build()
            a = [i for i in range(len(seqs)) for j in range(len(seqs))]
            b = [j for i in range(len(seqs)) for j in range(len(seqs))]
            inter, dist, ca, cb = e.pairs_ex(a, b)
            sizes = [e.set_size(i)[0] for i in range(len(seqs))]
        for t, (i, j) in enumerate(zip(a, b)):
            assert ca[t] == int(inter[t]) / sizes[i] and cb[t] == int(inter[t]) / sizes[j]
            if i == j:
                assert ca[t] == 1.0 and dist[t] == 0.0


def test_packed_sets_are_adopted_in_place(orc):
    """the multi-GPU exchange path on one device: describe the arenas of one context, copy their bytes,
    adopt them in a second context without unpacking, and get the same answers"""
    import torch

    from genome.distance_b200 import sharding

    seqs = _np_family(5, 9, 60000, [0.01, 0.05])
    for k in (21, 20):  # even K: the palindrome sub-sets travel too
        with gkd.Engine(k=k, workspace_bytes=4 << 20) as e, gkd.Engine(k=k) as f:
            for s in seqs:
                e.add(s)
            e.build()
            assert len(e.arenas()) > 1  # the small workspace forces several build batches
            gi, gd = e.all_vs_all()
            meta = e.arena_meta()
            bufs = []
            for p in sharding.plan_panels(meta, 0, len(seqs), 2):
                buf = e.arena_view(p["arena"], p["begin"], p["end"]).clone()
                torch.cuda.synchronize()
                first = f.adopt_sets(buf, p["table"])
                assert first == p["first"]
                bufs.append(buf)
            assert [f.set_size(i) for i in range(len(seqs))] == [e.set_size(i) for i in range(len(seqs))]
            fi, fd = f.all_vs_all()
            assert np.array_equal(fi, gi) and np.array_equal(fd, gd)
            assert np.array_equal(f.export_set(3), e.export_set(3))
            f.truncate(4)
            hi, hd = f.all_vs_all()
            assert np.array_equal(hi, gi[[0, 1, 2, 8, 9, 15]]) and len(f) == 4
            # a descriptor that does not fit the buffer is refused
            bad = sharding.plan_panels(meta, 0, 1, 1)[0]["table"].copy()
            bad["lows_off"] += np.uint64(1 << 40)
            with pytest.raises(gkd.GkdError) as err:
                f.adopt_sets(bufs[0], bad)
            assert err.value.code == -1


def test_literal_ambiguity_policy(orc):
    """GKD_AMBIG_LITERAL against the oracle's literal string sets (both policies are explicit switches: the
    reference does not pin what DnaKmers does with non-acgt characters).  Host strings, numpy and device
    inputs, multi-contig genomes, RNA, odd and even K, and the explicit-pair and rectangular calls."""
    import torch

    rng = random.Random(4242)
    base = _rand_dna(rng, 3000)
    genomes = []
    for g in range(6):
        s = list(_mutate(rng, base, 0.01 * g))
        for _ in range(2 + g):
            p = rng.randrange(len(s))
            run = rng.choice([1, 3, 25, 60])
            s[p:p + run] = rng.choice(["n", "N", "R", "y", "-", "x"]) * run
        s = "".join(s)
        genomes.append([s[:1700], s[1700:]] if g % 2 else [s])
    genomes.append([base])                       # no ambiguity at all
    genomes.append(["nnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnnn"])  # nothing but ambiguity
    genomes.append(["acgtn"])                     # shorter than most K
    for k in (4, 9, 21, 22):
        ssets = [orc.StrSet(g, k, orc.DNA, orc.AMBIG_LITERAL) for g in genomes]
        with gkd.Engine(k=k, ambig_policy=gkd.AMBIG_LITERAL) as e:
            for i, g in enumerate(genomes):
                if i == 2:    # device text
                    e.add([torch.tensor(list(c.encode()), dtype=torch.uint8, device="cuda") for c in g])
                elif i == 3:  # numpy text
                    e.add([np.frombuffer(c.encode(), dtype=np.uint8) for c in g])
                else:
                    e.add(g)
            e.build()
            assert [e.set_size(i)[0] for i in range(len(genomes))] == [len(s) for s in ssets], k
            gi, gd = e.all_vs_all()
            t = 0
            for i in range(len(genomes)):
                for j in range(i + 1, len(genomes)):
                    I = ssets[i].similarity(ssets[j])
                    assert int(gi[t]) == I and gd[t] == orc.distance(I, len(ssets[i]), len(ssets[j])), (k, i, j)
                    t += 1
            qi, qd = e.query_vs_ref([0, 1, 7], [2, 7, 6])
            assert int(qi[0, 0]) == ssets[0].similarity(ssets[2]) and qd[2, 1] == 0.0
            inter, dist, ca, cb = e.pairs_ex([0, 7], [1, 7])
            assert ca[0] == int(inter[0]) / len(ssets[0]) and cb[1] == 1.0
            with pytest.raises(gkd.GkdError):  # literal k-mers live on the host: not exchanged
                e.describe_sets(0, 1)
        # the default policy on the same input drops those k-mers
        with gkd.Engine(k=k) as e:
            e.add(genomes[0])
            e.build()
            assert e.set_size(0)[0] == len(orc.StrSet(genomes[0], k, orc.DNA, orc.AMBIG_SKIP)) < len(ssets[0])
    s = _rand_dna(rng, 500, "acgun")
    with gkd.Engine(k=7, alphabet=gkd.RNA, ambig_policy=gkd.AMBIG_LITERAL) as e:
        e.add(s)
        e.add(s.replace("u", "t"))
        e.build()
        assert e.set_size(0)[0] == len(orc.StrSet(s, 7, orc.RNA, orc.AMBIG_LITERAL)) and e.pair(0, 1)[2] == 0.0
    # the command restatement with the literal policy (oracle string mode) against the engine
    seqs = ["".join(c) for c in genomes[:6]]
    oi, od = orc.fasta_dist(seqs, 21, batch=2, threads=0, mode=0, ambig=orc.AMBIG_LITERAL)
    with gkd.Engine(k=21, ambig_policy=gkd.AMBIG_LITERAL) as e:
        for x in seqs:
            e.add(x)
        e.build()
        gi, gd = e.all_vs_all()
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


@pytest.mark.parametrize("kind", [0, 1])
def test_minhash_sketches_match_restatement(orc, kind):
    """SequenceKmers.hashSet(width) and Sketch.distance (hash unpinned, switchable):
    device sketches against the literal Python restatement -- DNA both strands (odd and even K, a set
    smaller than the width = a "dwarf"), RNA, protein, both string hashes, several widths."""
    rng = random.Random(31337 + kind)
    base = _rand_dna(rng, 40000)
    dna = [base, _mutate(rng, base, 0.02), _mutate(rng, base, 0.2), _rand_dna(rng, 30000), _rand_dna(rng, 150), "acgtn"]
    for k, width in ((21, 360), (8, 100), (12, 2000), (4, 4096)):
        psets = [orc.py_kmer_set(s, k) for s in dna]
        want = [orc.py_hash_set(p, width, kind) for p in psets]
        with gkd.Engine(k=k) as e:
            for s in dna:
                e.add(s)
            e.build()
            got = [e.hash_set(i, width, kind).tolist() for i in range(len(dna))]
            assert got == want, (k, width)
            a = [i for i in range(len(dna)) for j in range(len(dna))]
            b = [j for i in range(len(dna)) for j in range(len(dna))]
            d = e.sketch_distances(width, a, b, kind)
            for t, (i, j) in enumerate(zip(a, b)):
                assert d[t] == orc.py_sketch_distance(want[i], want[j]), (k, width, i, j)
            # the estimate tracks the exact distance for a mutated descendant (sanity, not parity)
            if k == 21:
                exact = e.pair(0, 1)[2]
                assert abs(d[1] - exact) < 0.15
    aa = "ACDEFGHIKLMNPQRSTVWY"
    prots = [_rand_dna(rng, 3000, aa), _rand_dna(rng, 50, aa)]
    prots.append(_mutate(rng, prots[0], 0.05, aa))
    for k in (8, 3):
        want = [orc.py_hash_set(orc.py_kmer_set(p, k, orc.PROT), 360, kind) for p in prots]
        with gkd.Engine(k=k, alphabet=gkd.PROT) as e:
            for p in prots:
                e.add(p)
            e.build()
            assert [e.hash_set(i, 360, kind).tolist() for i in range(3)] == want
            assert e.sketch_distances(360, [0, 0], [2, 1], kind).tolist() == [orc.py_sketch_distance(want[0], want[2]),
                                                                                orc.py_sketch_distance(want[0], want[1])]
    with gkd.Engine(k=9, alphabet=gkd.RNA) as e:
        s = _rand_dna(rng, 2000, "acgu")
        e.add(s)
        e.build()
        assert e.hash_set(0, 50, kind).tolist() == orc.py_hash_set(orc.py_kmer_set(s, 9, orc.RNA), 50, kind)
        with pytest.raises(gkd.GkdError):
            e.hash_set(0, 5000, kind)


def test_group_of_contexts_matches_single_context(orc):
    """gkd_group_*: several contexts in one process (here all on device 0; on a multi-GPU box one per device),
    one host thread per member, set arenas pulled with peer copies and adopted in place.  Results are indexed
    by global id exactly like the single-context call -- odd and even member counts (the half-way block of an
    even ring is split), ragged slices, several arenas per member, even K (palindrome sub-sets travel too)."""
    import torch

    rng = random.Random(808)
    base = _rand_dna(rng, 50000)
    seqs = [_mutate(rng, base, 0.01 * (i % 5)) if i % 4 else _rand_dna(rng, 20000 + 3000 * i) for i in range(13)]
    seqs[5] = ""
    for k in (21, 20):
        with gkd.Engine(k=k) as e:
            for s in seqs:
                e.add(s)
            e.build()
            want_i, want_d = e.all_vs_all()
        ndev = torch.cuda.device_count()
        for members in (2, 3, 4):
            devices = [m % ndev for m in range(members)]
            with gkd.Group(devices, k=k, workspace_bytes=4 << 20, panel=2) as g:
                per = (len(seqs) + members - 1) // members
                for i, s in enumerate(seqs):
                    assert g.add(min(i // per, members - 1), s) == i
                g.build()
                gi, gd = g.all_vs_all()
            assert np.array_equal(gi, want_i) and np.array_equal(gd, want_d), (k, members)
    with gkd.Group([0, 0], k=21) as g:
        g.add(1, "acgtacgtacgtacgtacgtacgtacgt")
        with pytest.raises(gkd.GkdError):  # members own contiguous id blocks: member 0 can no longer be filled
            g.add(0, "acgt")


def test_contexts_on_every_visible_device_in_one_process(orc):
    """function attributes (opt-in shared memory) are per device: every context sets them for its own device,
    so a process may hold contexts on all GPUs at once (one per device here, used alternately)"""
    import torch

    seqs = _np_family(9, 4, 300000, [0.02])
    want = None
    engines = [gkd.Engine(k=21, device=d) for d in range(torch.cuda.device_count())]
    try:
        for e in engines:
            for s in seqs:
                e.add(s)
        for e in engines:
            e.build()
        for e in engines:
            gi, gd = e.all_vs_all()
            if want is None:
                osets = [orc.IntSet(s, 21) for s in seqs]
                want = [osets[i].similarity(osets[j]) for i in range(4) for j in range(i + 1, 4)]
            assert gi.tolist() == want
    finally:
        for e in engines:
            e.close()


@pytest.mark.parametrize("algo", ["msd", "lsd"])
def test_both_kernel3_paths_build_identical_sets(orc, algo, monkeypatch):
    """kernel 3 has a fast path (MSD partition by the top bits of the mixed key + one shared-memory sort per bin,
    decoupled look-back for the output position) and the LSD radix path it falls back to.  GKD_SORT_ALGO pins
    one; both must produce the oracle's sets -- random genomes of several sizes (one bin ... thousands of
    bins), mostly-invalid text, a heavily repeated k-mer that overflows a bin (forces the fallback inside the
    MSD path), empty input, protein, and a batch that mixes all of them."""
    monkeypatch.setenv("GKD_SORT_ALGO", algo)
    rng = random.Random(77)
    big = np.empty(3_000_000, dtype=np.uint8)
    gkd.synth(big, 5, 1, 0, 0.0)
    genomes = [
        [big.tobytes()],
        [_rand_dna(rng, 70_000)],
        [_rand_dna(rng, 9000), _rand_dna(rng, 5000, "acgtn"), ""],
        ["n" * 50_000 + _rand_dna(rng, 300) + "n" * 50_000],          # 100,000 slots, a few hundred valid
        [""],
        [_rand_dna(rng, 4095 + 20), _rand_dna(rng, 4096 + 21)],        # bin-count boundaries
    ]
    repeats = [["a" * 40_000 + _rand_dna(rng, 20_000) + "acgt" * 5000]]  # > 5120 copies of one k-mer in one bin
    for k, batch in ((21, genomes), (21, genomes + repeats), (13, genomes[1:] + repeats), (5, genomes[1:3])):
        with gkd.Engine(k=k) as e:
            for g in batch:
                e.add(g)
            e.build()
            for i, g in enumerate(batch):
                o = orc.IntSet(g, k)
                assert e.set_size(i) == (len(o), o.count, o.palindromes), (algo, k, i)
                assert np.array_equal(e.export_set(i), o.keys()), (algo, k, i)
            gi, gd = e.all_vs_all()
            assert gd[0] == orc.IntSet(batch[0], k).distance(orc.IntSet(batch[1], k))
    aa = "ACDEFGHIKLMNPQRSTVWY"
    prots = [[_rand_dna(rng, 60_000, aa)], [_rand_dna(rng, 300, aa) for _ in range(40)], ["MKV"]]
    for k in (8, 5, 2):
        with gkd.Engine(k=k, alphabet=gkd.PROT) as e:
            for p in prots:
                e.add(p)
            e.build()
            for i, p in enumerate(prots):
                o = orc.IntSet(p, k, orc.PROT)
                assert e.set_size(i)[1] == o.count and np.array_equal(e.export_set(i), o.keys()), (algo, k, i)


@pytest.mark.parametrize("k,alpha", [(21, "DNA"), (12, "DNA"), (8, "PROT")])
def test_greedy_representatives_on_device(orc, k, alpha):
    """gkd_greedy_reps (pass 1 of DistanceRepsProcessor.java:185-201 / FastaDistanceRepsProcessor.java:124-146): the
    device-resident greedy pass must pick exactly the representatives the literal loop over oracle distances picks,
    for several thresholds and visiting orders (odd and even K, protein)."""
    rng = random.Random(5150 + k)
    letters = "ACDEFGHIKLMNPQRSTVWY" if alpha == "PROT" else "acgt"
    al, oal = (gkd.PROT, orc.PROT) if alpha == "PROT" else (gkd.DNA, orc.DNA)
    bases = [_rand_dna(rng, rng.randint(3000, 40000), letters) for _ in range(5)]
    seqs = [_mutate(rng, bases[i % 5], 0.004 * (i // 5), letters) for i in range(40)] + ["", letters[:2]]
    osets = [orc.IntSet(s, k, oal) for s in seqs]
    with gkd.Engine(k=k, alphabet=al) as e:
        for s in seqs:
            e.add(s)
        e.build()
        for max_dist in (0.05, 0.3, 0.6, 0.97):
            for order in (list(range(len(seqs))), list(range(len(seqs)))[::-1], rng.sample(range(len(seqs)), 25)):
                reps, want = [], []
                for g in order:
                    found = any(osets[r].distance(osets[g]) <= max_dist for r in reps)
                    want.append(0 if found else 1)
                    if not found:
                        reps.append(g)
                assert e.greedy_reps(order, max_dist).tolist() == want, (k, max_dist)
        assert 1 < sum(e.greedy_reps(list(range(len(seqs))), 0.3)) < len(seqs)
