"""Oracle pinned to the known-answer vectors of SURVEY.md section 8(c).

PARITY UNPINNED: /root/reference holds no tests or golden vectors; these vectors are
hand-derived from the call-site contract, and the three restatements (Python set of strings, C
string hash set, C canonical integer set) are checked against them and against each other.
"""
import json
import os
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kat_survey_8c.json")


def _kats():
    with open(GOLD) as f:
        return json.load(f)


@pytest.mark.parametrize("kat", _kats()["pairs"], ids=lambda k: k["name"])
def test_kat_all_modes(orc, kat):
    alpha = orc.PROT if kat["alphabet"] == "PROT" else orc.DNA
    k, A, B = kat["k"], kat["a"], kat["b"]
    exp = kat["both"]
    # literal python sets
    pa, pb = orc.py_kmer_set(A, k, alpha), orc.py_kmer_set(B, k, alpha)
    inter, dist = orc.py_distance(pa, pb)
    assert [len(pa), len(pb), inter] == exp
    assert orc.java_double(dist) == kat["distance"]
    # C string mode
    sa, sb = orc.StrSet(A, k, alpha), orc.StrSet(B, k, alpha)
    assert [len(sa), len(sb), sa.similarity(sb)] == exp
    assert sa.distance(sb) == dist
    assert sb.similarity(sa) == sa.similarity(sb)
    # C integer mode
    ia, ib = orc.IntSet(A, k, alpha), orc.IntSet(B, k, alpha)
    assert [len(ia), len(ib), ia.similarity(ib)] == exp
    assert ia.distance(ib) == dist
    if "canonical" in kat:
        c, pal = ia.intersect(ib)
        assert [ia.count, ib.count, c] == kat["canonical"]
        if "palindromes" in kat:
            assert [ia.palindromes, ib.palindromes, pal] == kat["palindromes"]


def test_java_double_layout(orc):
    for text in _kats()["double_to_string"]:
        assert orc.java_double(float(text)) == text
    assert orc.java_double(1.0 - 1.0 / 3.0) == "0.6666666666666667"
    assert orc.java_double(0.0) == "0.0"
    assert orc.java_double(1e7) == "1.0E7"
    assert orc.java_double(1234567.0) == "1234567.0"
    assert orc.java_double(1e-3) == "0.001"
    assert orc.java_double(9.999e-4) == "9.999E-4"


def test_distance_formula(orc):
    assert orc.distance(0, 10, 10) == 1.0
    assert orc.distance(10, 10, 10) == 0.0
    assert orc.distance(18, 24, 28) == 1.0 - 18.0 / 34.0
    # Java int wrap of |A|+|B| is restated, not "fixed"
    big = 2 ** 30
    wrapped = float(np.int32(np.uint32(2 * big & 0xFFFFFFFF)))
    assert orc.distance(5, big, big) == 1.0 - 5.0 / (wrapped - 5.0)


def test_canonical_key_encoding(orc):
    import ctypes as C

    L = orc.lib()
    v = C.c_int(0)
    assert L.orc_dna_canonical(b"acgt", 4, C.byref(v)) == 0b00011011 and v.value == 1  # palindrome
    assert L.orc_dna_canonical(b"tttt", 4, C.byref(v)) == 0  # revcomp aaaa is smaller
    assert L.orc_dna_canonical(b"ACGA", 4, C.byref(v)) == 0b00011000
    L.orc_dna_canonical(b"acna", 4, C.byref(v))
    assert v.value == 0
    assert L.orc_dna_revcomp_key(0b00011000, 4) == 0b11011011  # acga -> tcgt


def test_edge_cases(orc):
    # contig shorter than K, empty record, N runs, mixed case
    assert len(orc.StrSet("ACG", 5)) == 0 and orc.IntSet("ACG", 5).count == 0
    assert len(orc.StrSet("", 5)) == 0
    e = orc.StrSet("", 5)
    assert e.distance(orc.StrSet("ACGTACGT", 5)) == 1.0
    s1, s2 = orc.StrSet("ACGTNACGTAC", 4), orc.StrSet("acgtnacgtac", 4)
    assert len(s1) == len(s2) == s1.similarity(s2)
    i1 = orc.IntSet("ACGTNACGTAC", 4)
    assert len(i1) == len(s1)
    # k-mers never span contigs
    two = orc.StrSet(["ACGTT", "GGCAT"], 4)
    one = orc.StrSet("ACGTTGGCAT", 4)
    assert len(two) < len(one)
    assert len(orc.IntSet(["ACGTT", "GGCAT"], 4)) == len(two)
    # RNA reads u as t
    assert len(orc.StrSet("ACGUUGCA", 4, orc.RNA)) == len(orc.StrSet("ACGTTGCA", 4, orc.DNA))
    # literal policy keeps n-containing k-mers, skip policy drops them
    lit = orc.StrSet("ACGTNACGT", 4, orc.DNA, orc.AMBIG_LITERAL)
    assert len(lit) == len(orc.py_kmer_set("ACGTNACGT", 4, orc.DNA, orc.AMBIG_LITERAL))
    assert len(lit) > len(orc.StrSet("ACGTNACGT", 4))


dna = st.text(alphabet="ACGTacgtN", min_size=0, max_size=120)


@settings(max_examples=150, deadline=None)
@given(a=dna, b=dna, k=st.integers(min_value=2, max_value=12))
def test_modes_agree_random(orc, a, b, k):
    pa, pb = orc.py_kmer_set(a, k), orc.py_kmer_set(b, k)
    inter, dist = orc.py_distance(pa, pb)
    sa, sb = orc.StrSet(a, k), orc.StrSet(b, k)
    ia, ib = orc.IntSet(a, k), orc.IntSet(b, k)
    assert (len(sa), len(sb), sa.similarity(sb)) == (len(pa), len(pb), inter)
    assert (len(ia), len(ib), ia.similarity(ib)) == (len(pa), len(pb), inter)
    assert sa.distance(sb) == dist == ia.distance(ib)
    assert sa.distance(sb) == sb.distance(sa)


@settings(max_examples=60, deadline=None)
@given(a=st.text(alphabet="ACGT", min_size=12, max_size=200), k=st.integers(min_value=2, max_value=11))
def test_revcomp_invariance_and_identity(orc, a, k):
    rc = orc.py_revcomp(a.lower())
    s, r = orc.StrSet(a, k), orc.StrSet(rc, k)
    assert s.distance(r) == 0.0 and s.distance(s) == 0.0
    i = orc.IntSet(a, k)
    assert len(s) == 2 * i.count - i.palindromes
    if k % 2 == 1:
        assert i.palindromes == 0


prot = st.text(alphabet="ACDEFGHIKLMNPQRSTVWYX*acd", min_size=0, max_size=80)


@settings(max_examples=80, deadline=None)
@given(a=prot, b=prot, k=st.integers(min_value=1, max_value=8))
def test_protein_modes_agree(orc, a, b, k):
    pa, pb = orc.py_kmer_set(a, k, orc.PROT), orc.py_kmer_set(b, k, orc.PROT)
    inter, dist = orc.py_distance(pa, pb)
    sa, sb = orc.StrSet(a, k, orc.PROT), orc.StrSet(b, k, orc.PROT)
    ia, ib = orc.IntSet(a, k, orc.PROT), orc.IntSet(b, k, orc.PROT)
    assert (len(sa), len(sb), sa.similarity(sb)) == (len(pa), len(pb), inter)
    assert (len(ia), len(ib), ia.similarity(ib)) == (len(pa), len(pb), inter)
    assert sa.distance(sb) == dist == ia.distance(ib)


def test_fasta_dist_command_restatement(orc):
    rng = random.Random(7)
    base = "".join(rng.choice("acgt") for _ in range(3000))
    seqs = []
    for g in range(7):
        s = list(base)
        for p in range(len(s)):
            if rng.random() < 0.02 * g:
                s[p] = rng.choice("acgt")
        seqs.append("".join(s))
    seqs.append("".join(rng.choice("acgt") for _ in range(2500)))  # unrelated -> 1.0
    for batch in (20, 3, 1):  # uncached-column rebuild path must not change results
        for mode in (0, 1):
            inter, dist = orc.fasta_dist(seqs, 15, batch=batch, threads=3, mode=mode)
            t = 0
            for i in range(len(seqs)):
                for j in range(i + 1, len(seqs)):
                    pa, pb = orc.py_kmer_set(seqs[i], 15), orc.py_kmer_set(seqs[j], 15)
                    ei, ed = orc.py_distance(pa, pb)
                    assert (int(inter[t]), dist[t]) == (ei, ed)
                    t += 1
    assert dist[-1] == 1.0
    qi, qd = orc.query_vs_ref(seqs[:3], seqs[3:], 11, threads=2, mode=0)
    qi1, qd1 = orc.query_vs_ref(seqs[:3], seqs[3:], 11, threads=2, mode=1)
    assert (qi == qi1).all() and (qd == qd1).all() and qi.shape == (3, 5)


def test_parse_fasta(orc):
    recs = orc.parse_fasta(">s1 first one\nACGT\nAC\n\n>s2\nGG\r\n>s3   spaced  comment \n")
    assert recs == [("s1", "first one", "ACGTAC"), ("s2", "", "GG"), ("s3", "spaced  comment", "")]


def test_literal_ambiguity_policy_restatements_agree(orc):
    """ORC_AMBIG_LITERAL (recalled upstream behaviour, unpinned): windows with a character outside acgt stay
    in the set as literal strings and so do their reverse-complement windows (unknown base -> 'n').  The C
    string set, the command restatement and the literal Python sets must agree; SKIP differs when N occurs."""
    import random

    rng = random.Random(99)
    seqs = []
    base = "".join(rng.choice("acgt") for _ in range(600))
    for g in range(5):
        s = list(base)
        for _ in range(3 + g):
            p = rng.randrange(len(s))
            run = rng.choice([1, 2, 7, 30])
            s[p:p + run] = rng.choice(["n", "N", "r", "y", "-"]) * run
        seqs.append("".join(s))
    seqs.append(base)
    for k in (5, 8, 21):
        psets = [orc.py_kmer_set(s, k, orc.DNA, orc.AMBIG_LITERAL) for s in seqs]
        ssets = [orc.StrSet(s, k, orc.DNA, orc.AMBIG_LITERAL) for s in seqs]
        assert [len(x) for x in ssets] == [len(x) for x in psets]
        inter, dist = orc.fasta_dist(seqs, k, batch=3, threads=2, mode=0, ambig=orc.AMBIG_LITERAL)
        t = 0
        for i in range(len(seqs)):
            for j in range(i + 1, len(seqs)):
                want = orc.py_distance(psets[i], psets[j])
                assert (int(inter[t]), dist[t]) == want
                assert ssets[i].similarity(ssets[j]) == want[0]
                t += 1
        skip = orc.StrSet(seqs[0], k, orc.DNA, orc.AMBIG_SKIP)
        assert len(skip) < len(ssets[0])
        assert len(orc.StrSet(base, k, orc.DNA, orc.AMBIG_LITERAL)) == len(orc.StrSet(base, k, orc.DNA, orc.AMBIG_SKIP))
    with pytest.raises(MemoryError):
        orc.fasta_dist(seqs, 5, mode=1, ambig=orc.AMBIG_LITERAL)  # integer keys cannot hold literal k-mers


def test_numpy_generator_port_matches_the_engine_generator():
    """oracle/synth.py (what the CPU arm of bench.py uses) is byte-identical to gkd_synth_* (host path)"""
    import numpy as np

    import genome.distance_b200 as gkd
    from oracle import synth as osynth

    for protein in (False, True):
        for fam, mem, rate in ((0, 0, 0.0), (2, 5, 0.05), (7, 1, 0.2), (1, 3, 0.001)):
            a = np.empty(50021, dtype=np.uint8)
            gkd.synth(a, 0x5EED0000, fam, mem, rate, protein=protein)
            assert np.array_equal(a, osynth.synth(50021, 0x5EED0000, fam, mem, rate, protein=protein, chunk=7000))


def test_sketch_restatement_known_answers(orc):
    """String.hashCode and murmur3_x86_32 known answers, and the bottom-w estimator on hand-made signatures"""
    assert orc.java_string_hash("") == 0 and orc.java_string_hash("a") == 97
    assert orc.java_string_hash("hello") == 99162322
    assert orc.java_string_hash("acgtacgtacgtacgtacgta") == orc._s32(sum(ord(c) * 31 ** (20 - i) for i, c in enumerate("acgtacgtacgtacgtacgta")))
    # published murmur3_x86_32 vectors (seed 0)
    assert orc.murmur3_32(b"") == 0
    assert orc.murmur3_32(b"hello") & 0xFFFFFFFF == 0x248BFA47
    assert orc.murmur3_32(b"The quick brown fox jumps over the lazy dog") & 0xFFFFFFFF == 0x2E4FF723
    assert orc.py_sketch_distance([1, 2, 3], [1, 2, 3]) == 0.0
    assert orc.py_sketch_distance([1, 2, 3], [4, 5, 6]) == 1.0
    assert orc.py_sketch_distance([], [1]) == 1.0
    assert orc.py_sketch_distance([-5, 1, 7, 9], [-5, 2, 7, 10]) == 0.5   # union order -5* 1 2 7* 9 10: the w=4 smallest hold 2 matches
    assert orc.py_sketch_distance([-5, 1, 7, 9], [-5, 2, 8, 10]) == 0.75
    assert orc.py_sketch_distance([1, 2], [1, 2, 3, 4]) == 0.0
    ks = orc.py_kmer_set("acgtacgtta", 3)
    hs = orc.py_hash_set(ks, 4)
    assert hs == sorted(hs) and len(hs) == 4 and len(orc.py_hash_set(ks, 1000)) == len({orc.java_string_hash(k) for k in ks})
