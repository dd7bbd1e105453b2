"""The `gkd` command mirrors the reference sub-commands: option names, validation messages
(FastaDistanceProcessor.java:98-102, GenomeProcessor.java:84-98), headers and Double.toString text."""
import json
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GKD = os.path.join(ROOT, "genome", "distance_b200", "gkd")


def run(args, stdin=None):
    p = subprocess.run([GKD] + args, input=stdin, capture_output=True, text=True, timeout=300)
    return p.returncode, p.stdout, p.stderr


def test_validation_messages_match_reference(tmp_path):
    assert os.path.exists(GKD), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    rc, out, err = run(["fastaDist", "-K", "1"])
    assert rc == 1 and err.startswith("Kmer size must be at least 2.")
    rc, out, err = run(["fastaDist", "--batch", "0"])
    assert rc == 1 and err.startswith("Batch size must be at least 1.")
    rc, out, err = run(["fastaDist", "-i", str(tmp_path / "nope.fa")])
    assert rc == 1 and "is not found or unreadable." in err
    rc, out, err = run(["fastaDist", "--type", "BOGUS"])
    assert rc == 1 and "is not a valid value for \"--type\"" in err
    rc, out, err = run(["fastaDist", "--frobnicate"])
    assert rc == 1 and "is not a valid option" in err
    d = tmp_path / "g"
    d.mkdir()
    rc, out, err = run(["genomes", "-K", "3", str(d), str(d)])
    assert rc == 1 and err.startswith("Kmer size cannot be less than 4.")
    for bad in ("0", "1.5", "-0.2"):
        rc, out, err = run(["genomes", "-m", bad, str(d), str(d)])
        assert rc == 1 and "Maximum distance must be > 0 and <= 1." in err
    rc, out, err = run(["genomes", str(tmp_path / "missing"), str(d)])
    assert rc == 1 and err.strip().endswith("Main genome source \"%s\" is not found." % (tmp_path / "missing"))
    rc, out, err = run(["genomes", str(d), str(tmp_path / "missing2")])
    assert rc == 1 and "Genome source \"%s\" is not found." % (tmp_path / "missing2") in err
    rc, out, err = run(["genomes", str(d)])
    assert rc == 1 and "is required" in err
    rc, out, err = run(["methodCorr"])
    assert rc == 2 and "Invalid command methodCorr" in err


def _rand(rng, n, alpha="acgt"):
    return "".join(rng.choice(alpha) for _ in range(n))


def _mut(rng, s, rate):
    return "".join(rng.choice("acgt") if rng.random() < rate else c for c in s)


@pytest.mark.gpu
def test_fastadist_report_matches_oracle_text(orc, tmp_path):
    rng = random.Random(3)
    base = _rand(rng, 30000)
    recs = [("g%03d" % i, "synthetic len=%d seed=%d" % (30000, i), _mut(rng, base, 0.01 * i)) for i in range(6)]
    recs.append(("other", "", _rand(rng, 20000)))
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for label, comment, seq in recs:
            f.write(">%s %s\n" % (label, comment))
            for i in range(0, len(seq), 80):
                f.write(seq[i:i + 80] + "\n")
    want = ["seq1\tname1\tseq2\tname2\tdistance"]
    sets = [orc.StrSet(s, 21) for _, _, s in recs]
    for i in range(len(recs)):
        for j in range(i + 1, len(recs)):
            want.append("\t".join([recs[i][0], recs[i][1], recs[j][0], recs[j][1], orc.java_double(sets[i].distance(sets[j]))]))
    rc, out, err = run(["fastaDist", "-i", str(fa)])
    assert rc == 0, err
    got = out.rstrip("\n").split("\n")
    assert got[0] == want[0]
    assert sorted(got[1:]) == sorted(want[1:])  # the reference's row order is nondeterministic
    assert got[-1].endswith("\t1.0")
    assert "7 sequences read from input." in err
    # the same report through a group of contexts (--gpus path; three members on device 0 here): byte for byte
    met = tmp_path / "m.json"
    rc, out3, err = run(["fastaDist", "-i", str(fa), "--devices", "0,0,0", "--metrics", str(met)])
    assert rc == 0, err
    assert out3 == out and "sequences cached on 3 GPUs" in err
    import json as _json
    m = _json.loads(met.read_text())
    assert m["command"] == "fastaDist" and m["gpus"] == 3 and m["pairs"] == 21 and m["launches"] > 0
    rc, _, err = run(["fastaDist", "--gpus", "2"], stdin=">a\nacgt\n")
    assert rc != 0 and "--gpus needs an input file" in err
    # stdin + -o + protein type + explicit K
    prot = ">p1 a\nMKVLAAGIVGLLLAQWERTY\n>p2 b\nMKVLAAGIVGLLLSQWERTY\n"
    outp = tmp_path / "o.tbl"
    rc, out, err = run(["fastaDist", "--type", "PROT", "-K", "8", "-o", str(outp)], stdin=prot)
    assert rc == 0, err
    assert outp.read_text() == "seq1\tname1\tseq2\tname2\tdistance\np1\ta\tp2\tb\t0.7\n"


@pytest.mark.gpu
def test_genomes_report_matches_oracle_text(orc, tmp_path):
    rng = random.Random(8)
    base = [_rand(rng, 20000), _rand(rng, 7000)]

    def gto(dirp, gid, contigs):
        obj = {"id": gid, "scientific_name": "Synthetic " + gid, "features": [{"id": "fig|1", "location": [["c1", 1, "+", 9]]}],
               "contigs": [{"id": "c%d" % i, "dna": c, "genetic_code": 11} for i, c in enumerate(contigs)], "ncbi_taxonomy_id": 2}
        with open(dirp / (gid + ".gto"), "w") as f:
            json.dump(obj, f)

    bdir, qdir = tmp_path / "base", tmp_path / "query"
    bdir.mkdir()
    qdir.mkdir()
    refs = {"100.1": base, "100.2": [_mut(rng, c, 0.02) for c in base], "200.1": [_rand(rng, 15000)]}
    qs = {"300.1": [_mut(rng, c, 0.05) for c in base], "300.2": [_rand(rng, 9000)]}
    for gid, c in refs.items():
        gto(bdir, gid, c)
    for gid, c in qs.items():
        gto(qdir, gid, c)
    rsets = {g: orc.StrSet(c, 21) for g, c in refs.items()}
    want = ["genome1\tgenome2\tdistance"]
    for q in sorted(qs):
        qset = orc.StrSet(qs[q], 21)
        for r in sorted(refs):
            want.append("%s\t%s\t%s" % (q, r, orc.java_double(qset.distance(rsets[r]))))
    rc, out, err = run(["genomes", str(bdir), str(qdir)])
    assert rc == 0, err
    assert out.rstrip("\n").split("\n") == want  # deterministic order in this command
    assert "6 comparisons output." in err
    # downstream contract (DistanceCheckProcessor.java:159-169, GroupTypeSpec.java:84,90): three
    # tab-separated columns, column 3 parses as a Java double, unrelated genomes print exactly 1.0
    rows = [ln.split("\t") for ln in out.rstrip("\n").split("\n")[1:]]
    assert all(len(r) == 3 and 0.0 <= float(r[2]) <= 1.0 for r in rows)
    assert any(r[2] == "1.0" for r in rows)


def test_fastareps_validation():
    rc, out, err = run(["fastaReps", "-K", "1"])
    assert rc == 1 and err.startswith("Kmer size must be at least 2.")
    rc, out, err = run(["fastaReps", "-i", "/definitely/missing.fa"])
    assert rc == 1 and "is not found or invalid." in err


@pytest.mark.gpu
def test_fastareps_matches_reference_greedy(orc, tmp_path):
    """FastaDistanceRepsProcessor.java:117-147: a record is a representative unless a current
    representative is within --dist; output is `seq\\tname` in input order."""
    rng = random.Random(21)
    bases = [_rand(rng, 12000) for _ in range(3)]
    recs = []
    for i in range(14):
        b = bases[i % 3]
        recs.append(("r%02d" % i, "fam%d" % (i % 3), _mut(rng, b, 0.004 * (i // 3))))
    fa = tmp_path / "reps.fa"
    with open(fa, "w") as f:
        for label, comment, seq in recs:
            f.write(">%s %s\n%s\n" % (label, comment, seq))
    sets = [orc.StrSet(s, 21) for _, _, s in recs]
    for max_dist in (0.97, 0.3, 0.15, 0.05):
        reps, want = [], ["seq\tname"]
        for i, (label, comment, _) in enumerate(recs):
            if not any(sets[r].distance(sets[i]) <= max_dist for r in reps):
                reps.append(i)
                want.append(label + "\t" + comment)
        rc, out, err = run(["fastaReps", "-i", str(fa), "--dist", str(max_dist)])
        assert rc == 0, err
        assert out.rstrip("\n").split("\n") == want, max_dist
        assert "%d representatives found for %d sequences." % (len(reps), len(recs)) in err
    assert len(want) > 4  # the tightest threshold splits the families


def test_distreps_validation(tmp_path):
    d = tmp_path / "g"
    d.mkdir()
    rc, out, err = run(["distReps", "-K", "3", str(d)])
    assert rc == 1 and err.startswith("Kmer size must be at least 4.")
    for bad in ("0", "1", "1.5"):
        rc, out, err = run(["distReps", "--dist", bad, str(d)])
        assert rc == 1 and err.startswith("Distance must be strictly between 0 and 1.")
    rc, out, err = run(["distReps", str(tmp_path / "missing")])
    assert rc == 1 and "Genome source %s is not found." % (tmp_path / "missing") in err
    rc, out, err = run(["distReps"])
    assert rc == 1 and "is required" in err


def _name(gid):
    """genome names with non-ASCII characters: json.dump writes them as \\uXXXX escapes (one of them a surrogate
    pair), which the GTO reader must turn back into the UTF-8 bytes a Java PrintWriter would print"""
    return "Genus species " + gid + (" str. \u00e9\u2013\U0001d11e" if gid.endswith(("1", "4")) else "")


@pytest.mark.gpu
def test_distreps_matches_reference_greedy(orc, tmp_path):
    """DistanceRepsProcessor.java:185-274: greedy representatives, closest-representative assignment,
    list and stats files named rep%.4f_K%d.*"""
    rng = random.Random(44)
    bases = [[_rand(rng, 9000), _rand(rng, 4000)] for _ in range(3)]
    genomes = {}
    for i in range(11):
        b = bases[i % 3]
        genomes["%d.%d" % (100 + i % 3, i)] = [_mut(rng, c, 0.01 * (i // 3)) for c in b]
    src = tmp_path / "gtos"
    src.mkdir()
    for gid, contigs in genomes.items():
        with open(src / (gid + ".gto"), "w") as f:
            json.dump({"id": gid, "scientific_name": _name(gid),
                       "contigs": [{"id": "c%d" % j, "dna": c} for j, c in enumerate(contigs)]}, f)
    order = sorted(genomes)  # the source lists genomes in file-name order
    k, max_dist = 12, 0.6
    sets = {g: orc.StrSet(genomes[g], k) for g in order}
    reps = []
    for g in order:
        if not any(sets[r].distance(sets[g]) <= max_dist for r in reps):
            reps.append(g)
    want_list = ["genome_id\tgenome_name\trep_id\trep_name\tdistance"]
    counts = {}
    for g in order:
        if g in reps:
            rep, d = g, 0.0
        else:
            rep, d = None, 1.0
            for r in reps:
                x = sets[g].distance(sets[r])
                if x < d:
                    rep, d = r, x
        counts[rep] = counts.get(rep, 0) + 1
        want_list.append("\t".join([g, _name(g), rep, _name(rep), orc.java_double(d)]))
    out_dir = tmp_path / "out"
    rc, out, err = run(["distReps", "-K", str(k), "--dist", str(max_dist), "-D", str(out_dir), str(src)])
    assert rc == 0, err
    prefix = "rep%.4f_K%d" % (max_dist, k)
    assert (out_dir / (prefix + ".list.tbl")).read_text(encoding="utf-8").rstrip("\n").split("\n") == want_list
    stats = (out_dir / (prefix + ".stats.tbl")).read_text(encoding="utf-8").rstrip("\n").split("\n")
    assert stats[0] == "rep_id\trep_name\tsize"
    got = {ln.split("\t")[0]: int(ln.split("\t")[2]) for ln in stats[1:]}
    assert got == counts and 1 < len(reps) < len(order)
    sizes = [int(ln.split("\t")[2]) for ln in stats[1:]]
    assert sizes == sorted(sizes, reverse=True)
    # --clear erases what an earlier run left in the output directory
    (out_dir / "rep0.5000_K9.list.tbl").write_text("stale\n")
    rc, out, err = run(["distReps", "-K", str(k), "--dist", str(max_dist), "-D", str(out_dir), "--clear", str(src)])
    assert rc == 0, err
    assert sorted(p.name for p in out_dir.iterdir()) == [prefix + ".list.tbl", prefix + ".stats.tbl"]
    assert (out_dir / (prefix + ".list.tbl")).read_text(encoding="utf-8").rstrip("\n").split("\n") == want_list
