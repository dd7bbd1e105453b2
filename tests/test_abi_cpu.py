"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gkd.h declares, formats doubles like Java, and refuses to compute without a GPU."""
import ctypes as C
import os
import random
import re
import struct

import numpy as np
import pytest

import genome.distance_b200 as gkd
from genome.distance_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "gkd.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(gkd_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    lib = gkd.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gkd_abi_version() == _lib.ABI_VERSION == 2


def test_structs_match_header_layout(tmp_path):
    """ctypes mirrors are checked against the C compiler's view of include/gkd.h"""
    import subprocess

    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "gkd.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gkd_config), offsetof(gkd_config, workspace_bytes),'
        ' offsetof(gkd_config, segment_keys), sizeof(gkd_metrics), offsetof(gkd_metrics, keys_unique),'
        ' offsetof(gkd_metrics, intersect_launches), offsetof(gkd_config, ambig_policy), sizeof(gkd_packed_set),'
        ' offsetof(gkd_packed_set, n), offsetof(gkd_packed_set, pal_level), sizeof(gkd_outputs));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    cfg, met, ps = _lib.GkdConfig, _lib.GkdMetrics, _lib.GkdPackedSet
    assert got == [C.sizeof(cfg), cfg.workspace_bytes.offset, cfg.segment_keys.offset, C.sizeof(met),
                   met.keys_unique.offset, met.intersect_launches.offset, cfg.ambig_policy.offset, C.sizeof(ps),
                   ps.n.offset, ps.pal_level.offset, C.sizeof(_lib.GkdOutputs)]


def test_format_double_matches_java_layout(orc):
    rng = random.Random(11)
    vals = [1.0, 0.0, 0.5, 1 - 1 / 3, 1e-3, 9.5e-4, 1e-4, 0.9999999999999999, 1e7, 9999999.0, 123456.789, 1e-10,
            0.47058823529411764, 5e-324, 1.7976931348623157e308, -2.5, 100.0, 1e21, 1e22, 1e23]
    vals += [rng.random() for _ in range(2000)]
    vals += [1.0 - i / (i + rng.randint(1, 10 ** 7)) for i in range(1, 2000)]
    vals += [struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0] for _ in range(2000)]
    for v in vals:
        if v != v or v in (float("inf"), float("-inf")):
            continue
        s = gkd.format_double(v)
        assert s == orc.java_double(v), v
        assert float(s) == v
    for text in ("1.0", "0.0", "9.5E-4", "1.0E-4", "0.001", "1.0E7", "0.6666666666666667", "1234567.0"):
        assert gkd.format_double(float(text)) == text


def test_synth_host_is_deterministic_and_mutates_at_rate():
    a = np.empty(200000, dtype=np.uint8)
    b = np.empty(200000, dtype=np.uint8)
    gkd.synth(a, 7, 2, 0, 0.0)
    gkd.synth(b, 7, 2, 0, 0.0)
    assert (a == b).all() and set(a.tobytes()) == set(b"acgt")
    gkd.synth(b, 7, 2, 5, 0.05)
    frac = float((a != b).mean())
    assert 0.045 < frac < 0.055
    gkd.synth(b, 7, 3, 0, 0.0)
    assert 0.7 < float((a != b).mean()) < 0.8  # another family is unrelated
    p = np.empty(5000, dtype=np.uint8)
    gkd.synth(p, 7, 0, 1, 0.1, protein=True)
    assert set(p.tobytes()) <= set(b"ACDEFGHIKLMNPQRSTVWY")


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(gkd.GkdError) as e:
        gkd.Engine(k=21)
    assert e.value.code == _lib.GKD_ECUDA
    assert "no CPU fallback" in e.value.msg


def test_bad_config_is_einval_or_ecuda():
    # argument validation happens before the device probe
    for kw in (dict(k=33), dict(k=9, alphabet=gkd.PROT), dict(k=-1), dict(alphabet=7), dict(strand_mode=5),
               dict(ambig_policy=3)):
        with pytest.raises(gkd.GkdError) as e:
            gkd.Engine(**kw)
        assert e.value.code == _lib.GKD_EINVAL


def test_jni_glue_compiles_and_matches_the_java_declarations(tmp_path):
    """No JDK in the image: java/jni/gkd_jni.c is compiled against a stub jni.h (types and the JNIEnv members
    it uses) with -Wall -Wextra -Werror, linked against libgkd.so, and must define exactly one JNI function
    per `static native` method declared in GkdNative.java."""
    import subprocess

    so = tmp_path / "libgkd_jni.so"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared",
                           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "java", "jni", "gkd_jni.c"), "-L", os.path.dirname(_lib.LIB_PATH), "-lgkd",
                           "-o", str(so)])
    syms = subprocess.check_output(["nm", "-D", "--defined-only", str(so)]).decode()
    defined = set(re.findall(r"Java_org_theseed_sequence_gpu_GkdNative_(\w+)", syms))
    with open(os.path.join(ROOT, "java", "org", "theseed", "sequence", "gpu", "GkdNative.java")) as f:
        declared = set(re.findall(r"static native [\w\[\]]+ (\w+)\(", f.read()))
    assert declared and declared == defined, declared ^ defined
    # the engine class calls only natives that exist
    with open(os.path.join(ROOT, "java", "org", "theseed", "sequence", "gpu", "GpuKmerEngine.java")) as f:
        used = set(re.findall(r"GkdNative\.(\w+)\(", f.read()))
    assert used <= declared, used - declared
