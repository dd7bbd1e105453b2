"""numpy port of the synthetic-genome generator (the engine's copy is csrc/synth.cu).

TEST INFRASTRUCTURE: lets the CPU baseline / reference arm of bench.py and the CPU tests produce the
workload's inputs without loading the product library.  Counter-based: residue p of descendant `member`
of ancestor `family` depends only on (seed, family, member, p); tests check it is byte-identical to
gkd_synth_dna / gkd_synth_protein.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
_DNA = np.frombuffer(b"acgt", dtype=np.uint8)
_PROT = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


def _mix64(x):
    """splitmix64 finaliser on uint64 scalars or arrays (wrap-around arithmetic)"""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def _rate_to_u64(rate: float) -> int:
    if rate <= 0:
        return 0
    if rate >= 1:
        return 0xFFFFFFFFFFFFFFFF
    return int(rate * 18446744073709551616.0)


def synth(length: int, seed: int, family: int, member: int, rate: float, protein: bool = False,
          chunk: int = 1 << 22) -> np.ndarray:
    """`length` residues as a uint8 array (lower-case acgt, or the 20 amino-acid letters)"""
    out = np.empty(length, dtype=np.uint8)
    with np.errstate(over="ignore"):
        fam_key = _mix64(np.uint64(seed) ^ ((np.uint64(0xA5A5A5A5) + np.uint64(family)) * np.uint64(0xD6E8FEB86659FD93)))
        mem_key = _mix64(fam_key ^ ((np.uint64(0x5EED0000) + np.uint64(member)) * np.uint64(0xCA5A826395121157)))
        radix = np.uint64(20 if protein else 4)
        rate_u = np.uint64(_rate_to_u64(rate))
        letters = _PROT if protein else _DNA
        for p0 in range(0, length, chunk):
            p = np.arange(p0, min(length, p0 + chunk), dtype=np.uint64)
            code = _mix64(fam_key + p) % radix
            if member != 0 and int(rate_u) != 0:
                u = _mix64(mem_key + p)
                hit = u < rate_u
                delta = np.uint64(1) + _mix64(u ^ mem_key) % (radix - np.uint64(1))
                code = np.where(hit, (code + delta) % radix, code)
            out[p0:p0 + p.size] = letters[code.astype(np.int64)]
    return out
