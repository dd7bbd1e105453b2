"""Thin Python host layer over the C ABI: device memory plumbing for tests, bench and multi-GPU
sharding.  The product is libgkd.so; nothing here computes."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (AMBIG_LITERAL, AMBIG_SKIP, DNA, PROT, RNA, STRAND_BOTH, STRAND_CANONICAL, GkdConfig, GkdMetrics,
                   GkdOutputs, GkdPackedSet)

PACKED_DTYPE = np.dtype([("offs_off", "<u8"), ("lows_off", "<u8"), ("pal_offs_off", "<u8"), ("pal_lows_off", "<u8"),
                         ("n", "<u4"), ("n_pal", "<u4"), ("level", "<u4"), ("pal_level", "<u4")])
assert PACKED_DTYPE.itemsize == C.sizeof(GkdPackedSet)


class GkdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gkd error {code}: {msg}")
        self.code = code
        self.msg = msg


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class Engine:
    """One gkd context on one GPU.

    Mirrors the reference objects at batch granularity: `add` = `KmerType.createKmers` /
    `new GenomeKmers(genome)` / `new ProteinKmers(str)`; `all_vs_all` = FastaDistanceProcessor's pair
    loop; `query_vs_ref` = GenomeProcessor's; `pair` = `SequenceKmers.distance`.
    """

    def __init__(self, k: int = 0, alphabet: int = DNA, strand_mode: int = STRAND_BOTH, device: int = 0,
                 workspace_bytes: int = 0, segment_keys: int = 0, ambig_policy: int = AMBIG_SKIP):
        self._L = _lib.load()
        cfg = GkdConfig(device=device, k=k, alphabet=alphabet, strand_mode=strand_mode,
                        workspace_bytes=workspace_bytes, segment_keys=segment_keys, ambig_policy=ambig_policy)
        h = C.c_void_p()
        rc = self._L.gkd_create(C.byref(h), C.byref(cfg))
        if rc:
            raise GkdError(rc, (self._L.gkd_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self.alphabet = alphabet
        self._keep: List[object] = []
        self._adopted: List[Tuple[int, object]] = []  # (first id, buffer) of adopted panels

    # -- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc:
            raise GkdError(rc, (self._L.gkd_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.gkd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def _ptr_len(x) -> Tuple[int, int, object]:
        """(address, byte length, keep-alive) of str / bytes / numpy uint8 / torch uint8 (cpu or cuda)."""
        if isinstance(x, str):
            x = x.encode("latin-1")
        if isinstance(x, (bytes, bytearray)):
            buf = C.create_string_buffer(bytes(x), len(x)) if len(x) else C.create_string_buffer(1)
            return C.addressof(buf), len(x), buf
        if isinstance(x, np.ndarray):
            a = np.ascontiguousarray(x).view(np.uint8)
            return a.ctypes.data, a.size, a
        if _is_torch(x):
            t = x.contiguous()
            return t.data_ptr(), t.numel() * t.element_size(), t
        raise TypeError(f"unsupported sequence type {type(x)}")

    # -- ingest ------------------------------------------------------------------------------------
    def add(self, contigs) -> int:
        """Add one genome / record; `contigs` is one sequence or a list of contigs (or proteins)."""
        if isinstance(contigs, (str, bytes, bytearray, np.ndarray)) or _is_torch(contigs):
            contigs = [contigs]
        triples = [self._ptr_len(c) for c in contigs]
        n = len(triples)
        ptrs = (C.c_void_p * max(n, 1))(*[t[0] for t in triples])
        lens = (C.c_uint64 * max(n, 1))(*[t[1] for t in triples])
        out = C.c_uint32()
        if any(_is_torch(c) and c.is_cuda for c in contigs):
            import torch

            torch.cuda.synchronize(self.device)  # producers ran on torch's stream
        self._ck(self._L.gkd_add_sequences(self._h, ptrs, lens, n, C.byref(out)))
        # pinned / device text is read asynchronously and must stay alive until build() (gkd.h)
        self._keep.extend(t[2] for t in triples if _is_torch(t[2]))
        return out.value

    def add_fasta(self, path: str, per_record: bool = True) -> Tuple[int, int]:
        first, n = C.c_uint32(), C.c_uint32()
        self._ck(self._L.gkd_add_fasta_file(self._h, path.encode(), 1 if per_record else 0, C.byref(first), C.byref(n)))
        return first.value, n.value

    def label(self, i: int) -> str:
        return self._L.gkd_label(self._h, i).decode("latin-1")

    def comment(self, i: int) -> str:
        return self._L.gkd_comment(self._h, i).decode("latin-1")

    def __len__(self) -> int:
        return self._L.gkd_count(self._h)

    # -- sets --------------------------------------------------------------------------------------
    def build(self):
        self._ck(self._L.gkd_build_sets(self._h))
        self._keep.clear()

    def set_size(self, i: int) -> Tuple[int, int, int]:
        """(reference HashSet size, canonical count, palindromes)"""
        a, b, p = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._L.gkd_set_size(self._h, i, C.byref(a), C.byref(b), C.byref(p)))
        return a.value, b.value, p.value

    def export_set(self, i: int) -> np.ndarray:
        n = self.set_size(i)[1]
        out = np.empty(n, dtype=np.uint64)
        got = C.c_uint64()
        self._ck(self._L.gkd_export_set(self._h, i, out.ctypes.data if n else None, n, C.byref(got)))
        return out

    # -- set exchange (multi-GPU) --------------------------------------------------------------------
    def arenas(self) -> List[Tuple[int, int, int, int]]:
        """[(first_id, n_sets, device address, bytes)] of every live set arena, in id order"""
        out = []
        for a in range(self._L.gkd_arena_count(self._h)):
            first, n, base, nbytes = C.c_uint32(), C.c_uint32(), C.c_void_p(), C.c_uint64()
            self._ck(self._L.gkd_arena_info(self._h, a, C.byref(first), C.byref(n), C.byref(base), C.byref(nbytes)))
            out.append((first.value, n.value, base.value or 0, nbytes.value))
        return out

    def describe_sets(self, first_id: int, n_sets: int) -> np.ndarray:
        """packed layout (PACKED_DTYPE, offsets relative to the arena base) of consecutive sets of one arena"""
        table = np.zeros(n_sets, dtype=PACKED_DTYPE)
        self._ck(self._L.gkd_describe_sets(self._h, first_id, n_sets,
                                           table.ctypes.data_as(C.POINTER(GkdPackedSet))))
        return table

    def arena_meta(self):
        """[(first_id, n_sets, bytes, table)] of every live arena: what sharding.plan_panels consumes"""
        return [(first, n, nbytes, self.describe_sets(first, n)) for first, n, _, nbytes in self.arenas()]

    def arena_view(self, arena: int, begin: int, end: int):
        """zero-copy uint8 torch view of bytes [begin, end) of a set arena on this context's device
        (aliases engine memory: send it, do not write it)"""
        import torch

        base = self.arenas()[arena][2]
        size = max(end - begin, 1)

        class _Raw:
            __cuda_array_interface__ = {"shape": (size,), "typestr": "|u1", "data": (base + begin, False), "version": 3}

        return torch.as_tensor(_Raw(), device=f"cuda:{self.device}")

    def adopt_sets(self, buf, table: np.ndarray) -> int:
        """Register the sets described by `table` whose bytes are in `buf` (uint8 torch tensor on this
        device, e.g. a received panel) WITHOUT copying; `buf` is kept alive until truncate()/reset().
        The bytes must already be there (synchronise with the producer first)."""
        t = np.ascontiguousarray(table, dtype=PACKED_DTYPE)
        out = C.c_uint32()  # the caller has synchronised with whatever filled `buf`
        self._ck(self._L.gkd_adopt_sets(self._h, buf.data_ptr(), buf.numel(), t.ctypes.data_as(C.POINTER(GkdPackedSet)),
                                        t.size, C.byref(out)))
        self._adopted.append((out.value, buf))
        return out.value

    def import_set(self, keys) -> int:
        """Add a set from its key array (numpy uint64 on host or torch int64/uint64 on this device); copied."""
        if isinstance(keys, np.ndarray):
            a = np.ascontiguousarray(keys, dtype=np.uint64)
            ptr, n, keep = a.ctypes.data, a.size, a
        else:
            import torch

            t = keys.contiguous()
            if t.is_cuda:
                torch.cuda.synchronize(self.device)
            ptr, n, keep = t.data_ptr(), t.numel(), t
        out = C.c_uint32()
        self._ck(self._L.gkd_import_set(self._h, ptr if n else None, n, C.byref(out)))
        del keep
        return out.value

    def import_sets(self, keys, offsets) -> int:
        """Add many sets from key arrays stored back to back (`keys`: numpy uint64 or torch int64 tensor,
        host or this device; `offsets`: n+1 host integers).  Returns the id of the first new set."""
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        if isinstance(keys, np.ndarray):
            a = np.ascontiguousarray(keys, dtype=np.uint64)
            ptr, keep = a.ctypes.data, a
        else:
            import torch

            t = keys.contiguous()
            if t.is_cuda:
                torch.cuda.synchronize(self.device)
            ptr, keep = t.data_ptr(), t
        out = C.c_uint32()
        self._ck(self._L.gkd_import_sets(self._h, ptr if int(offs[-1]) else None,
                                         offs.ctypes.data_as(C.POINTER(C.c_uint64)), offs.size - 1, C.byref(out)))
        del keep
        return out.value

    # -- distances ---------------------------------------------------------------------------------
    @staticmethod
    def _outs(n: int, inter_out, dist_out, want_inter: bool, want_dist: bool):
        inter = inter_out if inter_out is not None else (np.empty(n, dtype=np.uint64) if want_inter else None)
        dist = dist_out if dist_out is not None else (np.empty(n, dtype=np.float64) if want_dist else None)

        def addr(x):
            if x is None:
                return None
            return x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr()

        return inter, dist, addr(inter), addr(dist)

    def all_vs_all(self, inter_out=None, dist_out=None, want_inter=True, want_dist=True):
        n = len(self)
        npairs = n * (n - 1) // 2
        inter, dist, pi, pd = self._outs(npairs, inter_out, dist_out, want_inter, want_dist)
        self._ck(self._L.gkd_all_vs_all(self._h, pi, pd))
        return inter, dist

    def all_vs_all_range(self, n: int, first: int, count: int, inter_out=None, dist_out=None):
        inter, dist, pi, pd = self._outs(count, inter_out, dist_out, True, True)
        self._ck(self._L.gkd_all_vs_all_range(self._h, n, first, count, pi, pd))
        return inter, dist

    def query_vs_ref(self, q: Sequence[int], r: Sequence[int]):
        qa, ra = np.ascontiguousarray(q, dtype=np.uint32), np.ascontiguousarray(r, dtype=np.uint32)
        inter, dist, pi, pd = self._outs(qa.size * ra.size, None, None, True, True)
        self._ck(self._L.gkd_query_vs_ref(self._h, qa.ctypes.data, qa.size, ra.ctypes.data, ra.size, pi, pd))
        return inter.reshape(qa.size, ra.size), dist.reshape(qa.size, ra.size)

    def pairs(self, a: Sequence[int], b: Sequence[int], inter_out=None, dist_out=None):
        aa, ba = np.ascontiguousarray(a, dtype=np.uint32), np.ascontiguousarray(b, dtype=np.uint32)
        inter, dist, pi, pd = self._outs(aa.size, inter_out, dist_out, True, True)
        self._ck(self._L.gkd_pairs(self._h, aa.ctypes.data, ba.ctypes.data, aa.size, pi, pd))
        return inter, dist

    def pairs_ex(self, a: Sequence[int], b: Sequence[int]):
        """(inter, dist, contain_a, contain_b) of an explicit pair list (gkd_pairs_ex)"""
        aa, ba = np.ascontiguousarray(a, dtype=np.uint32), np.ascontiguousarray(b, dtype=np.uint32)
        n = aa.size
        inter, dist = np.empty(n, dtype=np.uint64), np.empty(n, dtype=np.float64)
        ca, cb = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)
        o = GkdOutputs(inter.ctypes.data, dist.ctypes.data, ca.ctypes.data, cb.ctypes.data)
        self._ck(self._L.gkd_pairs_ex(self._h, aa.ctypes.data, ba.ctypes.data, n, C.byref(o)))
        return inter, dist, ca, cb

    # -- MinHash sketches --------------------------------------------------------------------------
    def hash_set(self, i: int, width: int, hash_kind: int = 0) -> np.ndarray:
        """SequenceKmers.hashSet(width): the smallest distinct hash codes of set i, ascending, as int32"""
        out = np.empty(width, dtype=np.int32)
        n = C.c_uint32()
        self._ck(self._L.gkd_hash_set(self._h, i, width, hash_kind, out.ctypes.data, C.byref(n)))
        return out[: n.value].copy()

    def sketch_distances(self, width: int, a: Sequence[int], b: Sequence[int], hash_kind: int = 0) -> np.ndarray:
        """Sketch.distance for a pair list (sketches built and compared on the device)"""
        aa, ba = np.ascontiguousarray(a, dtype=np.uint32), np.ascontiguousarray(b, dtype=np.uint32)
        dist = np.empty(aa.size, dtype=np.float64)
        self._ck(self._L.gkd_sketch_distances(self._h, width, hash_kind, aa.ctypes.data, ba.ctypes.data, aa.size,
                                              dist.ctypes.data))
        return dist

    def greedy_reps(self, order: Sequence[int], max_dist: float) -> np.ndarray:
        """greedy representative pass on the device: flags[i] = 1 iff order[i] became a representative"""
        o = np.ascontiguousarray(order, dtype=np.uint32)
        flags = np.zeros(o.size, dtype=np.uint8)
        self._ck(self._L.gkd_greedy_reps(self._h, o.ctypes.data, o.size, float(max_dist), flags.ctypes.data))
        return flags

    def pair(self, a: int, b: int) -> Tuple[int, int, float]:
        i, u, d = C.c_uint64(), C.c_uint64(), C.c_double()
        self._ck(self._L.gkd_pair(self._h, a, b, C.byref(i), C.byref(u), C.byref(d)))
        return i.value, u.value, d.value

    # -- misc --------------------------------------------------------------------------------------
    def reset(self):
        self._ck(self._L.gkd_reset(self._h))
        self._keep.clear()
        self._adopted.clear()

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of this context (wrap with torch.cuda.ExternalStream to time on it)"""
        return self._L.gkd_stream(self._h) or 0

    def save_sets(self, path: str):
        """write every set (with label and comment) to a .kset cache file"""
        self._ck(self._L.gkd_save_sets(self._h, path.encode()))

    def load_sets(self, path: str) -> Tuple[int, int]:
        """append the sets of a .kset cache file; returns (first id, count)"""
        first, n = C.c_uint32(), C.c_uint32()
        self._ck(self._L.gkd_load_sets(self._h, path.encode(), C.byref(first), C.byref(n)))
        return first.value, n.value

    def truncate(self, n_keep: int):
        """drop the sets with id >= n_keep (a streamed panel) and recycle their arena"""
        self._ck(self._L.gkd_truncate(self._h, n_keep))
        self._adopted = [x for x in self._adopted if x[0] < n_keep]

    def metrics(self) -> dict:
        m = GkdMetrics()
        self._ck(self._L.gkd_get_metrics(self._h, C.byref(m)))
        return {f: getattr(m, f) for f, _ in GkdMetrics._fields_ if not f.startswith("reserved")}


def format_double(v: float) -> str:
    buf = C.create_string_buffer(64)
    _lib.load().gkd_format_double(v, buf, 64)
    return buf.value.decode()


def synth(dst, seed: int, family: int, member: int, rate: float, protein: bool = False, device: int = 0):
    """Fill `dst` (numpy uint8 array or torch uint8 tensor, host or device) with a synthetic sequence."""
    L = _lib.load()
    if isinstance(dst, np.ndarray):
        ptr, n, dev = dst.ctypes.data, dst.size, -1
    else:
        ptr, n, dev = dst.data_ptr(), dst.numel(), (device if dst.is_cuda else -1)
    fn = L.gkd_synth_protein if protein else L.gkd_synth_dna
    rc = fn(dev, ptr, n, seed, family, member, float(rate))
    if rc:
        raise GkdError(rc, "synth failed")
    return dst


class Group:
    """Several devices in one process behind the C ABI (gkd_group_*): one context per listed device ordinal
    (an ordinal may repeat), one host thread per member, peer copies of set arenas, no reduction."""

    def __init__(self, devices: Sequence[int], k: int = 0, alphabet: int = DNA, strand_mode: int = STRAND_BOTH,
                 workspace_bytes: int = 0, panel: int = 0):
        self._L = _lib.load()
        cfg = GkdConfig(device=0, k=k, alphabet=alphabet, strand_mode=strand_mode, workspace_bytes=workspace_bytes)
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._L.gkd_group_create(C.byref(h), C.byref(cfg), devs, len(devices))
        if rc:
            raise GkdError(rc, (self._L.gkd_group_last_error(None) or b"").decode())
        self._h = h
        if panel:
            self._ck(self._L.gkd_group_set_panel(self._h, panel))

    def _ck(self, rc: int):
        if rc:
            raise GkdError(rc, (self._L.gkd_group_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.gkd_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __len__(self) -> int:
        return self._L.gkd_group_count(self._h)

    def add(self, member: int, contigs) -> int:
        """one genome on one member (members must be filled in ascending order); returns its global id"""
        if isinstance(contigs, (str, bytes, bytearray, np.ndarray)):
            contigs = [contigs]
        triples = [Engine._ptr_len(c) for c in contigs]
        n = len(triples)
        ptrs = (C.c_void_p * max(n, 1))(*[t[0] for t in triples])
        lens = (C.c_uint64 * max(n, 1))(*[t[1] for t in triples])
        out = C.c_uint32()
        self._ck(self._L.gkd_group_add_sequences(self._h, member, ptrs, lens, n, C.byref(out)))
        return out.value

    def add_fasta(self, path: str) -> int:
        n = C.c_uint32()
        self._ck(self._L.gkd_group_add_fasta_file(self._h, path.encode(), C.byref(n)))
        return n.value

    def label(self, i: int) -> str:
        return self._L.gkd_group_label(self._h, i).decode("latin-1")

    def comment(self, i: int) -> str:
        return self._L.gkd_group_comment(self._h, i).decode("latin-1")

    def build(self):
        self._ck(self._L.gkd_group_build(self._h))

    def all_vs_all(self):
        n = len(self)
        npairs = n * (n - 1) // 2
        inter, dist = np.empty(npairs, dtype=np.uint64), np.empty(npairs, dtype=np.float64)
        self._ck(self._L.gkd_group_all_vs_all(self._h, inter.ctypes.data, dist.ctypes.data))
        return inter, dist
