"""ctypes binding of the CPU oracle (oracle/libsufr_oracle.so).

Test infrastructure only: the oracle is the *checker* for the CUDA product.  Nothing under
``sufr_b200/`` imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "libsufr_oracle.so"


class OracleArgs(C.Structure):
    _fields_ = [
        ("text", C.c_void_p),
        ("text_len", C.c_uint64),
        ("is_dna", C.c_int32),
        ("allow_ambiguity", C.c_int32),
        ("ignore_softmask", C.c_int32),
        ("has_max_query_len", C.c_int32),
        ("max_query_len", C.c_uint64),
        ("seed_mask", C.c_char_p),
        ("num_partitions", C.c_uint64),
        ("random_seed", C.c_uint64),
        ("threads", C.c_int32),
        ("index_bits", C.c_int32),
        ("sequence_starts", C.c_void_p),
        ("num_sequences", C.c_uint64),
        ("sequence_names", C.POINTER(C.c_char_p)),
    ]


_lib = None


def build_oracle(force: bool = False) -> Path:
    src = ORACLE_DIR / "sufr_oracle.cpp"
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(ORACLE_DIR), "-s"])
    return ORACLE_SO


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.oracle_build.restype = C.c_void_p
        L.oracle_build.argtypes = [C.POINTER(OracleArgs), C.c_int, C.c_char_p, C.c_size_t]
        L.oracle_free.argtypes = [C.c_void_p]
        for name in ("oracle_num_suffixes", "oracle_text_len", "oracle_num_partitions_built",
                     "oracle_num_n_ranges", "oracle_file_size"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("oracle_text", "oracle_sa", "oracle_lcp", "oracle_n_ranges", "oracle_file_bytes"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_index_bits.restype = C.c_int
        L.oracle_index_bits.argtypes = [C.c_void_p]
        L.oracle_phase_times.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_find_lcp.restype = C.c_uint64
        L.oracle_find_lcp.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
        L.oracle_is_less.restype = C.c_int
        L.oracle_is_less.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.oracle_upper_bound.restype = C.c_uint64
        L.oracle_upper_bound.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.oracle_seed_mask_valid.restype = C.c_int
        L.oracle_seed_mask_valid.argtypes = [C.c_char_p]
        L.oracle_seed_mask.restype = C.c_int64
        L.oracle_seed_mask.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.oracle_find_lcp_full_offset.restype = C.c_uint64
        L.oracle_find_lcp_full_offset.argtypes = [C.c_uint64, C.c_char_p]
        L.oracle_read_sequence_file.restype = C.c_void_p
        L.oracle_read_sequence_file.argtypes = [C.c_char_p, C.c_uint8, C.c_char_p, C.c_size_t]
        L.oracle_seq_free.argtypes = [C.c_void_p]
        L.oracle_seq_len.restype = C.c_uint64
        L.oracle_seq_len.argtypes = [C.c_void_p]
        L.oracle_seq_bytes.restype = C.c_void_p
        L.oracle_seq_bytes.argtypes = [C.c_void_p]
        L.oracle_seq_count.restype = C.c_uint64
        L.oracle_seq_count.argtypes = [C.c_void_p]
        L.oracle_seq_starts.restype = C.c_void_p
        L.oracle_seq_starts.argtypes = [C.c_void_p]
        L.oracle_seq_name.restype = C.c_char_p
        L.oracle_seq_name.argtypes = [C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


@dataclass
class SeqData:
    """Mirror of libsufr::types::SequenceFileData (types.rs:271-282)."""
    seq: bytes
    start_positions: List[int]
    sequence_names: List[str]


def read_sequence_file(path, sequence_delimiter: bytes = b"%") -> SeqData:
    L = lib()
    err = C.create_string_buffer(512)
    h = L.oracle_read_sequence_file(str(path).encode(), sequence_delimiter[0], err, 512)
    if not h:
        raise OracleError(err.value.decode())
    try:
        n = L.oracle_seq_len(h)
        seq = C.string_at(L.oracle_seq_bytes(h), n)
        k = L.oracle_seq_count(h)
        starts = list(np.ctypeslib.as_array(C.cast(L.oracle_seq_starts(h), C.POINTER(C.c_uint64)), (k,)).copy()) if k else []
        names = [L.oracle_seq_name(h, i).decode() for i in range(k)]
    finally:
        L.oracle_seq_free(h)
    return SeqData(seq, [int(s) for s in starts], names)


@dataclass
class OracleResult:
    index_bits: int
    text: bytes
    sa: np.ndarray
    lcp: np.ndarray
    num_suffixes: int
    n_ranges: List[tuple]
    file_bytes: bytes
    phase_times: dict
    num_partitions_built: int


class Oracle:
    """Handle on a built oracle SufrBuilder; exposes the private methods the reference unit-tests."""

    def __init__(self, text: bytes, *, is_dna=False, allow_ambiguity=False, ignore_softmask=False,
                 max_query_len: Optional[int] = None, seed_mask: Optional[str] = None,
                 num_partitions: int = 16, random_seed: int = 42, threads: int = 1,
                 index_bits: int = 0, sequence_starts=(0,), sequence_names=("1",), do_sort=True):
        L = lib()
        self._L = L
        self._text = bytes(text)
        self._buf = (C.c_uint8 * max(1, len(self._text))).from_buffer_copy(self._text or b"\0")
        starts = np.asarray(list(sequence_starts), dtype=np.uint64)
        names = [s.encode() for s in sequence_names]
        name_arr = (C.c_char_p * max(1, len(names)))(*names) if names else (C.c_char_p * 1)()
        a = OracleArgs()
        a.text = C.addressof(self._buf)
        a.text_len = len(self._text)
        a.is_dna, a.allow_ambiguity, a.ignore_softmask = int(is_dna), int(allow_ambiguity), int(ignore_softmask)
        a.has_max_query_len = int(max_query_len is not None)
        a.max_query_len = int(max_query_len or 0)
        a.seed_mask = seed_mask.encode() if seed_mask is not None else None
        a.num_partitions = num_partitions
        a.random_seed = random_seed
        a.threads = threads
        a.index_bits = index_bits
        a.sequence_starts = starts.ctypes.data if len(starts) else None
        a.num_sequences = len(starts)
        a.sequence_names = name_arr
        err = C.create_string_buffer(512)
        self._h = L.oracle_build(C.byref(a), int(do_sort), err, 512)
        if not self._h:
            raise OracleError(err.value.decode())
        self._sorted = bool(do_sort)

    def close(self):
        if getattr(self, "_h", None):
            self._L.oracle_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- private-method mirrors (sufr_builder.rs unit tests)
    def find_lcp(self, s1, s2, length, skip=0):
        return int(self._L.oracle_find_lcp(self._h, s1, s2, length, skip))

    def is_less(self, s1, s2):
        return bool(self._L.oracle_is_less(self._h, s1, s2))

    def upper_bound(self, suffix, pivots):
        pv = np.asarray(list(pivots), dtype=np.uint64)
        return int(self._L.oracle_upper_bound(self._h, suffix, pv.ctypes.data if len(pv) else None, len(pv)))

    def result(self) -> OracleResult:
        L, h = self._L, self._h
        bits = L.oracle_index_bits(h)
        dt = np.uint32 if bits == 32 else np.uint64
        n = L.oracle_text_len(h)
        s = L.oracle_num_suffixes(h)
        text = C.string_at(L.oracle_text(h), n)
        if s:
            ct = C.c_uint32 if bits == 32 else C.c_uint64
            sa = np.ctypeslib.as_array(C.cast(L.oracle_sa(h), C.POINTER(ct)), (s,)).copy().astype(dt)
            lcp = np.ctypeslib.as_array(C.cast(L.oracle_lcp(h), C.POINTER(ct)), (s,)).copy().astype(dt)
        else:
            sa = np.zeros(0, dt)
            lcp = np.zeros(0, dt)
        k = L.oracle_num_n_ranges(h)
        nr = []
        if k:
            flat = np.ctypeslib.as_array(C.cast(L.oracle_n_ranges(h), C.POINTER(C.c_uint64)), (2 * k,))
            nr = [(int(flat[2 * i]), int(flat[2 * i + 1])) for i in range(k)]
        # the serialized file only when it is small enough to be useful as a bytes object (ctypes sizes are C ints)
        fsize = L.oracle_file_size(h) if self._sorted else 0
        fb = C.string_at(L.oracle_file_bytes(h), fsize) if 0 < fsize < (1 << 30) else b""
        t = (C.c_double * 6)()
        L.oracle_phase_times(h, t)
        keys = ["transform_s", "nscan_s", "pivots_s", "partition_s", "sort_s", "stitch_s"]
        return OracleResult(bits, text, sa, lcp, int(s), nr, fb, dict(zip(keys, list(t))),
                            int(L.oracle_num_partitions_built(h)))


def oracle_build(text: bytes, **kw) -> OracleResult:
    o = Oracle(text, **kw)
    try:
        return o.result()
    finally:
        o.close()


def seed_mask(mask: str):
    """Returns (bytes, positions, differences, weight) or None if invalid (types.rs:80-97)."""
    L = lib()
    cap = max(1, len(mask))
    pos = np.zeros(cap, np.uint64)
    dif = np.zeros(cap, np.uint64)
    byt = np.zeros(cap, np.uint8)
    w = L.oracle_seed_mask(mask.encode(), pos.ctypes.data, dif.ctypes.data, byt.ctypes.data, cap)
    if w < 0:
        return None
    return list(map(int, byt[:len(mask)])), list(map(int, pos[:w])), list(map(int, dif[:w])), int(w)


def seed_mask_valid(mask: str) -> bool:
    return bool(lib().oracle_seed_mask_valid(mask.encode()))


def find_lcp_full_offset(lcp: int, mask: Optional[str]) -> int:
    return int(lib().oracle_find_lcp_full_offset(lcp, mask.encode() if mask else None))


# ------------------------------------------------------------------ independent "spec" sorter
def spec_build(text: bytes, *, is_dna=False, allow_ambiguity=False, ignore_softmask=False,
               max_query_len=None, seed_mask_str=None):
    """Closed-form semantics of SURVEY.md section 8a ("Semantics distilled"), in pure Python.

    Independent of the literal restatement: used to cross-check it on small random inputs for the
    modes whose result is a pure function of the input (full sort, seed mask; MQL only without ties).
    Returns (transformed_text, sa_list, lcp_list).
    """
    t = bytearray(text)
    for i, b in enumerate(t):
        if 97 <= b <= 122:
            t[i] = ord("N") if ignore_softmask else (b & 0x5F)
    t = bytes(t)
    n = len(t)
    idx = [i for i, b in enumerate(t)
           if b == ord("$") or not is_dna or b in b"ACGT" or allow_ambiguity]
    if seed_mask_str is not None:
        positions = [i for i, c in enumerate(seed_mask_str) if c == "1"]

        def key(p):
            return bytes(t[p + o] for o in positions if p + o < n)
        order = sorted(idx, key=lambda p: (key(p), -p))

        def common(a, b):
            ka, kb = key(a), key(b)
            c = 0
            while c < len(ka) and c < len(kb) and ka[c] == kb[c]:
                c += 1
            return c
    else:
        q = max_query_len or 0

        def key(p):
            return t[p:p + q] if q else t[p:]
        order = sorted(idx, key=lambda p: (key(p), -p))

        def common(a, b):
            c = 0
            lim = min(n - a, n - b)
            if q:
                lim = min(lim, q)
            while c < lim and t[a + c] == t[b + c]:
                c += 1
            return c
    lcp = [0] * len(order)
    for j in range(1, len(order)):
        lcp[j] = common(order[j - 1], order[j])
    return t, order, lcp
