"""Host-side mirror of the reference's `create` interface, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference:

* :class:`SufrBuilderArgs`  -- libsufr/src/types.rs:527-582
* :class:`SufrBuilder`      -- libsufr/src/sufr_builder.rs:38-220 (``SufrBuilder::<T>::new`` builds AND
  writes the file; the public fields of the struct are attributes here)
* :class:`SuffixArray`      -- ``SuffixArray::write`` u32/u64 dispatch, libsufr/src/suffix_array.rs:460-470
* :class:`SeedMask`         -- libsufr/src/types.rs:36-200
* :func:`read_sequence_file`, :func:`find_lcp_full_offset` -- libsufr/src/util.rs:51-89, :19-37

Everything that computes goes through ``libsufr_b200.so`` (CUDA, sm_100a).  Nothing here touches the
oracle and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST

OUTFILE_VERSION = 6          # types.rs:16
SENTINEL_CHARACTER = b"$"    # types.rs:20


class SufrError(RuntimeError):
    """anyhow::Error of the reference; ``code`` is the C ABI return code."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def _check(rc: int):
    if rc != _lib.OK:
        raise SufrError(rc, _lib.lib().sufr_b200_last_error().decode(errors="replace"))


# ------------------------------------------------------------------ types.rs:36-200
@dataclass
class SeedMask:
    mask: str
    bytes: List[int]
    positions: List[int]
    differences: List[int]
    weight: int

    @staticmethod
    def new(mask: str) -> "SeedMask":
        n = max(1, len(mask))
        b = (C.c_uint8 * n)()
        p = (C.c_uint64 * n)()
        d = (C.c_uint64 * n)()
        w = _lib.lib().sufr_b200_seed_mask(mask.encode(), b, p, d)
        if w < 0:
            raise SufrError(_lib.ERR_ARGUMENT, f"Invalid seed mask '{mask}'")  # types.rs:82
        return SeedMask(mask, list(b)[:len(mask)], list(p)[:w], list(d)[:w], int(w))

    @staticmethod
    def is_valid(mask: str) -> bool:
        return _lib.lib().sufr_b200_seed_mask(mask.encode(), None, None, None) >= 0

    def __str__(self):
        return self.mask


def find_lcp_full_offset(lcp: int, seed_mask: Optional[str]) -> int:
    """util.rs:19-37"""
    return int(_lib.lib().sufr_b200_find_lcp_full_offset(lcp, seed_mask.encode() if seed_mask else None))


# ------------------------------------------------------------------ types.rs:271-282, util.rs:51-89
@dataclass
class SequenceFileData:
    seq: bytes
    start_positions: List[int]
    sequence_names: List[str]


def read_sequence_file(path, sequence_delimiter: bytes = b"%") -> SequenceFileData:
    s = _lib.Sequences()
    _check(_lib.lib().sufr_b200_read_sequence_file(str(path).encode(), sequence_delimiter[0], C.byref(s)))
    try:
        seq = C.string_at(s.seq, s.seq_len)
        starts = [int(s.start_positions[i]) for i in range(s.num_sequences)]
        names = [s.sequence_names[i].decode() for i in range(s.num_sequences)]
    finally:
        _lib.lib().sufr_b200_sequences_free(C.byref(s))
    return SequenceFileData(seq, starts, names)


# ------------------------------------------------------------------ types.rs:527-582
@dataclass
class SufrBuilderArgs:
    text: bytes
    path: Optional[str] = None
    low_memory: bool = True
    max_query_len: Optional[int] = None
    is_dna: bool = False
    allow_ambiguity: bool = False
    ignore_softmask: bool = False
    sequence_starts: Sequence[int] = (0,)
    sequence_names: Sequence[str] = ("1",)
    num_partitions: int = 16
    seed_mask: Optional[str] = None
    random_seed: int = 42


class _CArgs:
    """Keeps the ctypes buffers behind a SufrB200Args alive."""

    def __init__(self, a: SufrBuilderArgs, *, text_ptr: Optional[int] = None, text_len: Optional[int] = None,
                 rank: int = 0, world_size: int = 1):
        self.c = _lib.Args()
        if text_ptr is None:
            self._text = np.frombuffer(a.text, dtype=np.uint8) if len(a.text) else np.zeros(1, np.uint8)
            self.c.text = self._text.ctypes.data
            self.c.text_len = len(a.text)
        else:
            self.c.text = text_ptr
            self.c.text_len = text_len
        self.c.path = a.path.encode() if a.path is not None else None
        self.c.low_memory = int(a.low_memory)
        self.c.has_max_query_len = int(a.max_query_len is not None)
        self.c.max_query_len = int(a.max_query_len or 0)
        self.c.is_dna = int(a.is_dna)
        self.c.allow_ambiguity = int(a.allow_ambiguity)
        self.c.ignore_softmask = int(a.ignore_softmask)
        self._starts = np.asarray(list(a.sequence_starts), dtype=np.uint64)
        self.c.sequence_starts = self._starts.ctypes.data if len(self._starts) else None
        if len(a.sequence_names) != len(a.sequence_starts):
            # the C ABI carries ONE count for both arrays (make_sufr_frame reads num_sequences entries of each)
            raise SufrError(_lib.ERR_ARGUMENT, f"sequence_names has {len(a.sequence_names)} entries but "
                                               f"sequence_starts has {len(a.sequence_starts)}")
        names = [s.encode() for s in a.sequence_names]
        self._names = (C.c_char_p * max(1, len(names)))(*names)
        self.c.sequence_names = self._names
        self.c.num_sequences = len(self._starts)
        self.c.num_partitions = a.num_partitions
        self.c.seed_mask = a.seed_mask.encode() if a.seed_mask is not None else None
        self.c.random_seed = a.random_seed
        self.c.rank = rank
        self.c.world_size = world_size


class Context:
    """One CUDA stream + device memory pool (SufrB200Ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(_lib.lib().sufr_b200_ctx_create(device, C.byref(self._h)))
        self.device = device

    def reserve(self, text_len: int, index_bits: int = 32):
        _check(_lib.lib().sufr_b200_ctx_reserve(self._h, text_len, index_bits))

    def trim(self):
        _lib.lib().sufr_b200_ctx_trim(self._h)

    def close(self):
        if self._h:
            _lib.lib().sufr_b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class BuildResult:
    """Owner of one SufrB200Result (host or device memory)."""

    def __init__(self, ctx: Context, cargs: _CArgs, res: "_lib.Result"):
        self._ctx, self._cargs, self.c = ctx, cargs, res

    # -- scalars
    @property
    def index_bits(self): return int(self.c.index_bits)
    @property
    def text_len(self): return int(self.c.text_len)
    @property
    def num_suffixes(self): return int(self.c.num_suffixes)
    @property
    def total_suffixes(self): return int(self.c.total_suffixes)
    @property
    def shard_offset(self): return int(self.c.shard_offset)
    @property
    def first_suffix(self): return int(self.c.first_suffix)
    @property
    def last_suffix(self): return int(self.c.last_suffix)
    @property
    def on_device(self): return self.c.memory == MEM_DEVICE
    @property
    def timings(self): return self.c.timings.as_dict()
    @property
    def kernel_launches(self): return int(self.c.kernel_launches)
    @property
    def h2d_bytes(self): return int(self.c.h2d_bytes)
    @property
    def d2h_bytes(self): return int(self.c.d2h_bytes)
    @property
    def n_ranges(self) -> List[Tuple[int, int]]:
        return [(int(self.c.n_ranges[2 * i]), int(self.c.n_ranges[2 * i + 1])) for i in range(self.c.num_n_ranges)]

    def _np(self, ptr, count, dtype):
        if self.on_device:
            raise SufrError(_lib.ERR_ARGUMENT, "result lives in device memory")
        if count == 0:
            return np.zeros(0, dtype)
        ct = {np.uint8: C.c_uint8, np.uint32: C.c_uint32, np.uint64: C.c_uint64}[dtype]
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), (count,))

    @property
    def dtype(self):
        return np.uint32 if self.c.index_bits == 32 else np.uint64

    @property
    def text(self) -> bytes:
        if not self.c.text:
            raise SufrError(_lib.ERR_ARGUMENT, "the transformed text is only returned on rank 0 of a sharded build")
        return self._np(self.c.text, self.text_len, np.uint8).tobytes()

    @property
    def sa(self) -> np.ndarray:
        return self._np(self.c.sa, self.num_suffixes, self.dtype)

    @property
    def lcp(self) -> np.ndarray:
        return self._np(self.c.lcp, self.num_suffixes, self.dtype)

    def _tensor(self, ptr, count, typestr):
        """Zero-copy torch view of a device buffer (valid until free())."""
        import torch

        class _Dev:
            pass
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (max(count, 0),), "typestr": typestr, "data": (int(ptr), False),
                                      "version": 2}
        if count == 0:
            return torch.zeros(0, dtype=torch.int64, device=f"cuda:{self._ctx.device}")
        return torch.as_tensor(d, device=f"cuda:{self._ctx.device}")

    def sa_tensor(self):
        """Device result: the suffix array as an int32 / int64 torch tensor (same bits as u32 / u64)."""
        return self._tensor(self.c.sa, self.num_suffixes, "<i4" if self.c.index_bits == 32 else "<i8")

    def lcp_tensor(self):
        return self._tensor(self.c.lcp, self.num_suffixes, "<i4" if self.c.index_bits == 32 else "<i8")

    def text_tensor(self):
        return self._tensor(self.c.text, self.text_len, "|u1")

    def device_pointers(self):
        """(text, sa, lcp) raw device addresses of a device result."""
        return int(self.c.text or 0), int(self.c.sa or 0), int(self.c.lcp or 0)

    def set_shard_layout(self, shard_offset: int, total_suffixes: int):
        """Sharded builds: position of this shard in the whole suffix array (from the ranks' counts)."""
        self.c.shard_offset = shard_offset
        self.c.total_suffixes = total_suffixes

    def patch_seam(self, prev_last_suffix: int):
        _check(_lib.lib().sufr_b200_patch_seam(self._ctx.handle, C.byref(self._cargs.c), C.byref(self.c),
                                               prev_last_suffix))

    def verify(self, prev_last_suffix: Optional[int] = None) -> dict:
        """Full on-device check of a DEVICE result (sufr_b200_verify): every SA entry indexed and unique, every
        adjacent pair in order, every LCP value exact.  Returns the report with ``ok`` added; for a sharded
        build ``ok`` covers this shard only (the caller sums num_suffixes against expected_suffixes)."""
        rep = _lib.VerifyReport()
        _check(_lib.lib().sufr_b200_verify(self._ctx.handle, C.byref(self._cargs.c), C.byref(self.c),
                                           int(prev_last_suffix is not None), int(prev_last_suffix or 0), C.byref(rep)))
        d = rep.as_dict()
        errors = d["order_errors"] + d["lcp_errors"] + d["out_of_range"] + d["not_indexed"] + d["duplicates"]
        d["ok"] = errors == 0 and (self._cargs.c.world_size > 1 or d["expected_suffixes"] == self.num_suffixes)
        return d

    def write(self):
        _check(_lib.lib().sufr_b200_write(C.byref(self._cargs.c), C.byref(self.c)))

    def free(self):
        if self.c.owner:
            if not self._ctx.handle:  # context already destroyed (interpreter shutdown): nothing left to return to
                self.c.owner = None
                return
            _lib.lib().sufr_b200_result_free(self._ctx.handle, C.byref(self.c))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def build(args: SufrBuilderArgs, *, index_bits: int = 0, ctx: Optional[Context] = None, device: int = 0,
          result_memory: int = MEM_HOST, device_text: Optional[Tuple[int, int]] = None,
          rank: int = 0, world_size: int = 1) -> BuildResult:
    """SufrBuilder::new up to (not including) write(): sufr_b200_build.

    ``device_text=(ptr, len)`` passes a text that already lives in device memory.
    """
    ctx = ctx or default_context(device)
    if device_text is not None:
        cargs = _CArgs(args, text_ptr=device_text[0], text_len=device_text[1], rank=rank, world_size=world_size)
        text_memory = MEM_DEVICE
    else:
        cargs = _CArgs(args, rank=rank, world_size=world_size)
        text_memory = MEM_HOST
    res = _lib.Result()
    _check(_lib.lib().sufr_b200_build(ctx.handle, C.byref(cargs.c), index_bits, text_memory, result_memory,
                                      C.byref(res)))
    return BuildResult(ctx, cargs, res)


class SufrIndex:
    """Device-resident index over a DEVICE build result: the read side that consumes the LCP array
    (``SufrFile::subsample_suffix_array`` sufr_file.rs:429-456, ``suffix_search`` :777-836, ``locate`` :1132-1169,
    ``SufrSearch`` sufr_search.rs:104-350).  ``count`` / ``locate`` take a batch of queries; one GPU thread each."""

    def __init__(self, result: "BuildResult", args: SufrBuilderArgs):
        self._res, self._args = result, args
        self._h = C.c_void_p()
        _check(_lib.lib().sufr_b200_index_create(result._ctx.handle, C.byref(result._cargs.c), C.byref(result.c),
                                                 C.byref(self._h)))
        self._sub_mql = None
        if args.seed_mask is not None:
            self._built = SeedMask.new(args.seed_mask).weight        # sufr_file.rs:528
        else:
            self._built = args.max_query_len or result.text_len       # sufr_file.rs:521-527

    def close(self):
        if self._h:
            _lib.lib().sufr_b200_index_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def subsample(self, max_query_len: int) -> int:
        kept = C.c_uint64()
        _check(_lib.lib().sufr_b200_index_subsample(self._h, max_query_len, C.byref(kept)))
        self._sub_mql = max_query_len
        return int(kept.value)

    def search(self, queries: Sequence[str], max_query_len: Optional[int] = None, low_memory: bool = False):
        """SufrFile::suffix_search: per query the half-open rank range in the full suffix array, or None.
        low_memory=False mirrors set_suffix_array_mem (sufr_file.rs:514-560): when the effective max_query_len is
        shorter than what the index was built with, the LCP-subsampled array is searched."""
        qs = [q.encode() if isinstance(q, str) else bytes(q) for q in queries]
        offs = np.zeros(len(qs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(q) for q in qs])
        blob = np.frombuffer(b"".join(qs) or b"\0", dtype=np.uint8)
        use_sub = 0
        if not low_memory:
            eff = min(max_query_len, self._built) if max_query_len else self._built
            if eff != self._built:
                if self._sub_mql != eff:
                    self.subsample(eff)
                use_sub = 1
        b = np.zeros(max(1, len(qs)), dtype=np.uint64)
        e = np.zeros(max(1, len(qs)), dtype=np.uint64)
        _check(_lib.lib().sufr_b200_index_search(self._h, blob.ctypes.data, offs.ctypes.data, len(qs),
                                                 int(max_query_len is not None), int(max_query_len or 0), use_sub,
                                                 b.ctypes.data, e.ctypes.data))
        none = np.uint64(0xFFFFFFFFFFFFFFFF)
        return [None if b[i] == none else (int(b[i]), int(e[i])) for i in range(len(qs))]

    def count(self, queries, max_query_len=None, low_memory=False) -> List[int]:
        return [0 if r is None else r[1] - r[0] for r in self.search(queries, max_query_len, low_memory)]

    def suffixes(self, rank_begin: int, count: int) -> np.ndarray:
        out = np.zeros(max(1, count), dtype=np.uint64)
        _check(_lib.lib().sufr_b200_index_suffixes(self._h, rank_begin, count, out.ctypes.data))
        return out[:count]

    def locate(self, queries, max_query_len=None, low_memory=False):
        """SufrFile::locate: per query a list of (rank, suffix, sequence_name, sequence_position) in rank order."""
        import bisect
        starts, names = list(self._args.sequence_starts), list(self._args.sequence_names)
        out = []
        for r in self.search(queries, max_query_len, low_memory):
            hits = []
            if r is not None:
                for k, suf in enumerate(self.suffixes(r[0], r[1] - r[0]).tolist()):
                    i = bisect.bisect_right(starts, suf) - 1
                    hits.append((r[0] + k, suf, names[i], suf - starts[i]))
            out.append(hits)
        return out


def create_multi(args: SufrBuilderArgs, devices: Sequence[int], index_bits: int = 0) -> dict:
    """``sufr::create`` on several GPUs of one box in one call (sufr_b200_create_multi): builds and writes
    ``args.path``; returns the counts and timings of the whole build."""
    cargs = _CArgs(args)
    res = _lib.Result()
    devs = (C.c_int * len(devices))(*[int(d) for d in devices])
    _check(_lib.lib().sufr_b200_create_multi(C.byref(cargs.c), devs, len(devices), index_bits, C.byref(res)))
    try:
        return {"num_suffixes": int(res.num_suffixes), "text_len": int(res.text_len), "index_bits": int(res.index_bits),
                "timings": res.timings.as_dict(), "kernel_launches": int(res.kernel_launches),
                # (ctypes.string_at takes a C int: texts of 2 GiB and more go through numpy)
                "text": np.ctypeslib.as_array(C.cast(res.text, C.POINTER(C.c_uint8)), (int(res.text_len),)).tobytes()
                if res.text and res.text_len else b"",
                "n_ranges": [(int(res.n_ranges[2 * i]), int(res.n_ranges[2 * i + 1])) for i in range(res.num_n_ranges)]}
    finally:
        _lib.lib().sufr_b200_result_free(None, C.byref(res))


class SufrBuilder:
    """``SufrBuilder::<T>::new(args)``: builds the suffix and LCP arrays on the GPU and writes the
    `.sufr` file at ``args.path`` (default "out.sufr", sufr_builder.rs:215).  ``index_bits`` plays the
    role of the type parameter T (32 -> u32, 64 -> u64)."""

    def __init__(self, args: SufrBuilderArgs, index_bits: int = 32, *, ctx: Optional[Context] = None,
                 device: int = 0, write: bool = True):
        if index_bits not in (32, 64):
            raise SufrError(_lib.ERR_ARGUMENT, "index_bits must be 32 or 64")
        self._res = build(args, index_bits=index_bits, ctx=ctx, device=device)
        r = self._res
        self.version = OUTFILE_VERSION
        self.is_dna = args.is_dna
        self.allow_ambiguity = args.allow_ambiguity
        self.ignore_softmask = args.ignore_softmask
        self.text_len = r.text_len
        self.num_suffixes = r.num_suffixes
        self.num_sequences = len(args.sequence_starts)
        self.sequence_starts = list(args.sequence_starts)
        self.sequence_names = list(args.sequence_names)
        self.text = r.text
        self.sort_type = ("Mask", SeedMask.new(args.seed_mask)) if args.seed_mask is not None else \
            ("MaxQueryLen", args.max_query_len or 0)
        self.n_ranges = r.n_ranges
        self.path = args.path if args.path is not None else "out.sufr"
        self.index_bits = index_bits
        # not part of the reference struct (its arrays live in temp files): kept for library users
        self.suffix_array = r.sa
        self.lcp_array = r.lcp
        self.timings = r.timings
        if write:
            r.write()


class SuffixArray:
    """The write half of libsufr::suffix_array::SuffixArray."""

    @staticmethod
    def write(args: SufrBuilderArgs, **kw) -> str:
        # suffix_array.rs:460-470: u32 iff text.len() < u32::MAX
        bits = 32 if len(args.text) < 0xFFFFFFFF else 64
        return SufrBuilder(args, bits, **kw).path
