#!/usr/bin/env python
"""SURVEY 8(f) rank 4, measured: batched search over a device-resident index (sufr_b200_index_search, one GPU thread per
query; mirrors SufrFile::suffix_search / SufrSearch::search, sufr_search.rs:104-350).

    python tools/search_bench.py [--bases 1000000000] [--queries 4000000] [--len 24] [--cpu-queries 200000]

Builds the suffix array of a synthetic genome on the GPU, searches a batch of queries (half of them substrings of the
text, half random), CHECKS every answer on the device (the suffix at the first rank starts with the query, the
suffixes just outside the range do not, the range of an absent query is empty), and times the call with host buffers
(queries uploaded, rank ranges downloaded inside the timed region).  A vectorised numpy restatement of the same binary
search, one thread, is timed beside it on a slice of the batch."""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bench  # noqa: E402
import sufr_b200 as S  # noqa: E402
from sufr_b200 import _lib  # noqa: E402


def cpu_search(text, sa, q, L):
    """lower / upper bound of every query (rows of q, L bytes each) in the suffix array, by L-byte prefixes."""
    n, s = len(text), len(sa)
    pad = np.concatenate([text, np.zeros(L, dtype=np.uint8)])

    def key(positions):  # big-endian integer words of the L bytes at the positions (suffixes past the end: zero fill)
        w = pad[positions[:, None] + np.arange(L)[None, :]]
        return [w[:, i:i + 8].copy().view(">u8")[:, 0] if i + 8 <= L else None for i in range(0, L, 8)]

    qk = [q[:, i:i + 8].copy().view(">u8")[:, 0] for i in range(0, L, 8)]

    def less(ak, bk, or_equal):
        res = np.zeros(len(ak[0]), dtype=bool)
        eq = np.ones(len(ak[0]), dtype=bool)
        for a, b in zip(ak, bk):
            res |= eq & (a < b)
            eq &= a == b
        return res | eq if or_equal else res

    out = []
    for upper in (False, True):
        lo = np.zeros(len(q), dtype=np.int64)
        hi = np.full(len(q), s, dtype=np.int64)
        while True:
            act = lo < hi
            if not act.any():
                break
            mid = (lo + hi) // 2
            sk = key(sa[np.minimum(mid, s - 1)].astype(np.int64))
            go_right = less(sk, qk, upper) & act  # suffix < query (lower bound) or <= query (upper bound)
            lo = np.where(go_right, mid + 1, lo)
            hi = np.where(act & ~go_right, mid, hi)
        out.append(lo)
    return out[0], out[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=int, default=1_000_000_000)
    ap.add_argument("--queries", type=int, default=4_000_000)
    ap.add_argument("--len", type=int, default=24)
    ap.add_argument("--cpu-queries", type=int, default=200_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    L = a.len
    assert L % 8 == 0, "the numpy comparator packs queries into 8-byte words"
    text_len, starts = bench.record_layout(a.bases)
    ctx = S.Context(0)
    d_text = torch.empty(text_len, dtype=torch.uint8, device="cuda")
    st = np.asarray(starts, dtype=np.uint64)
    assert _lib.lib().sufr_b200_synth_dna(ctx.handle, d_text.data_ptr(), text_len, bench.SEED, st.ctypes.data, len(st), ord("%")) == 0
    bargs = S.SufrBuilderArgs(text=b"", is_dna=True, sequence_starts=starts, sequence_names=[f"chr{i + 1}" for i in range(len(starts))])
    r = S.build(bargs, index_bits=32, ctx=ctx, result_memory=S.MEM_DEVICE, device_text=(d_text.data_ptr(), text_len))
    idx = S.SufrIndex(r, bargs)
    # queries: substrings of the text that contain no delimiter, and random strings
    g = torch.Generator(device="cuda").manual_seed(7)
    half = a.queries // 2
    p = torch.randint(0, text_len - L - 1, (half,), device="cuda", generator=g)
    sub = d_text[p[:, None] + torch.arange(L, device="cuda")[None, :]]
    rnd = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (a.queries - half, L), device="cuda", generator=g)]
    q = torch.cat([sub, rnd])[torch.randperm(a.queries, device="cuda", generator=g)].contiguous()
    blob = q.cpu().numpy().reshape(-1)
    offs = (np.arange(a.queries + 1, dtype=np.uint64) * L)
    b = np.zeros(a.queries, dtype=np.uint64)
    e = np.zeros(a.queries, dtype=np.uint64)

    def call():
        rc = _lib.lib().sufr_b200_index_search(idx._h, blob.ctypes.data, offs.ctypes.data, a.queries, 0, 0, 0, b.ctypes.data, e.ctypes.data)
        assert rc == 0, _lib.lib().sufr_b200_last_error()

    call()
    times = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        call()
        times.append(time.perf_counter() - t0)
    dt = min(times)
    # ---- check every answer on the device
    none = np.uint64(0xFFFFFFFFFFFFFFFF)
    found = b != none
    sa = r.sa_tensor()
    s = r.num_suffixes
    tb = torch.from_numpy(np.where(found, b, 0).astype(np.int64)).cuda()
    te = torch.from_numpy(np.where(found, e, 0).astype(np.int64)).cuda()
    tf = torch.from_numpy(found).cuda()
    ar = torch.arange(L, device="cuda")[None, :]
    padded = torch.cat([d_text, torch.zeros(L, dtype=torch.uint8, device="cuda")])

    def prefix_at(rank):
        return padded[sa[rank.clamp(0, s - 1)].long()[:, None] + ar]

    first_ok = (prefix_at(tb) == q).all(dim=1)
    last_ok = (prefix_at(te - 1) == q).all(dim=1)
    before_ok = (tb == 0) | ~(prefix_at(tb - 1) == q).all(dim=1)
    after_ok = (te >= s) | ~(prefix_at(te) == q).all(dim=1)
    bad_found = int((tf & ~(first_ok & last_ok & before_ok & after_ok)).sum())
    # (absent queries are checked through the CPU comparator below)
    # ---- numpy restatement on a slice, timed and compared
    k = min(a.cpu_queries, a.queries)
    h_text = d_text.cpu().numpy()
    h_sa = sa.cpu().numpy()
    t0 = time.perf_counter()
    lo, hi = cpu_search(h_text, h_sa, q[:k].cpu().numpy(), L)
    cpu_dt = time.perf_counter() - t0
    gb = np.where(found[:k], b[:k], 0).astype(np.int64)
    ge = np.where(found[:k], e[:k], 0).astype(np.int64)
    agree = int(((hi - lo == ge - gb) & ((hi == lo) | (lo == gb))).sum())
    out = {"metric": "queries/s (suffix_search, rank ranges)", "value": a.queries / dt, "unit": "queries/s",
           "ms_per_batch": 1e3 * dt, "queries": a.queries, "query_len": L, "found": int(found.sum()),
           "index": {"text_len": text_len, "suffixes": int(s), "index_bits": 32},
           "h2d_bytes_per_batch": int(blob.nbytes + offs.nbytes), "d2h_bytes_per_batch": int(b.nbytes + e.nbytes),
           "check": {"ranges_checked_on_device": int(found.sum()), "bad_ranges": bad_found,
                     "compared_with_cpu_restatement": k, "agree": agree, "ok": bad_found == 0 and agree == k},
           "cpu_baseline": {"value": k / cpu_dt, "unit": "queries/s", "cores": 1, "kind": "port",
                            "sample": f"{k} queries of the batch, vectorised numpy binary search on {L}-byte prefixes, {cpu_dt:.2f} s"},
           "note": "timed region: queries uploaded from host memory, search kernel, rank ranges downloaded; best of %d" % a.reps}
    print(json.dumps(out))
    idx.close()
    r.free()


if __name__ == "__main__":
    main()
