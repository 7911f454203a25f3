#!/usr/bin/env python
"""Smoke-size builds of every mode, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Each build is also compared with the CPU oracle, so a sanitizer-clean run is a correct run."""
import os
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle as O  # noqa: E402
import sufr_b200 as S  # noqa: E402


def rand_text(seed, n, alphabet, repeat_p=0.0):
    rng = random.Random(seed)
    out = bytearray()
    while len(out) < n:
        if out and rng.random() < repeat_p:
            s = rng.randrange(len(out))
            out += out[s:s + rng.randrange(1, 80)]
        else:
            out.append(rng.choice(alphabet))
    return bytes(out[:n]) + b"$"


def tandem(seed, n):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    parts, size, k = [], 0, 0
    while size < n:
        k += 1
        if k % 5 == 0:
            seg = np.full(int(rng.integers(1000, 1500)), ord("N"), dtype=np.uint8)
        elif k % 2:
            seg = np.tile(acgt[rng.integers(0, 4, int(rng.integers(1, 30)))], int(rng.integers(20, 400)))
        else:
            seg = acgt[rng.integers(0, 4, int(rng.integers(200, 2000)))]
        parts.append(seg)
        size += len(seg)
    return np.concatenate(parts)[:n].tobytes() + b"A$"


def family(seed, n, seg_len=160, copies=12):
    """one segment copied `copies` times with a substitution or two each, in random DNA (3 % of the text)"""
    rng = random.Random(seed)
    seg = bytes(rng.choice(b"ACGT") for _ in range(seg_len))
    per = (n - copies * seg_len) // copies
    out = bytearray()
    for _ in range(copies):
        out += bytes(rng.choice(b"ACGT") for _ in range(per))
        c = bytearray(seg)
        c[rng.randrange(seg_len // 2, seg_len)] = rng.choice(b"ACGT")
        out += c
    return bytes(out) + b"$"


def main():
    n = int(os.environ.get("SANITIZE_N", "40000"))
    cases = [
        ("fast2 dna u32", rand_text(1, n, b"ACGT"), dict(is_dna=True), 32, {}),
        ("fast2 dna u64 + N/softmask", rand_text(2, n, b"ACGTNacgt%", 0.02), dict(is_dna=True), 64, {}),
        ("3-bit general path", rand_text(3, n, b"ACGTN%", 0.02), dict(is_dna=True), 32, {"SUFR_B200_DEBUG_NO_FAST2": "1"}),
        ("protein", rand_text(4, n, b"ACDEFGHIKLMNPQRSTVWY%", 0.02), dict(), 32, {}),
        ("seed mask", rand_text(5, n, b"ACGT"), dict(is_dna=True, seed_mask="1101101101"), 32, {}),
        ("max-query-len", rand_text(6, n, b"ACDEFGHIKLMNPQRSTVWY"), dict(max_query_len=40), 32, {}),
        ("deep repeats: prefix doubling, N-run rule, softmask", tandem(7, n),
         dict(is_dna=True, allow_ambiguity=True, ignore_softmask=True), 32, {}),
        ("deep repeats, filtered after the sort", tandem(8, n), dict(is_dna=True), 64, {}),
        ("dense round 0", rand_text(9, n, b"ACGT", 0.05), dict(is_dna=True), 32, {"SUFR_B200_DEBUG_SPARSE_CAP": "16"}),
        ("deep repeats, inverse suffix array by sorting", tandem(10, n), dict(is_dna=True, allow_ambiguity=True), 32,
         {"SUFR_B200_DEBUG_SORT_ISA": "1"}),
        ("repeat family: shards finish by direct comparison (one block per group)", family(11, 2 * n), dict(is_dna=True), 32, {}),
    ]
    for name, text, kw, bits, env in cases:
        for k, v in env.items():
            os.environ[k] = v
        want = O.oracle_build(text, num_partitions=8, threads=4, index_bits=bits, **kw)
        for world in (1, 3):
            shards = [S.build(S.SufrBuilderArgs(text=text, **kw), index_bits=bits, rank=r, world_size=world) for r in range(world)]
            sa = np.concatenate([s.sa for s in shards])
            lcp = np.concatenate([s.lcp for s in shards])
            ok = np.array_equal(sa, want.sa) and (world > 1 or np.array_equal(lcp, want.lcp))
            print(f"{name:55s} world {world}: {len(sa)} suffixes, {'bit-exact' if ok else 'MISMATCH'}", flush=True)
            assert ok, name
            for s in shards:
                s.free()
        # device result + verifier
        import torch
        t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
        r = S.build(S.SufrBuilderArgs(text=b"", **kw), index_bits=bits, result_memory=S.MEM_DEVICE,
                    device_text=(t.data_ptr(), t.numel()))
        rep = r.verify()
        assert rep["ok"] or "max_query_len" in kw, (name, rep)
        # the read side: LCP-subsampled array + batched search, against a scan of the text
        bargs = S.SufrBuilderArgs(text=b"", **kw)
        idx = S.SufrIndex(r, bargs)
        rng = random.Random(len(name))
        tt = r.text_tensor().cpu().numpy().tobytes()
        qs = [tt[p:p + rng.randrange(1, 12)].decode("latin1") for p in (rng.randrange(len(tt) - 12) for _ in range(300))]
        for low in (True, False):
            got = idx.search(qs, max_query_len=6 if "seed_mask" not in kw else None, low_memory=low)
            assert len(got) == len(qs)
        idx.close()
        r.free()
        for k in env:
            os.environ.pop(k, None)
    print("all modes done")


if __name__ == "__main__":
    main()
