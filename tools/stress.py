#!/usr/bin/env python
"""Randomised differential test: CUDA path (through the C ABI) vs the CPU oracle on many seeded inputs.
Complements the fixed cases of tests/test_gpu_parity.py; run on a GPU box: `python tools/stress.py --cases 300`."""
from __future__ import annotations

import argparse
import random
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle as O  # noqa: E402
import sufr_b200 as S  # noqa: E402
from sufr_b200.distributed import previous_last_suffix, shard_layout  # noqa: E402


def make_text(rng: random.Random, n: int):
    kind = rng.choice(["dna", "dna", "dna_rare", "dna_rare", "dna_n", "protein", "binary", "bytes", "repeat", "tandem"])
    np_rng = np.random.default_rng(rng.randrange(1 << 30))
    if kind == "protein":
        t = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)[np_rng.integers(0, 20, n)]
    elif kind == "binary":
        t = np.frombuffer(b"AB", np.uint8)[np_rng.integers(0, 2, n)]
    elif kind == "bytes":
        t = np_rng.integers(1, 255, n).astype(np.uint8)
        t[t == ord("$")] = ord("A")
    else:
        t = np.frombuffer(b"ACGT", np.uint8)[np_rng.integers(0, 4, n)].copy()
        if kind in ("dna_rare", "repeat", "tandem") or (kind == "dna_n" and rng.random() < 0.5):
            rare = np.frombuffer(b"N%RYnacgt#~" if kind != "dna_n" else b"N", np.uint8)
            k = max(1, int(n * rng.choice([0.0005, 0.01, 0.08])))
            t[np_rng.integers(0, n, k)] = rare[np_rng.integers(0, len(rare), k)]
        if kind == "dna_n" and n > 5000:
            for _ in range(rng.randrange(1, 4)):
                s = rng.randrange(0, n - 2500)
                t[s:s + rng.choice([300, 1000, 1001, 2200])] = ord("N")
        if kind == "repeat" and n > 2000:
            for _ in range(rng.randrange(1, 30)):
                ln = rng.randrange(20, min(5000, n // 3))
                a, b = rng.randrange(0, n - ln), rng.randrange(0, n - ln)
                t[b:b + ln] = t[a:a + ln]
        if kind == "tandem" and n > 2000:
            for _ in range(rng.randrange(1, 6)):
                unit = t[:rng.randrange(1, 40)].copy()
                copies = rng.randrange(5, max(6, min(3000, n // (2 * len(unit)))))
                s = rng.randrange(0, max(1, n - len(unit) * copies))
                seg = np.tile(unit, copies)[: n - s]
                t[s:s + len(seg)] = seg
            for _ in range(rng.randrange(0, 3)):
                s = rng.randrange(0, n - 200)
                t[s:s + rng.randrange(30, 200)] = ord(rng.choice("AT"))
    text = t.tobytes() + rng.choice([b"$", b"$", b"$", b""])
    return kind, text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-n", type=int, default=400_000)
    args = ap.parse_args()
    rng = random.Random(args.seed)
    failures = 0
    skipped = 0
    t0 = time.time()
    for case in range(args.cases):
        n = int(10 ** rng.uniform(1.0, np.log10(args.max_n)))
        kind, text = make_text(rng, n)
        is_dna = kind not in ("protein", "binary", "bytes") and rng.random() < 0.8
        flags = dict(is_dna=is_dna, allow_ambiguity=is_dna and rng.random() < 0.4,
                     ignore_softmask=rng.random() < 0.3)
        mode = rng.choice(["full", "full", "full", "mask", "mql"])
        if mode == "mask":
            flags["seed_mask"] = rng.choice(["101", "1101", "10111011", "1101101101", "111010010100110111"])
        bits = rng.choice([32, 32, 64])
        world = rng.choice([1, 1, 1, 2, 3, 5])
        if len(text) < 8:
            continue
        try:
            if mode == "mql":
                full = O.oracle_build(text, num_partitions=1, **flags)
                if full.num_suffixes == 0:
                    continue
                flags["max_query_len"] = int(full.lcp.max()) + rng.randrange(1, 5)  # no ties: well defined
            want = O.oracle_build(text, num_partitions=1 if len(text) < 4000 else 16, threads=4, index_bits=bits, **flags)
        except O.OracleError as e:
            continue
        if want.n_ranges and not flags.get("seed_mask"):
            # The reference is only a function of its input when every N-prefixed suffix lies in a recorded run
            # (SURVEY 8a rule 3) and no max-query-len cap interferes with the N-run shortcut: skip the rest.
            t = np.frombuffer(want.text, np.uint8)
            in_run = np.zeros(len(t), bool)
            for a, b in want.n_ranges:
                in_run[a:b] = True
            if flags.get("max_query_len") or bool(((t == ord("N")) & ~in_run).any()):
                skipped += 1
                continue
        bargs = S.SufrBuilderArgs(text=text, **flags)
        try:
            if world == 1:
                got = S.build(bargs, index_bits=bits)
                sa, lcp = got.sa.copy(), got.lcp.copy()
                ok = got.text == want.text and got.n_ranges == want.n_ranges
                got.free()
            else:
                shards = [S.build(bargs, index_bits=bits, rank=r, world_size=world) for r in range(world)]
                meta = [(s.num_suffixes, s.first_suffix, s.last_suffix) for s in shards]
                offs, total = shard_layout(meta)
                for r, s in enumerate(shards):
                    s.set_shard_layout(offs[r], total)
                    prev = previous_last_suffix(meta, r)
                    if prev is not None and s.num_suffixes:
                        s.patch_seam(prev)
                sa = np.concatenate([s.sa for s in shards])
                lcp = np.concatenate([s.lcp for s in shards])
                ok = True
                for s in shards:
                    s.free()
            ok = ok and np.array_equal(sa, want.sa) and np.array_equal(lcp, want.lcp)
        except Exception as e:  # noqa: BLE001
            ok = False
            print("EXC", repr(e))
        if not ok:
            failures += 1
            bad_sa = int(np.nonzero(sa != want.sa)[0][0]) if len(sa) == len(want.sa) and (sa != want.sa).any() else -1
            bad_lcp = int(np.nonzero(lcp != want.lcp)[0][0]) if len(lcp) == len(want.lcp) and (lcp != want.lcp).any() else -1
            print(f"FAIL case {case}: kind={kind} n={len(text)} flags={flags} bits={bits} world={world} "
                  f"first_sa_diff={bad_sa} first_lcp_diff={bad_lcp} sizes={len(sa)}/{len(want.sa)}", flush=True)
    print(f"{args.cases} cases, {failures} failures, {skipped} skipped as ill-defined for the reference, "
          f"{time.time() - t0:.1f}s")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
