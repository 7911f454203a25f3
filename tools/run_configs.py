#!/usr/bin/env python
"""Runs the BASELINE.json configs on one GPU at a chosen scale, verifies them with the size-independent
checks of tools/verify.py and prints one JSON line per config (kept under profiles/)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import workloads  # noqa: E402
from verify import check_pairs, check_positions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="config1,config3,config4,config5,config2b")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the BASELINE size")
    ap.add_argument("--pairs", type=int, default=3000)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    import sufr_b200 as S

    full = {"config1": 10_000_000, "config2a": 3_100_000_000, "config2b": 3_100_000_000, "config3": 1_000_000_000,
            "config4": 500_000_000, "config5": 1_000_000_000}
    ctx = S.Context(0)
    for name in args.configs.split(","):
        size = max(1000, int(full[name] * args.scale))
        t0 = time.time()
        w = workloads.ALL[name](size)
        gen_s = time.time() - t0
        bargs = S.SufrBuilderArgs(text=w.text, sequence_starts=w.sequence_starts, sequence_names=w.sequence_names,
                                  **w.flags)
        best = None
        res = None
        for _ in range(args.repeat):
            if res is not None:
                res.free()
            t0 = time.time()
            res = S.build(bargs, index_bits=w.index_bits, ctx=ctx)
            wall = time.time() - t0
            if best is None or res.timings["total_ms"] < best["total_ms"]:
                best = dict(res.timings)
                best["wall_s"] = wall
        sa, lcp, text = res.sa, res.lcp, res.text
        rng = np.random.default_rng(11)
        ranks = rng.integers(0, max(1, res.num_suffixes), args.pairs)
        # also probe the deepest LCPs: they exercise the doubling / PLCP path
        if res.num_suffixes > 10:
            deep = np.argsort(lcp[:: max(1, res.num_suffixes // 2_000_000)])[-50:] * max(1, res.num_suffixes // 2_000_000)
            ranks = np.concatenate([ranks, deep])
        t0 = time.time()
        bad = check_pairs(text, sa, lcp, ranks, seed_mask=w.flags.get("seed_mask"),
                          max_query_len=w.flags.get("max_query_len"), n_ranges=res.n_ranges)
        pos_ok = check_positions(text, sa, is_dna=w.flags.get("is_dna", False),
                                 allow_ambiguity=w.flags.get("allow_ambiguity", False))
        line = {
            "config": name, "text_len": res.text_len, "num_suffixes": res.num_suffixes, "index_bits": w.index_bits,
            "flags": {k: v for k, v in w.flags.items()}, "device_ms": best["total_ms"],
            "suffixes_per_s": res.num_suffixes / (best["total_ms"] * 1e-3), "phases_ms": best,
            "refine_rounds": int(res.c.refine_rounds), "doubling_rounds": int(res.c.doubling_rounds),
            "kernel_launches": res.kernel_launches, "max_lcp": int(lcp.max()) if res.num_suffixes else 0,
            "n_ranges": len(res.n_ranges), "pairs_checked": int(len(ranks)), "pair_mismatches": len(bad),
            "positions_ok": pos_ok, "first_mismatches": [tuple(map(str, b)) for b in bad[:3]],
            "gen_s": round(gen_s, 1), "verify_s": round(time.time() - t0, 1),
        }
        if name == "config3":
            line["precondition_no_lcp_ge_q"] = bool(int(lcp.max()) < 32)
        print(json.dumps(line), flush=True)
        res.free()
        ctx.trim()


if __name__ == "__main__":
    main()
