#!/usr/bin/env python
"""The north-star run: ONE call builds the suffix / LCP arrays of the synthetic 3.1 Gbp genome (u64 indices) on the
GPUs of one box and writes one `.sufr` file; the file is then read back and fully verified on a GPU.

    python tools/north_star.py [--gpus N] [--bases 3100000000] [--index-bits 64] [--out /dev/shm/north_star.sufr]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
import sufr_b200 as S  # noqa: E402
from sufr_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--bases", type=int, default=3_100_000_000)
    ap.add_argument("--index-bits", type=int, default=64)
    ap.add_argument("--out", default="/dev/shm/north_star.sufr")
    ap.add_argument("--runs", type=int, default=2)
    a = ap.parse_args()
    text_len, starts = bench.record_layout(a.bases)
    names = [f"chr{i + 1}" for i in range(len(starts))]
    ctx = S.Context(0)
    d_text = torch.empty(text_len, dtype=torch.uint8, device="cuda:0")
    st = np.asarray(starts, dtype=np.uint64)
    assert _lib.lib().sufr_b200_synth_dna(ctx.handle, d_text.data_ptr(), text_len, bench.SEED, st.ctypes.data, len(st), ord("%")) == 0
    h_text = d_text.cpu().numpy()
    del d_text
    torch.cuda.empty_cache()
    args = S.SufrBuilderArgs(text=memoryview(h_text), path=a.out, is_dna=True, sequence_starts=starts, sequence_names=names)
    runs = []
    for _ in range(a.runs):
        if os.path.exists(a.out):
            os.unlink(a.out)
        t0 = time.perf_counter()
        info = S.create_multi(args, list(range(a.gpus)), index_bits=a.index_bits)
        wall = time.perf_counter() - t0
        runs.append({"wall_s": round(wall, 3), "device_ms_slowest_shard": info["timings"]["total_ms"],
                     "h2d_ms": info["timings"]["h2d_ms"], "write_ms_slowest_shard": info["timings"]["d2h_ms"],
                     "kernel_launches": info["kernel_launches"]})
        nsuf = info["num_suffixes"]
    size = os.path.getsize(a.out)
    # ---- read the file back and verify it on GPU 0
    t0 = time.perf_counter()
    import sufrfile
    with open(a.out, "rb") as f:
        head = f.read(1 << 20)
    hdr = sufrfile.parse_header(head)
    w = 4 if hdr["text_len"] < 0xFFFFFFFF and a.index_bits != 64 else 8
    dt = np.uint32 if w == 4 else np.uint64
    text = torch.from_numpy(np.fromfile(a.out, dtype=np.uint8, count=hdr["text_len"], offset=hdr["text_pos"])).cuda()
    pad = torch.zeros(text.numel() + 16, dtype=torch.uint8, device="cuda:0")
    pad[: text.numel()] = text
    sa = torch.from_numpy(np.fromfile(a.out, dtype=dt, count=hdr["num_suffixes"], offset=hdr["sa_pos"]).view(np.int64 if w == 8 else np.int32)).cuda()
    lcp = torch.from_numpy(np.fromfile(a.out, dtype=dt, count=hdr["num_suffixes"], offset=hdr["lcp_pos"]).view(np.int64 if w == 8 else np.int32)).cuda()
    read_s = time.perf_counter() - t0
    res = _lib.Result()
    res.index_bits = 8 * w
    res.memory = S.MEM_DEVICE
    res.text_len = hdr["text_len"]
    res.num_suffixes = res.total_suffixes = hdr["num_suffixes"]
    res.text = pad.data_ptr()
    res.sa = sa.data_ptr()
    res.lcp = lcp.data_ptr()
    cargs = S.builder._CArgs(S.SufrBuilderArgs(text=b"", is_dna=True))
    rep = _lib.VerifyReport()
    rc = _lib.lib().sufr_b200_verify(ctx.handle, C.byref(cargs.c), C.byref(res), 0, 0, C.byref(rep))
    assert rc == 0, _lib.lib().sufr_b200_last_error()
    r = rep.as_dict()
    errors = r["order_errors"] + r["lcp_errors"] + r["out_of_range"] + r["not_indexed"] + r["duplicates"]
    same_text = bool(torch.equal(text.cpu(), torch.from_numpy(h_text)))  # --dna on uppercase ACGT: transform is identity
    same_text = same_text and info["text"] == h_text.tobytes()             # and the builder's `text` field
    print(json.dumps({
        "what": f"sufr_b200_create_multi: {a.bases} bp synthetic genome in 24 records, u{8 * w} SA + LCP, {a.gpus} GPU(s), "
                f"host text -> one .sufr file at {a.out}",
        "gpus": a.gpus, "text_len": text_len, "num_suffixes": nsuf, "file_bytes": size, "runs": runs,
        "suffixes_per_s_wall": nsuf / min(x["wall_s"] for x in runs),
        "verify": {"method": "file read back, sufr_b200_verify on GPU 0 (every pair, every position)", "read_back_s": round(read_s, 1),
                   "header_num_suffixes": hdr["num_suffixes"], "pairs_checked": r["pairs_checked"], "errors": errors,
                   "expected_suffixes": r["expected_suffixes"], "text_section_identical": same_text,
                   "names": hdr.get("names", [])[:2], "ok": errors == 0 and r["expected_suffixes"] == hdr["num_suffixes"] and same_text}}))
    os.unlink(a.out)


if __name__ == "__main__":
    main()
