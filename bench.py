#!/usr/bin/env python
"""bench.py -- suffixes/second of SA+LCP construction (BASELINE.json metric) on B200.

A "step" is one complete build (encode -> keys -> sort -> refine -> LCP -> finish) of the synthetic
3.1 Gbp DNA text of BASELINE.json configs[1] (24 records with chromosome-proportional lengths, iid
uniform ACGT, '%' between records, '$' at the end; u64 SA/LCP as the config names).

  value : text already resident in HBM, SA/LCP left in HBM (device-timed, whole job over all ranks)
  e2e   : the same build through the C ABI with HOST buffers -- pinned text copied H2D, SA/LCP/text
          copied D2H -- all inside the timed region
  roofline     : the dominant kernel (radix-sort downsweep), CUDA-event timed per launch inside the step
  cpu_baseline : the oracle restatement of the reference's rayon path on a bounded prefix of the same text

`--impl reference` times the reference's CPU algorithm (oracle port; the Rust reference cannot be compiled
in this image) with all host threads on a bounded sample of the same workload.

N > 1 (torchrun): the text is replicated, rank r builds key-range shard r (no data-path collective), one
all_gather of (count, first, last) repairs the seam LCPs.  Fixed total work => "scaling": "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# human chromosome lengths (Mbp, GRCh38 1..22, X, Y): only the proportions matter
CHROM_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51,
             156, 57]
SEED = 2


def record_layout(total_bases: int):
    """24 record lengths proportional to the human chromosomes, summing to total_bases; returns
    (text_len, record_starts) with one delimiter between records and the trailing '$'."""
    tot = sum(CHROM_MBP)
    lens = [max(1, total_bases * c // tot) for c in CHROM_MBP]
    lens[0] += total_bases - sum(lens)
    starts, pos = [], 0
    for ln in lens:
        starts.append(pos)
        pos += ln + 1  # delimiter (or the final '$')
    return pos, starts


def synth_prefix_numpy(n: int, text_len: int, starts, seed: int = SEED) -> bytes:
    """CPU restatement of synth_dna_kernel / synth_marks_kernel for the first n bytes of the text."""
    i = np.arange(1, n + 1, dtype=np.uint64)
    c = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * c + i * c
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    out = np.frombuffer(b"ACGT", dtype=np.uint8)[(z >> np.uint64(62)).astype(np.int64)].copy()
    for s in starts[1:]:
        if 1 <= s <= n:
            out[s - 1] = ord("%")
    if n == text_len:
        out[n - 1] = ord("$")
    return out.tobytes()


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md): through NVML every
    10 ms when pynvml is importable (a multi-GPU timed region lasts ~150 ms), else through nvidia-smi."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]  # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index: int, uuid: str = None):
        self.index, self.uuid = index, uuid
        self.samples, self._stop, self._t = [], threading.Event(), None  # [sm, sm_max, watts, 4 x "Active"/"Not Active"]
        self.source = "nvidia-smi"

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _run(self):
        try:
            nv, h = self._nvml_handle()
            sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.source = "nvml"
            while not self._stop.is_set():
                mask = reasons_fn(h)
                self.samples.append([nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), sm_max, nv.nvmlDeviceGetPowerUsage(h) / 1000.0]
                                    + ["Active" if mask & b else "Not Active" for b in self.BITS])
                self._stop.wait(0.01)
            return
        except Exception:
            if self.samples:
                return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = [nm for k, nm in enumerate(self.NAMES) if any(str(s[3 + k]).lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(s[2]) for s in self.samples), "source": self.source}


def oracle_run(sample: bytes, threads: int, partitions: int):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle as O
    t0 = time.perf_counter()
    r = O.oracle_build(sample, is_dna=True, num_partitions=partitions, threads=threads, index_bits=32)
    dt = time.perf_counter() - t0
    return r, dt


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle port) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    text_len, starts = record_layout(args.bases)
    sample_n = min(args.cpu_sample, text_len)
    sample = synth_prefix_numpy(sample_n, text_len, starts)
    if sample_n < text_len:
        sample = sample[:-1] + b"$"
    parts = max(16, 4 * cores)
    for _ in range(args.warmup):
        oracle_run(sample, cores, parts)
    t = 0.0
    nsuf = 0
    for _ in range(args.steps):
        r, dt = oracle_run(sample, cores, parts)
        t += dt
        nsuf += r.num_suffixes
    value = nsuf / t
    line = {
        "impl": "reference", "metric": "suffixes/sec (SA+LCP)", "value": value, "unit": "suffixes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, text_len),
        "cpu_baseline": {"value": value, "unit": "suffixes/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample_n} bytes of the workload text (u32 indices, -n {parts}), "
                                   f"oracle restatement of sufr_builder.rs with {cores} threads, partitions in RAM"},
        "e2e": {"value": value, "unit": "suffixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, text_len):
    return {"workload": f"sufr create --dna, {args.bases} bp synthetic iid ACGT in 24 records (BASELINE configs[1]), "
                        f"u{args.index_bits} SA+LCP", "text_len": text_len, "index_bits": args.index_bits,
            "flags": "--dna -n 16", "seed": SEED,
            "l2_policy": "inputs (>= 3 GB text, >= 37 GB key/position arrays) are far larger than the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=int, default=3_100_000_000)
    ap.add_argument("--index-bits", type=int, default=64, choices=[32, 64])
    ap.add_argument("--cpu-sample", type=int, default=32_000_000)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--verify", type=int, default=4000, help="sampled SA/LCP checks after the timed region")
    ap.add_argument("--repetitive", action="store_true",
                    help="also time the repetitive variant (config 2b, generated on the host: adds ~1 min)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import sufr_b200 as S
    from sufr_b200.distributed import finish_shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    text_len, starts = record_layout(args.bases)
    ctx = S.Context(local_rank)
    d_text = torch.empty(text_len, dtype=torch.uint8, device=dev)
    from sufr_b200 import _lib
    import ctypes as C
    st = np.asarray(starts, dtype=np.uint64)
    rc = _lib.lib().sufr_b200_synth_dna(ctx.handle, d_text.data_ptr(), text_len, SEED, st.ctypes.data, len(st), ord("%"))
    assert rc == 0, _lib.lib().sufr_b200_last_error()
    bargs = S.SufrBuilderArgs(text=b"", is_dna=True, num_partitions=16, sequence_starts=starts,
                              sequence_names=[f"chr{i + 1}" for i in range(len(starts))])

    def step_device():
        r = S.build(bargs, index_bits=args.index_bits, ctx=ctx, result_memory=S.MEM_DEVICE,
                    device_text=(d_text.data_ptr(), text_len), rank=rank, world_size=world)
        if world > 1:
            finish_shard(r)
        return r

    # ---------------- value: resident text -> resident SA/LCP
    for _ in range(args.warmup):
        step_device().free()
    barrier()
    results = []
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    with ClockSampler(local_rank, gpu_uuid) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = step_device()
            results.append((r.num_suffixes, r.total_suffixes, r.timings, r.kernel_launches))
            last = r
            if _ + 1 < args.steps:
                r.free()
        barrier()
        elapsed = time.perf_counter() - t0
    el = torch.tensor([elapsed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed = float(el.item())
    total_suffixes = results[-1][1]
    value = total_suffixes * args.steps / elapsed
    tm = results[-1][2]
    launches = sum(x[3] for x in results)

    # ---------------- sampled verification of the last result (outside the timed region)
    verify = verify_sample(last, d_text, args.verify, rank, world) if args.verify else None
    last.free()

    # ---------------- e2e: host text in, host SA/LCP out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, S, ctx, d_text, text_len, bargs, rank, world, dev, barrier)

    # ---------------- CPU baseline on a bounded prefix (rank 0, N=1 only)
    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        sample_n = min(args.cpu_sample, text_len)
        sample = d_text[:sample_n].cpu().numpy().tobytes()
        if sample_n < text_len:
            sample = sample[:-1] + b"$"
        parts = max(16, 4 * cores)
        r, dt = oracle_run(sample, cores, parts)
        cpu = {"value": r.num_suffixes / dt, "unit": "suffixes/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_n} bytes of the workload text (u32 indices, -n {parts}); oracle "
                         f"restatement of sufr_builder.rs, {cores} threads, partitions in RAM; {dt:.2f} s"}

    # ---------------- repetitive variant of the same workload (BASELINE: "random and repetitive FASTA"), N=1 only
    variants = None
    if args.repetitive and world == 1:
        variants = {"repetitive": run_repetitive(args, S, ctx, dev)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dk_ms = tm["dominant_kernel_ms"] / max(1, tm["dominant_kernel_launches"])
        achieved = tm["dominant_kernel_bytes"] / (dk_ms * 1e-3) / 1e9 if dk_ms > 0 else 0.0
        w = args.index_bits // 8
        job_bytes = text_len + 2 * total_suffixes * w
        line = {
            "metric": "suffixes/sec (SA+LCP)", "value": value, "unit": "suffixes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64" if args.index_bits == 64 else "u32", "data": "synthetic",
            "config": workload_config(args, text_len),
            "clocks": clocks.summary(),
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "rsort::downsweep_kernel<u64,u32> (main sort)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s",
                         "traffic": traffic_estimate(tm["dominant_kernel_bytes"]), "traffic_source": TRAFFIC_SOURCE,
                         "algorithmic_bytes": tm["dominant_kernel_bytes"], "launch_ms": dk_ms, "launches_per_step": tm["dominant_kernel_launches"],
                         "share_of_step": tm["dominant_kernel_ms"] / (1e3 * elapsed / args.steps),
                         "job": {"algorithmic_bytes": job_bytes,
                                 "achieved": job_bytes * args.steps / elapsed / 1e9 / world,
                                 "frac": job_bytes * args.steps / elapsed / 1e9 / world / peak,
                                 "note": "A = n + 2*s*sizeof(T) per SURVEY 8(d), per GPU"}},
            "cpu_baseline": cpu,
            "variants": variants,
            "phases_ms": {k: v for k, v in tm.items() if k.endswith("_ms")},
            "verify": verify,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


TRAFFIC_SOURCE = ("dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on the "
                  "200 Mbp workload (profiles/r1_v7_downsweep_raw.csv: 2.781 GB + 2.686 GB for 4.800 GB "
                  "algorithmic), scaled to this launch's element count")


def traffic_estimate(algorithmic_bytes):
    """DRAM bytes per launch of the dominant kernel (see TRAFFIC_SOURCE)."""
    return algorithmic_bytes * (2.781094 + 2.685538) / 4.800000576


def run_repetitive(args, S, ctx, dev):
    """One warm-up + two timed device-resident builds of workloads.config2_repetitive at the same size."""
    import torch
    import workloads
    sys.path.insert(0, str(ROOT / "tools"))
    w = workloads.config2_repetitive(args.bases)
    t = torch.frombuffer(bytearray(w.text), dtype=torch.uint8).to(dev)
    bargs = S.SufrBuilderArgs(text=b"", is_dna=True, sequence_starts=w.sequence_starts, sequence_names=w.sequence_names)
    times, last = [], None
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = S.build(bargs, index_bits=args.index_bits, ctx=ctx, result_memory=S.MEM_DEVICE,
                    device_text=(t.data_ptr(), t.numel()))
        torch.cuda.synchronize()
        if i:
            times.append(time.perf_counter() - t0)
        info = (r.num_suffixes, int(r.c.refine_rounds), int(r.c.doubling_rounds), r.timings)
        # against the TRANSFORMED text the index is built on (soft-masked stretches are upper-cased)
        v = verify_sample(r, r.text_tensor(), 1000, 0, 2) if i == 2 else None
        r.free()
    ms = 1e3 * sum(times) / len(times)
    return {"workload": "config 2b: ~50 % of the bases are mutated copies of earlier 0.3-6 kb segments, some N / "
                        "soft-masked stretches", "ms_per_step": ms, "value": info[0] / (ms * 1e-3),
            "unit": "suffixes/s", "refine_rounds": info[1], "doubling_rounds": info[2],
            "phases_ms": {k: v for k, v in info[3].items() if k.endswith("_ms")}, "verify": v}


def verify_sample(res, d_text, k, rank, world):
    """Size-independent checks at full size: the shard's SA holds distinct positions in suffix order and
    the LCP values are exact, on k sampled adjacent pairs (compared on the host from text windows)."""
    import torch
    s = res.num_suffixes
    if s < 2:
        return {"pairs": 0}
    sa, lcp = res.sa_tensor(), res.lcp_tensor()
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + rank)
    j = torch.randint(1, s, (k,), generator=g).to(sa.device)
    # u32 results are viewed as int32 tensors: positions >= 2^31 come out negative, mask them back
    wrap = 0xFFFFFFFF if sa.dtype == torch.int32 else 0x7FFFFFFFFFFFFFFF
    a = sa[j - 1].cpu().numpy().astype(np.int64) & wrap
    b = sa[j].cpu().numpy().astype(np.int64) & wrap
    l = lcp[j].cpu().numpy().astype(np.int64) & wrap
    n = d_text.numel()
    W = 256
    bad = 0
    for x, y, ll in zip(a.tolist(), b.tolist(), l.tolist()):
        ta = bytes(d_text[x:min(n, x + W)].cpu().numpy().tobytes())
        tb = bytes(d_text[y:min(n, y + W)].cpu().numpy().tobytes())
        c = 0
        while c < len(ta) and c < len(tb) and ta[c] == tb[c]:
            c += 1
        if c >= W:
            continue  # deeper than the window: skip
        ok_order = ta[c:c + 1] < tb[c:c + 1] if c < len(ta) and c < len(tb) else len(ta) < len(tb)
        if not ok_order or c != ll:
            bad += 1
    out = {"pairs": k, "mismatches": bad}
    if world == 1:
        tot = int(sa.sum(dtype=torch.int64).item())
        if sa.dtype == torch.int32:
            tot += int((sa < 0).sum().item()) << 32
        # all positions except the delimiters ('%' is not indexed under --dna)
        expect = n * (n - 1) // 2 - int(torch.nonzero(d_text == ord("%")).sum().item())
        out["position_sum_ok"] = (tot == expect)
    return out


def run_e2e(args, S, ctx, d_text, text_len, bargs, rank, world, dev, barrier):
    import torch
    import torch.distributed as dist
    from sufr_b200.distributed import finish_shard
    h_text = torch.empty(text_len, dtype=torch.uint8, pin_memory=True) if rank == 0 or world == 1 else None
    if h_text is not None:
        h_text.copy_(d_text)
    torch.cuda.synchronize()
    staging = torch.empty(text_len, dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        if world == 1:
            return S.build(_with_text(bargs, h_text), index_bits=args.index_bits, ctx=ctx,
                           result_memory=S.MEM_HOST)
        # N > 1: rank 0 uploads, NCCL broadcast replicates the text, every rank builds and downloads its shard
        if rank == 0:
            staging.copy_(h_text, non_blocking=True)
        dist.broadcast(staging, src=0)
        torch.cuda.synchronize()
        r = S.build(bargs, index_bits=args.index_bits, ctx=ctx, result_memory=S.MEM_HOST,
                    device_text=(staging.data_ptr(), text_len), rank=rank, world_size=world)
        finish_shard(r)
        return r

    step().free()  # warm-up: page-locks the result buffers once
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    h2d = 0
    out_bytes = 0
    tot = 0
    for _ in range(args.e2e_steps):
        r = step()
        d2h = r.d2h_bytes                      # bytes that crossed PCIe (compact encoding, see sufr_b200.h)
        h2d = r.h2d_bytes if world == 1 else (text_len if rank == 0 else 0)
        out_bytes = (r.text_len if rank == 0 else 0) + 2 * r.num_suffixes * (args.index_bits // 8)
        tot = r.total_suffixes
        _ = int(r.sa[:1024].sum()) if r.num_suffixes else 0  # read the step's result on the host
        r.free()
    barrier()
    elapsed = time.perf_counter() - t0
    el = torch.tensor([elapsed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed = float(el.item())
    return {"value": tot * args.e2e_steps / elapsed, "unit": "suffixes/s",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_result_bytes_per_step": out_bytes,
            "ms_per_step": 1e3 * elapsed / args.e2e_steps, "steps": args.e2e_steps,
            "note": "pinned host text -> H2D -> build -> D2H (LCP as bytes + exceptions, u64 SA as u32, widened by host "
                    "threads into the pinned u64 result arrays), per step; bytes are those of rank 0"}


def _with_text(bargs, h_text):
    """SufrBuilderArgs whose text is a zero-copy view of the pinned host tensor."""
    import copy
    b = copy.copy(bargs)
    b.text = memoryview(h_text.numpy())
    return b


if __name__ == "__main__":
    main()
