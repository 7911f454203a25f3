"""Synthetic inputs for the five BASELINE.json configs (generator specs: SURVEY.md section 8(d)).

Numpy on the host, seeded and size-parametrised: the GPU parity tests run them at sizes the oracle
finishes in seconds, tools/run_configs.py at the full sizes.  (bench.py's headline workload, config 2
variant (a), is generated on the device by sufr_b200_synth_dna; `config2_random` here is its host twin.)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

CHROM_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51,
             156, 57]
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
AMINO = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


@dataclass
class Workload:
    name: str
    text: bytes
    flags: dict
    sequence_starts: List[int] = field(default_factory=lambda: [0])
    sequence_names: List[str] = field(default_factory=lambda: ["1"])
    index_bits: int = 32
    note: str = ""


def _join_records(records: List[np.ndarray], delim: int = ord("%")):
    starts, parts, pos = [], [], 0
    for i, r in enumerate(records):
        if i:
            parts.append(np.array([delim], dtype=np.uint8))
            pos += 1
        starts.append(pos)
        parts.append(r)
        pos += len(r)
    parts.append(np.array([ord("$")], dtype=np.uint8))
    return np.concatenate(parts), starts


def config1(n: int = 10_000_000, seed: int = 1) -> Workload:
    """sufr create --dna -n 16 on one record of n iid-uniform ACGT (u32)."""
    rng = np.random.default_rng(seed)
    text, starts = _join_records([ACGT[rng.integers(0, 4, n)]])
    return Workload("config1", text.tobytes(), dict(is_dna=True, num_partitions=16), starts, ["seq1"])


def _chrom_lengths(total: int):
    tot = sum(CHROM_MBP)
    lens = [max(1, total * c // tot) for c in CHROM_MBP]
    lens[0] += total - sum(lens)
    return lens


def config2_random(total: int = 3_100_000_000, seed: int = 2) -> Workload:
    """24 records with chromosome-proportional lengths, iid ACGT; u64 indices as BASELINE names it."""
    rng = np.random.default_rng(seed)
    recs = [ACGT[rng.integers(0, 4, ln)] for ln in _chrom_lengths(total)]
    text, starts = _join_records(recs)
    return Workload("config2a", text.tobytes(), dict(is_dna=True), starts, [f"chr{i + 1}" for i in range(24)], 64)


def config2_repetitive(total: int = 3_100_000_000, seed: int = 2) -> Workload:
    """Variant (b): ~50 % of the bases are copies of earlier 0.3-6 kb segments with 1-10 % substitutions;
    a few lowercase (soft-masked) and N stretches (no -a: N starts are not indexed)."""
    rng = np.random.default_rng(seed + 1000)
    out = np.empty(total, dtype=np.uint8)
    pos = 0
    while pos < total:
        if pos > 10_000 and rng.random() < 0.5:
            ln = int(rng.integers(300, 6001))
            ln = min(ln, total - pos)
            src = int(rng.integers(0, pos - ln)) if pos > ln else 0
            seg = out[src:src + ln].copy()
            rate = rng.uniform(0.01, 0.10)
            mut = rng.random(ln) < rate
            seg[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
            out[pos:pos + ln] = seg
        else:
            ln = int(min(rng.integers(300, 6001), total - pos))
            out[pos:pos + ln] = ACGT[rng.integers(0, 4, ln)]
        pos += ln
    # sprinkle soft-masked blocks and N stretches (about 2 % and 0.5 % of the text)
    nblocks = max(1, total // 200_000)
    for _ in range(nblocks):
        s = int(rng.integers(0, max(1, total - 5000)))
        ln = int(rng.integers(100, 4000))
        out[s:s + ln] |= 0x20  # lowercase
    for _ in range(max(1, nblocks // 4)):
        s = int(rng.integers(0, max(1, total - 5000)))
        out[s:s + int(rng.integers(50, 4000))] = ord("N")
    lens = _chrom_lengths(total)
    recs, p = [], 0
    for ln in lens:
        recs.append(out[p:p + ln])
        p += ln
    text, starts = _join_records(recs)
    return Workload("config2b", text.tobytes(), dict(is_dna=True), starts, [f"chr{i + 1}" for i in range(24)], 64)


def config3(total: int = 1_000_000_000, seed: int = 3, record_len: int = 100_000) -> Workload:
    """Protein: records of `record_len` iid residues over the 20 amino acids, --max-query-len 32.
    Precondition of a well-defined reference result: no two indexed suffixes share 32 residues (checked by
    the harness on the output: no LCP >= 32)."""
    rng = np.random.default_rng(seed)
    nrec = max(1, total // record_len)
    recs = [AMINO[rng.integers(0, 20, record_len)] for _ in range(nrec)]
    text, starts = _join_records(recs)
    return Workload("config3", text.tobytes(), dict(max_query_len=32), starts, [f"p{i}" for i in range(nrec)])


def config4(n: int = 500_000_000, seed: int = 4) -> Workload:
    """--dna --seed-mask 1101101101 on one record of iid ACGT.  (BASELINE adds --max-query-len 64, which the
    reference rejects together with a mask: sufr/src/lib.rs:95, sufr_builder.rs:163-165.)"""
    rng = np.random.default_rng(seed)
    text, starts = _join_records([ACGT[rng.integers(0, 4, n)]])
    return Workload("config4", text.tobytes(), dict(is_dna=True, seed_mask="1101101101"), starts, ["seq1"])


def config5(total: int = 1_000_000_000, seed: int = 5, max_unit: int = 200, max_copies: int = 10_000) -> Workload:
    """--dna --allow-ambiguity --ignore-softmask on low-entropy DNA: tandem-repeat arrays (unit 1..max_unit bp,
    10..max_copies copies), AT-rich stretches, soft-masked (lowercase) blocks and literal N blocks.
    Generator constraint (SURVEY 8a rule 3): after the transform every maximal N run (literal N or lowercase,
    adjacent ones merge) is >= 1000 long, so the reference result is a function of the input."""
    rng = np.random.default_rng(seed)
    parts, size = [], 0
    last_was_n = False
    while size < total:
        kind = rng.random()
        if kind < 0.45:  # tandem array
            unit = ACGT[rng.integers(0, 4, int(rng.integers(1, max_unit + 1)))]
            copies = int(rng.integers(10, max_copies + 1))
            seg = np.tile(unit, copies)
            last_was_n = False
        elif kind < 0.70:  # AT-rich low-entropy stretch
            ln = int(rng.integers(1000, 200_000))
            seg = ACGT[rng.choice(4, size=ln, p=[0.45, 0.05, 0.05, 0.45])]
            last_was_n = False
        elif kind < 0.85:  # plain random
            seg = ACGT[rng.integers(0, 4, int(rng.integers(1000, 100_000)))]
            last_was_n = False
        elif last_was_n:
            continue  # keep N runs separated by real bases
        elif kind < 0.93:  # soft-masked block (becomes N under --ignore-softmask)
            ln = int(rng.integers(1000, 50_000))
            seg = ACGT[rng.integers(0, 4, ln)] | 0x20
            last_was_n = True
        else:  # literal N block
            seg = np.full(int(rng.integers(1000, 50_000)), ord("N"), dtype=np.uint8)
            last_was_n = True
        if size + len(seg) > total:
            if last_was_n:
                seg = ACGT[rng.integers(0, 4, total - size)]  # never truncate an N run below 1000
            else:
                seg = seg[: total - size]
        parts.append(seg)
        size += len(seg)
    text, starts = _join_records([np.concatenate(parts)])
    return Workload("config5", text.tobytes(),
                    dict(is_dna=True, allow_ambiguity=True, ignore_softmask=True), starts, ["seq1"])


ALL = {"config1": config1, "config2a": config2_random, "config2b": config2_repetitive, "config3": config3,
       "config4": config4, "config5": config5}
