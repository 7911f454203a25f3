#!/usr/bin/env python
"""End-to-end timing of the command-line path: FASTA file -> `sufr-b200 create` -> .sufr file
on a RAM disk.  Prints the CLI's own log (ingest, device phases, transfer, file write) and the wall time, then
checks the written file with the size-independent checks of tools/verify.py."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    bases = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    records = 8
    work = Path(os.environ.get("SUFR_B200_TMP", "/dev/shm")) / "sufr_b200_cli_e2e"
    work.mkdir(parents=True, exist_ok=True)
    fa, out = work / "genome.fa", work / "genome.sufr"
    rng = np.random.default_rng(7)
    t = time.time()
    with open(fa, "wb") as f:
        for r in range(records):
            f.write(b">chr%d synthetic\n" % (r + 1))
            m = bases // records
            a = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, m)]
            lines = m // 60
            b = np.empty((lines, 61), np.uint8)
            b[:, :60] = a[: lines * 60].reshape(lines, 60)
            b[:, 60] = 10
            f.write(b.tobytes())
    gen_s = time.time() - t
    res = {"bases": bases, "fasta_bytes": fa.stat().st_size, "gen_s": round(gen_s, 1), "runs": []}
    for _ in range(2):
        if out.exists():
            out.unlink()
        t = time.time()
        p = subprocess.run([str(ROOT / "sufr_b200" / "sufr-b200"), "create", "--dna", "--log", "info", "-o", str(out), str(fa)],
                           capture_output=True, text=True)
        wall = time.time() - t
        res["runs"].append({"wall_s": round(wall, 3), "rc": p.returncode, "log": p.stdout.strip().splitlines(),
                            "stderr": p.stderr.strip()[-300:]})
    # verify the file: header, text, and sampled adjacent pairs of the suffix array
    import sufrfile
    from verify import check_pairs, check_positions
    sf = sufrfile.parse_sufr(out.read_bytes())
    ranks = np.random.default_rng(3).integers(0, sf.num_suffixes, 2000)
    bad = check_pairs(sf.text, sf.sa, sf.lcp, ranks)
    res["verify"] = {"num_suffixes": int(sf.num_suffixes), "file_bytes": out.stat().st_size, "pairs": len(ranks),
                     "mismatches": len(bad), "positions_ok": bool(check_positions(sf.text, sf.sa, is_dna=True)),
                     "num_sequences": sf.num_sequences, "names": sf.sequence_names[:2]}
    print(json.dumps(res))
    fa.unlink()
    out.unlink()


if __name__ == "__main__":
    main()
