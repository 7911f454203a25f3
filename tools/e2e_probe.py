#!/usr/bin/env python
"""Where the end-to-end time of a host-to-host build goes: one 3.1 Gbp u64 build per setting of the host-side knobs
(SUFR_B200_WIDEN_THREADS, plain instead of compact transfer), with the library's own e2e log on stderr."""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["SUFR_B200_LOG_E2E"] = "1"
import torch  # noqa: E402
import bench  # noqa: E402
import sufr_b200 as S  # noqa: E402
from sufr_b200 import _lib  # noqa: E402

bases = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
text_len, starts = bench.record_layout(bases)
ctx = S.Context(0)
d_text = torch.empty(text_len, dtype=torch.uint8, device="cuda")
st = np.asarray(starts, dtype=np.uint64)
assert _lib.lib().sufr_b200_synth_dna(ctx.handle, d_text.data_ptr(), text_len, bench.SEED, st.ctypes.data, len(st), ord("%")) == 0
h_text = torch.empty(text_len, dtype=torch.uint8, pin_memory=True)
h_text.copy_(d_text)
del d_text
torch.cuda.synchronize()
bargs = S.SufrBuilderArgs(text=memoryview(h_text.numpy()), is_dna=True, sequence_starts=starts,
                          sequence_names=[f"chr{i + 1}" for i in range(len(starts))])
settings = [("default", {}), ("4 widening threads", {"SUFR_B200_WIDEN_THREADS": "4"}),
            ("8 widening threads", {"SUFR_B200_WIDEN_THREADS": "8"}),
            ("plain u64 transfer", {"SUFR_B200_DEBUG_NO_COMPACT_D2H": "1"})]
if len(sys.argv) > 2 and sys.argv[2] == "default":  # e.g. to compare SUFR_B200_DEBUG_NO_AVX512=1 between two runs
    settings = settings[:1]
for bits in (64, 32):
    for name, env in settings:
        for k, v in env.items():
            os.environ[k] = v
        for rep in range(2):
            t0 = time.perf_counter()
            r = S.build(bargs, index_bits=bits, ctx=ctx, result_memory=S.MEM_HOST)
            dt = time.perf_counter() - t0
            tm = r.timings
            if rep:
                print(f"u{bits} {name:22s}: {1e3 * dt:7.1f} ms wall | h2d {tm['h2d_ms']:.1f} build {tm['total_ms']:.1f} "
                      f"d2h(events) {tm['d2h_ms']:.1f} | d2h bytes {r.d2h_bytes / 1e9:.2f} GB", flush=True)
            r.free()
        for k in env:
            os.environ.pop(k, None)
