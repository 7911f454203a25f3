#!/usr/bin/env python
"""One build of the headline workload inside a cudaProfilerStart/Stop range (after two warm-up builds), for

    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernels> \
        -o gpurun_out/prof python tools/profile_build.py [bases] [index_bits] [rank world_size]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bench  # noqa: E402
import sufr_b200 as S  # noqa: E402
from sufr_b200 import _lib  # noqa: E402

bases = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
shard = dict(rank=int(sys.argv[3]), world_size=int(sys.argv[4])) if len(sys.argv) > 4 else {}
text_len, starts = bench.record_layout(bases)
ctx = S.Context(0)
d_text = torch.empty(text_len, dtype=torch.uint8, device="cuda")
st = np.asarray(starts, dtype=np.uint64)
assert _lib.lib().sufr_b200_synth_dna(ctx.handle, d_text.data_ptr(), text_len, bench.SEED, st.ctypes.data, len(st), ord("%")) == 0
bargs = S.SufrBuilderArgs(text=b"", is_dna=True, sequence_starts=starts, sequence_names=[f"chr{i + 1}" for i in range(len(starts))])
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    r = S.build(bargs, index_bits=bits, ctx=ctx, result_memory=S.MEM_DEVICE, device_text=(d_text.data_ptr(), text_len), **shard)
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.profiler.stop()
    print(f"build {i}: {r.num_suffixes} suffixes, {r.timings['total_ms']:.2f} ms on device, {r.kernel_launches} launches", flush=True)
    r.free()
