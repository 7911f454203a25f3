"""Single-process probe of one shard build (rank r of w) on a synthetic text: used under ncu to see the
per-rank fixed costs of the sharded path."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch, bench
import sufr_b200 as S
from sufr_b200 import _lib
bases = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
text_len, starts = bench.record_layout(bases)
ctx = S.Context(0)
d = torch.empty(text_len, dtype=torch.uint8, device="cuda")
st = np.asarray(starts, dtype=np.uint64)
assert _lib.lib().sufr_b200_synth_dna(ctx.handle, d.data_ptr(), text_len, 2, st.ctypes.data, len(st), ord("%")) == 0
args = S.SufrBuilderArgs(text=b"", is_dna=True)
for _ in range(2):
    r = S.build(args, index_bits=64, ctx=ctx, result_memory=S.MEM_DEVICE, device_text=(d.data_ptr(), text_len), rank=0, world_size=world)
    print(r.num_suffixes, {k: round(v, 2) for k, v in r.timings.items() if k.endswith("_ms")})
    r.free()
