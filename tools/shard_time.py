#!/usr/bin/env python
"""Device time of single key-range shards of a BASELINE config, emulated on ONE GPU (rank r of N builds only its key
range): usage: tools/shard_time.py <scale> <world>, e.g. `tools/shard_time.py 1.0 8` for shards 1 and 7 of 8 of config 2b."""
import sys, time, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import torch, bench, sufr_b200 as S, workloads
w = workloads.ALL["config2b"](int(bench.FULL_SIZES["config2b"] * float(sys.argv[1])))
t = torch.frombuffer(bytearray(w.text), dtype=torch.uint8).cuda()
args = S.SufrBuilderArgs(text=b"", sequence_starts=w.sequence_starts, sequence_names=w.sequence_names, **w.flags)
world = int(sys.argv[2])
for rank in (1, world - 1):
    for i in range(2):
        r = S.build(args, index_bits=w.index_bits, result_memory=S.MEM_DEVICE, device_text=(t.data_ptr(), t.numel()), rank=rank, world_size=world)
        torch.cuda.synchronize()
        tm = r.timings
        if i: print(f"shard {rank} of {world}: {tm['total_ms']:.1f} ms (refine {tm['refine_ms']:.1f}), {r.c.refine_rounds} word rounds, {r.num_suffixes} suffixes", flush=True)
        r.free()
