#!/usr/bin/env python
"""One BASELINE config as key-range shards under torchrun (one process per GPU): device time of the slowest rank,
full verification of every shard.  usage: torchrun --nproc-per-node N tools/sharded_config.py config2b [scale]"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
import sufr_b200 as S  # noqa: E402
import workloads  # noqa: E402
from sufr_b200.distributed import finish_shard  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2b"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
w = workloads.ALL[name](int(bench.FULL_SIZES[name] * scale))
t = torch.frombuffer(bytearray(w.text), dtype=torch.uint8).to(dev)
ctx = S.Context(local)
args = S.SufrBuilderArgs(text=b"", sequence_starts=w.sequence_starts, sequence_names=w.sequence_names, **w.flags)
times = []
for i in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = S.build(args, index_bits=w.index_bits, ctx=ctx, result_memory=S.MEM_DEVICE, device_text=(t.data_ptr(), t.numel()),
                rank=rank, world_size=world)
    meta = finish_shard(r) if world > 1 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if i:
        times.append(dt)
    info = (r.num_suffixes, r.total_suffixes, int(r.c.refine_rounds), int(r.c.doubling_rounds), r.timings)
    ver = bench.verify_full(r, meta, rank, world, dev) if i == 2 else None
    r.free()
el = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
gathered = [None] * world
mine = {"rank": rank, "suffixes": info[0], "refine_rounds": info[2], "doubling_rounds": info[3],
        "phases_ms": {k: round(v, 2) for k, v in info[4].items() if k.endswith("_ms")}}
if world > 1:
    dist.all_gather_object(gathered, mine)
else:
    gathered = [mine]
if rank == 0:
    ms = 1e3 * float(el.item())
    print(json.dumps({"config": name, "text_len": t.numel(), "n_gpus": world, "ms_per_build": ms, "suffixes": info[1],
                      "suffixes_per_s": info[1] / (ms * 1e-3), "verify": ver, "ranks": gathered}))
if world > 1:
    dist.destroy_process_group()
