import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
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
