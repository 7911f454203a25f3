#!/usr/bin/env python
"""Per-round log of the refinement (SUFR_B200_LOG_ROUNDS=1) for one BASELINE config: where the time of a deep-repeat
build goes.  usage: tools/round_log.py config5 [scale]"""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
if not os.environ.get("SUFR_NO_ROUND_LOG"):
    os.environ["SUFR_B200_LOG_ROUNDS"] = "1"
import torch  # noqa: E402
import sufr_b200 as S  # noqa: E402
import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config5"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
full = {"config2b": 3_100_000_000, "config5": 1_000_000_000, "config3": 1_000_000_000, "config4": 500_000_000}
w = workloads.ALL[name](int(full[name] * scale))
t = torch.frombuffer(bytearray(w.text), dtype=torch.uint8).cuda()
args = S.SufrBuilderArgs(text=b"", sequence_starts=w.sequence_starts, sequence_names=w.sequence_names, **w.flags)
for i in range(int(os.environ.get("SUFR_BUILDS", "2"))):
    print(f"--- build {i}", file=sys.stderr, flush=True)
    if i == 1 and os.environ.get("SUFR_PROFILE"):  # ncu --profile-from-start off: the second build only
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    t0 = time.time()
    r = S.build(args, index_bits=w.index_bits, result_memory=S.MEM_DEVICE, device_text=(t.data_ptr(), t.numel()))
    torch.cuda.synchronize()
    if i == 1 and os.environ.get("SUFR_PROFILE"):
        torch.cuda.profiler.stop()
    print(f"--- {name}: {r.num_suffixes} suffixes, wall {time.time() - t0:.3f} s, phases {r.timings}", file=sys.stderr, flush=True)
    r.free()
