"""Per-kernel GPU durations of a bench.py run measured in the pipeline (CUPTI through torch.profiler: warm caches,
no serialisation), to set beside the cold-cache ncu launch lists.  Usage: python scripts/kernel_times.py --workload c5"""
import os
import runpy
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:] + ["--no-e2e", "--no-cpu-baseline", "--no-other-workloads", "--no-ncu-side-run", "--steps", "50", "--warmup", "5"]
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    try:
        runpy.run_path(sys.argv[0], run_name="__main__")
    except SystemExit:
        pass
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
for e in rows[:25]:
    print("%-90s n=%5d mean=%9.2f us total=%10.1f us" % (e.key[:90], e.count, e.device_time_total / max(e.count, 1), e.device_time_total))
