"""Where a batch of config-4 plan solves (reference guess, reference IPOPT options, limited-memory Hessian) spends its GPU
time: torch.profiler totals per kernel over the whole solve.   usage: profile_plans.py [-b BATCH]"""
import runpy
import sys

import torch
from torch.profiler import ProfilerActivity, profile

batch = sys.argv[sys.argv.index("-b") + 1] if "-b" in sys.argv else "148"
sys.argv = ["check_periodic_step.py", "-s", "--ref-options", "--lbfgs", "-b", batch]
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    runpy.run_path("tools/check_periodic_step.py", run_name="__main__")
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
