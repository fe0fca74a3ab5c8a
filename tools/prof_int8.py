#!/usr/bin/env python
"""Kernel-level timeline of one Linear8bitLt forward (BASELINE config 3) with torch.profiler."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch
import bnb_b200
from torch.profiler import profile, ProfilerActivity
m, k, n = 4096, 4096, 16384
torch.manual_seed(0)
A = torch.randn(m, k, device="cuda").half()
A[:, [7, 100, 2000, 3000]] = 8.0
lin = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
with torch.no_grad():
    for _ in range(3):
        lin(A)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            lin(A)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
