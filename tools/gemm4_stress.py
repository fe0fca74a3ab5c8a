#!/usr/bin/env python
"""Stress the fused 4-bit GEMM under cold caches: flush L2, launch, compare with the fp64 reference, count bad launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from bnb_b200 import functional as F  # noqa: E402

torch.manual_seed(0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (batch, N, K) in [(32, 4096, 4096), (128, 4096, 4096), (16, 14336, 4096), (256, 4096, 14336), (80, 4096, 4096)]:
    dt = torch.bfloat16
    W = (torch.randn(N, K, device="cuda") * 0.02).to(dt)
    x = torch.randn(batch, K, device="cuda").to(dt)
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
    ref = (x.double() @ F.dequantize_4bit(q, st).to(dt).double().t())
    thr = 0.05 * ref.abs().mean()
    bad_launches, worst = 0, 0
    for it in range(iters):
        flush.fill_(it & 255)
        if it % 3 == 0:
            torch.cuda.synchronize()
        y = F.gemm_4bit(x, q.t(), st).double()
        nbad = int(((y - ref).abs() > thr).sum())
        bad_launches += nbad > 0
        worst = max(worst, nbad)
    print(dict(batch=batch, N=N, K=K, launches=iters, bad_launches=bad_launches, worst_bad_outputs=worst), flush=True)
