#!/usr/bin/env python
"""Accuracy of the fused 4-bit GEMM routes against the fp64 product with exact (fp32 code * fp32 absmax) weights."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bnb_b200 import functional as F  # noqa: E402

torch.manual_seed(0)
for (batch, N, K, dt) in [(16, 256, 1024, torch.bfloat16), (64, 256, 1024, torch.bfloat16), (5, 130, 1024, torch.bfloat16),
                          (32, 4096, 4096, torch.bfloat16), (16, 1024, 4096, torch.float16)]:
    W = (torch.randn(N, K, device="cuda") * 0.02).to(dt)
    x = torch.randn(batch, K, device="cuda").to(dt)
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
    W32 = F.dequantize_4bit(q, F.QuantState(absmax=st.absmax, shape=st.shape, code=st.code, blocksize=64, quant_type="nf4", dtype=torch.float32))
    exact = (x.double() @ W32.double().t()).cpu().numpy()
    y = F.gemm_4bit(x, q.t(), st).double().cpu().numpy()
    yref = torch.nn.functional.linear(x, F.dequantize_4bit(q, st).to(dt)).double().cpu().numpy()
    rl2 = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))   # noqa: E731
    rms = float(np.sqrt(np.mean(exact ** 2)))
    d = np.abs(y - exact)
    print(dict(batch=batch, N=N, K=K, dtype=str(dt), relL2_kernel=rl2(y, exact), relL2_ref_composition=rl2(yref, exact),
               max_abs_over_rms=float(d.max() / rms), worst_excess_over_2p8rel=float((d - 2.0 ** -8 * np.abs(exact)).max() / rms)))
