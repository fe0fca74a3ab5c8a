#!/usr/bin/env python
"""Run ONE kernel of the hot path a few times on fresh buffers (target for ncu).
    python tools/run_one.py gemv 28672 8192 | quant bf16 | dequant bf16 | stats | igemm"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from bnb_b200 import functional as F  # noqa: E402

what = sys.argv[1]
torch.manual_seed(0)
if what == "gemv":
    N, K = int(sys.argv[2]), int(sys.argv[3])
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    packs = [q.clone() for _ in range(4)]
    x = torch.randn(1, K, device="cuda").bfloat16()
    for i in range(8):
        y = F.gemv_4bit(x, packs[i % 4].t(), state=st)
elif what in ("quant", "dequant"):
    dt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[sys.argv[2]]
    n = 4096 * 4096
    src = [torch.randn(n, device="cuda").to(dt) for _ in range(4)]
    for i in range(8):
        q, st = F.quantize_4bit(src[i % 4], blocksize=64, quant_type="nf4")
        if what == "dequant":
            F.dequantize_4bit(q, st)
elif what == "gemm4":
    N, K, batch = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
    packs = [q.clone() for _ in range(4)]
    x = torch.randn(batch, K, device="cuda").bfloat16()
    for i in range(8):
        y = F.gemm_4bit(x, packs[i % 4].t(), st)
elif what == "stats":
    A = [torch.randn(4096, 4096, device="cuda").half() for _ in range(4)]
    for i in range(8):
        F.double_quant(A[i % 4], threshold=6.0)
elif what == "igemm":
    m, k, n = 4096, 4096, 16384
    CA = torch.randint(-127, 128, (m, k), dtype=torch.int8, device="cuda")
    CB = torch.randint(-127, 128, (n, k), dtype=torch.int8, device="cuda")
    SCA = torch.rand(m, device="cuda") + 0.5
    SCB = torch.rand(n, device="cuda") + 0.5
    for i in range(3):
        F.int8_linear_dequant(CA, CB, SCA, SCB)
torch.cuda.synchronize()
print("done", what)
