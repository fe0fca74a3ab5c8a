import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch
from bnb_b200 import functional as F
torch.manual_seed(0)
for batch, N, K in ((16, 14336, 4096), (32, 14336, 4096), (64, 14336, 4096)):
    W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    x = torch.randn(batch, K, device="cuda").to(torch.bfloat16)
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(2):
        flush.fill_(i)
        torch.cuda.synchronize()
        print(f"--- batch {batch} N {N} K {K} call {i}", flush=True)
        y = F.gemm_4bit(x, q.t(), st)
        torch.cuda.synchronize()
