import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch
from bnb_b200 import functional as F
torch.manual_seed(0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, K = 14336, 4096
W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
x = torch.randn(batch, K, device="cuda").to(torch.bfloat16)
q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(6):
    flush.fill_(i)
    y = F.gemm_4bit(x, q.t(), st)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
