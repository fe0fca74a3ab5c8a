import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch
from bnb_b200 import functional as F
m, k = 4096, 4096
torch.manual_seed(0)
A = torch.randn(m, k, device="cuda").half()
A[:, [7, 100, 2000, 3000]] = 8.0
for _ in range(3):
    rs, cs, nnz = F.get_colrow_absmax(A, threshold=6.0)
    out = F.double_quant(A, threshold=6.0)
    W = torch.randn(m, k, device="cuda")
    q, st = F.quantize_4bit(W, blocksize=64, quant_type="nf4")
    q2, st2 = F.quantize_4bit(W.bfloat16(), blocksize=64, quant_type="nf4")
    d = F.dequantize_4bit(q2, st2)
torch.cuda.synchronize()
print("ok")
