"""One launch each of the int8 GEMM (config 3) and the wide fused 4-bit GEMM (batch 256), for `ncu -k regex:...` captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch
from bnb_b200 import functional as F
which = sys.argv[1] if len(sys.argv) > 1 else "igemm"
torch.manual_seed(0)
if which == "igemm":
    m, k, n = 4096, 4096, 16384
    CA = torch.randint(-127, 128, (m, k), dtype=torch.int8, device="cuda")
    CB = torch.randint(-127, 128, (n, k), dtype=torch.int8, device="cuda")
    out32 = torch.empty(m, n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        F.igemmlt(CA, CB, ((m, k), "row"), ((n, k), "row"), out=out32, Sout=((m, n), "row"))
else:
    N, K, batch = 14336, 4096, 256
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(batch, K, device="cuda").bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(3):
        flush.fill_(i)
        y = F.gemm_4bit(x, q.t(), st)
torch.cuda.synchronize()
print("ok")
