import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch, bnb_b200
m, k, n = 4096, 4096, 16384
torch.manual_seed(0)
A16 = torch.randn(m, k, device="cuda").half()
A16[:, [7, 100, 2000, 3000]] = 8.0
lin = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
with torch.no_grad():
    for _ in range(4):
        y = lin(A16)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
