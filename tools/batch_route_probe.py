"""Batch 2..8: GEMV kernels with the batch in the MMA n dimension (current route) vs the small-batch tcgen05 GEMM
(forced by passing a strided-looking output spec through peer_outs=[]... here: by passing a zero bias)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch
from bnb_b200 import functional as F
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kbench import time_graph

for (N, K) in [(4096, 4096), (11008, 4096), (14336, 4096), (8192, 8192), (28672, 8192), (8192, 28672)]:
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    del W
    nbuf = max(2, int(200e6 // (N * K // 2)) + 1)
    qs = [q.clone() for _ in range(nbuf)]
    zero_bias = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    for batch in (2, 4, 8):
        x = torch.randn(batch, K, device="cuda").bfloat16()
        outs = [torch.empty(batch, N, dtype=torch.bfloat16, device="cuda") for _ in range(nbuf)]
        F.gemm_4bit(x, qs[0].t(), st, out=outs[0]); F.gemm_4bit(x, qs[0].t(), st, bias=zero_bias, out=outs[0])
        a = time_graph([(lambda i=i: F.gemm_4bit(x, qs[i].t(), st, out=outs[i])) for i in range(nbuf)], 20)
        b = time_graph([(lambda i=i: F.gemm_4bit(x, qs[i].t(), st, bias=zero_bias, out=outs[i])) for i in range(nbuf)], 20)
        print(json.dumps({"shape": [N, K], "batch": batch, "gemv_batch_route_us": round(a, 2), "tcgen05_small_us": round(b, 2)}), flush=True)
