import sys, os, ctypes as ct
ROOT='/root/repo'
sys.path[:0]=[ROOT, os.path.join(ROOT,'bitsandbytes-sycl_b200')]
import torch
from bnb_b200 import functional as F
torch.manual_seed(0)
def run(N,K,n,dt=torch.bfloat16):
    W=(torch.randn(N,K,device='cuda')*0.02).to(dt)
    q,st=F.quantize_4bit(W,blocksize=64,compress_statistics=True,quant_type='nf4')
    x=torch.randn(n,K,device='cuda').to(dt)
    out=torch.empty(n,N,device='cuda',dtype=dt)
    s2=st.state2; code=st.code; off=float(st.offset)
    def call():
        prev=F.pre_call(x.device)
        getattr(F.lib,'cgemm_4bit_inference_nested_bf16')(ct.c_int32(N),ct.c_int32(n),ct.c_int32(K),F.get_ptr(x),F.get_ptr(q),F.get_ptr(st.absmax),F.get_ptr(s2.absmax),F.get_ptr(s2.code),ct.c_float(off),F.get_ptr(code),F.get_ptr(out),ct.c_int32(N),ct.c_int32(K//2),ct.c_int32(N),ct.c_int32(64),ct.c_int32(s2.blocksize))
        F.post_call(prev)
    call(); torch.cuda.synchronize()
    import copy
    st32=copy.copy(st); st32.dtype=torch.float32
    ref=x.double()@F.dequantize_4bit(q,st32).double().t()
    err=float((out.double()-ref).norm()/ref.norm())
    # timing
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): call()
    g.replay(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)*1e3/50
    # K4 for comparison
    y=F.gemm_4bit(x,q.t(),st); torch.cuda.synchronize()
    g2=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for _ in range(10): F.gemm_4bit(x,q.t(),st)
    g2.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): g2.replay()
    e1.record(); torch.cuda.synchronize()
    us2=e0.elapsed_time(e1)*1e3/50
    print(f"N={N} K={K} n={n}: rel-L2 {err:.2e}  nested-gemv-batch {us:.1f} us   K4 {us2:.1f} us", flush=True)
for (N,K) in [(14336,4096),(4096,14336)]:
    for n in (2,5,8):
        run(N,K,n)
