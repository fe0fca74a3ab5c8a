import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch
from bnb_b200 import functional as F
torch.manual_seed(0)
batch, N, K = 16, 14336, 4096
dt = torch.bfloat16
W = (torch.randn(N, K, device="cuda") * 0.02).to(dt)
x = torch.randn(batch, K, device="cuda").to(dt)
q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
Wd = F.dequantize_4bit(q, st).to(dt).double()
ref = x.double() @ Wd.t()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
thr = 0.05 * ref.abs().mean()
# per-stage contributions to identify which k-block is missing / duplicated
for it in range(90):
    flush.fill_(it & 255)
    if it % 3 == 0:
        torch.cuda.synchronize()
    y = F.gemm_4bit(x, q.t(), st).double()
    d = y - ref
    bad = d.abs() > thr
    if int(bad.sum()):
        rows = bad.any(0).nonzero().flatten()
        cols = bad.any(1).nonzero().flatten()
        tile = int(rows[0]) // 128
        print(f"launch {it}: {int(bad.sum())} bad; rows {int(rows.min())}..{int(rows.max())} (tile {tile}, {len(rows)} rows), batch cols {cols.tolist()}")
        r0 = tile * 128
        xs = x.double().view(batch, K // 64, 64)                       # [b, s, 64]
        am_blocks = st.absmax.view(N, K // 64).double()
        for rr in rows.tolist()[:6]:
            wrow = Wd[rr].view(K // 64, 64)                              # [s, 64]
            contrib = torch.einsum("bsk,sk->sb", xs, wrow)               # [s, b]
            dv = d[:, rr]                                                # [b]
            coef = (contrib @ dv) / (contrib * contrib).sum(1)
            resid = (dv[None, :] - coef[:, None] * contrib).norm(dim=1) / dv.norm()
            s1 = int(resid.argmin())
            # model 2: stage s multiplied x_s by the unscaled/scaled weights of stage s2 of the same row
            cross = torch.einsum("bsk,tk->stb", xs, wrow)                # [s, t, b]: x_s . W_t
            m2 = cross - contrib[:, None, :]                             # error if stage s used W_t
            r2 = (dv[None, None, :] - m2).norm(dim=2) / dv.norm()
            s2 = int(r2.argmin()); sa, sb = s2 // (K // 64), s2 % (K // 64)
            # model 3: absmax of another block (same row) applied to stage s: error = (am_t/am_s - 1) contrib_s
            ratio = am_blocks[rr][None, :] / am_blocks[rr][:, None] - 1.0   # [s, t]
            m3 = ratio[:, :, None] * contrib[:, None, :]
            r3 = (dv[None, None, :] - m3).norm(dim=2) / dv.norm()
            s3 = int(r3.argmin()); ta, tb = s3 // (K // 64), s3 % (K // 64)
            print(f"   row {rr} (tile row {rr - r0}): |d|/|y| {float(dv.norm() / ref[:, rr].norm()):.3f}; scale-one-stage: s={s1} coef={float(coef[s1]):.3f} resid={float(resid[s1]):.3f};"
                  f" W of other stage: s={sa} used {sb} resid={float(r2.view(-1)[s2]):.3f}; absmax of other block: s={ta} used {tb} resid={float(r3.view(-1)[s3]):.3f}")
print("done")
