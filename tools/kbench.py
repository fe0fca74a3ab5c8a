#!/usr/bin/env python
"""Per-kernel micro-benchmarks (CUDA events, rotating buffers larger than L2) for every kernel of the hot
path.  Prints one JSON line per kernel with the algorithmic bytes/ops, time and roofline fraction.

    python tools/kbench.py [--only quant,dequant,gemv,int8] [--iters 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import bnb_b200  # noqa: E402
from bnb_b200 import functional as F  # noqa: E402

PEAK_HBM = 6449.7
PEAK_BF16 = 1551.4
try:
    _p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_HBM, PEAK_BF16 = float(_p["hbm_gbs"]), float(_p["bf16_tflops"])
except Exception:
    pass


def time_graph(fns, iters):
    """fns: list of callables over DISTINCT buffers; captured into one CUDA graph; returns us per call."""
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * len(fns))


def report(name, us, nbytes=None, ops=None, **extra):
    line = {"kernel": name, "us": round(us, 3)}
    if nbytes is not None:
        gbs = nbytes / us / 1e3
        line.update(algorithmic_bytes=nbytes, GBps=round(gbs, 1), hbm_frac=round(gbs / PEAK_HBM, 4))
    if ops is not None:
        tops = ops / us / 1e6
        line.update(ops=ops, TOPS=round(tops, 1), frac_of_2x_bf16_measured=round(tops / (2 * PEAK_BF16), 4),
                    frac_of_4500_nominal=round(tops / 4500.0, 4))
    line.update(extra)
    print(json.dumps(line), flush=True)


def bench_quant(iters):
    n = 4096 * 4096
    for dt, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        nbuf = 8
        src = [torch.randn(n, device="cuda").to(dt) for _ in range(nbuf)]
        outs = [torch.empty((n // 2, 1), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
        ams = [torch.empty(n // 64, dtype=torch.float32, device="cuda") for _ in range(nbuf)]
        for qt in ("nf4", "fp4"):
            fns = [(lambda i=i: F.quantize_4bit(src[i], absmax=ams[i], out=outs[i], blocksize=64, quant_type=qt)) for i in range(nbuf)]
            us = time_graph(fns, iters)
            report(f"quantize_{qt}_{name}_bs64", us, n * src[0].element_size() + n // 2 + n // 64 * 4)
        code = F.create_dynamic_map().cuda()
        outs8 = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
        fns = [(lambda i=i: F.quantize_blockwise(src[i], code=code, absmax=ams[i], out=outs8[i], blocksize=64)) for i in range(nbuf)]
        report(f"quantize_8bit_{name}_bs64", time_graph(fns, iters), n * src[0].element_size() + n + n // 64 * 4)
        # dequantize
        q, st = F.quantize_4bit(src[0], blocksize=64, quant_type="nf4")
        qs = [q.clone() for _ in range(nbuf)]
        for odt, oname in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            dst = [torch.empty(n, dtype=odt, device="cuda") for _ in range(nbuf)]
            fns = [(lambda i=i: F.dequantize_4bit(qs[i], absmax=ams[i], out=dst[i], blocksize=64, quant_type="nf4")) for i in range(nbuf)]
            report(f"dequantize_nf4_to_{oname}_bs64", time_graph(fns, iters), n // 2 + n // 64 * 4 + n * dst[0].element_size())
            del dst
        del src, outs, ams, outs8, qs
        torch.cuda.empty_cache()


def gemv_bytes(N, K, nested=True):
    nb = N * K // 64
    return N * K // 2 + (nb + 4 * ((nb + 255) // 256) if nested else 4 * nb) + 2 * K + 2 * N + 1088


def bench_gemv(iters, shapes=None, dtypes=(torch.bfloat16,)):
    shapes = shapes or [(4096, 4096), (11008, 4096), (4096, 11008), (8192, 8192), (28672, 8192), (8192, 28672)]
    for dt in dtypes:
        for (N, K) in shapes:
            nbuf = max(2, min(24, int(300e6 // (N * K // 2)) + 1))
            if os.environ.get("KBENCH_NBUF"):      # e.g. 2: every launch finds its weights in L2 (upper bound of any prefetch scheme)
                nbuf = int(os.environ["KBENCH_NBUF"])
            packs = []
            W = (torch.randn(N, K, device="cuda") * 0.02).to(dt)
            q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
            del W
            for i in range(nbuf):
                packs.append(q.clone())
            x = torch.randn(1, K, device="cuda").to(dt)
            outs = [torch.empty(1, N, dtype=dt, device="cuda") for _ in range(nbuf)]
            fns = [(lambda i=i: F.gemv_4bit(x, packs[i].t(), out=outs[i], state=st)) for i in range(nbuf)]
            us = time_graph(fns, iters)
            extra = {}
            if os.environ.get("BNB_B200_GEMV_PROBE") == "1":
                import ctypes as ct
                buf = (ct.c_ulonglong * 12)()
                F.lib.cbnb_debug_gemv_probe(buf)
                if buf[1]:
                    extra = {"cta_us": round(buf[1] / 1e3, 2), "sm_mhz_in_kernel": round(buf[0] * 1e3 / buf[1], 1),
                             "phase_us": {"loads_issued": buf[7] / 1e3, "code2_in": buf[8] / 1e3, "lut_done": buf[9] / 1e3, "tables": buf[2] / 1e3, "prev_done": buf[3] / 1e3, "x_ready": buf[4] / 1e3,
                                          "warp0_done": buf[5] / 1e3, "all_done": buf[6] / 1e3}}
            report(f"gemv_nf4_nested_{str(dt).split('.')[-1]}_{N}x{K}", us, gemv_bytes(N, K), buffers=nbuf, **extra)
            del packs, outs
            torch.cuda.empty_cache()


def bench_gemm4(iters):
    """K4: fused NF4 GEMM (tcgen05) vs the reference's composition (dequantize_4bit + F.linear), BASELINE config 4."""
    for (N, K) in [(14336, 4096), (4096, 14336)]:
        W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
        del W
        nbuf = 6                                   # 6 x 29 MB of packed weights + absmax > L2
        qs = [q.clone() for _ in range(nbuf)]
        for batch in (16, 32, 64, 128, 256):
            x = torch.randn(batch, K, device="cuda").bfloat16()
            outs = [torch.empty(batch, N, dtype=torch.bfloat16, device="cuda") for _ in range(nbuf)]
            F.gemm_4bit(x, qs[0].t(), st, out=outs[0])   # allocates the split-K workspace outside the capture
            fns = [(lambda i=i: F.gemm_4bit(x, qs[i].t(), st, out=outs[i])) for i in range(nbuf)]
            us = time_graph(fns, iters)
            nbytes = N * K // 2 + 4 * N * K // 64 + 2 * batch * (K + N)
            flops = 2.0 * batch * N * K
            fns = [(lambda i=i: torch.nn.functional.linear(x, F.dequantize_4bit(qs[i], st))) for i in range(nbuf)]
            us_ref = time_graph(fns, max(2, iters // 2))
            print(json.dumps({"kernel": f"gemm4_nf4_bf16_{N}x{K}_b{batch}", "us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1),
                              "hbm_frac": round(nbytes / us / 1e3 / PEAK_HBM, 4), "TFLOPs": round(flops / us / 1e6, 1),
                              "bf16_frac": round(flops / us / 1e6 / PEAK_BF16, 4),
                              "reference_composition_us": round(us_ref, 2), "speedup_vs_composition": round(us_ref / us, 2)}), flush=True)
            del outs


def bench_int8(iters):
    m, k, n = 4096, 4096, 16384
    A = torch.randn(m, k, device="cuda").half()
    A[:, [7, 100, 2000, 3000]] = 8.0
    fns = [lambda: F.get_colrow_absmax(A, threshold=6.0)]
    report("get_col_row_stats_4096x4096_thr6", time_graph(fns, iters), m * k * 2)
    rs, cs, nnz = F.get_colrow_absmax(A, threshold=0.0)
    oc = torch.empty(m, k, dtype=torch.int8, device="cuda")
    orow = torch.empty(m, k, dtype=torch.int8, device="cuda")
    fns = [lambda: F.double_quant(A, col_stats=cs, row_stats=rs, out_col=oc, out_row=orow)]
    report("double_rowcol_quant_4096x4096", time_graph(fns, iters), m * k * 4)
    CA = torch.randint(-127, 128, (m, k), dtype=torch.int8, device="cuda")
    CB = torch.randint(-127, 128, (n, k), dtype=torch.int8, device="cuda")
    out32 = torch.empty(m, n, dtype=torch.int32, device="cuda")
    fns = [lambda: F.igemmlt(CA, CB, ((m, k), "row"), ((n, k), "row"), out=out32, Sout=((m, n), "row"))]
    report("igemm_rowmajor_int32_4096x16384x4096", time_graph(fns, iters), ops=2.0 * m * n * k)
    SCA = torch.rand(m, device="cuda") + 0.5
    SCB = torch.rand(n, device="cuda") + 0.5
    bias = torch.randn(n, device="cuda").half()
    out16 = torch.empty(m, n, dtype=torch.float16, device="cuda")
    fns = [lambda: F.int8_linear_dequant(CA, CB, SCA, SCB, bias=bias, out=out16)]
    report("igemm_rowmajor_fused_dequant_fp16_4096x16384x4096", time_graph(fns, iters), ops=2.0 * m * n * k)
    # reference-shaped path: col32 / col_turing operands through the ABI wrappers + mm_dequant
    C32A, SA = F.transform(CA, "col32")
    CxB, SB = F.transform(CB, "col_turing")
    fns = [lambda: F.transform(CA, "col32")]
    report("transform_row2col32_4096x4096", time_graph(fns, iters), 2 * m * k)
    o32, So = F.igemmlt(C32A, CxB, SA, SB)
    fns = [lambda: F.mm_dequant(o32, So, SCA, SCB, bias=bias, out=out16)]
    report("dequant_mm_int32_fp16_4096x16384", time_graph(fns, iters), m * n * 6)
    fns = [lambda: F.igemmlt(C32A, CxB, SA, SB, out=o32, Sout=So)]
    report("cigemmlt_turing_32_abi_4096x16384x4096", time_graph(fns, max(2, iters // 4)), ops=2.0 * m * n * k)
    # end-to-end Linear8bitLt forward (threshold 6.0)
    lin = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
    with torch.no_grad():
        lin(A)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lin(A)
        e1.record()
        torch.cuda.synchronize()
    report("Linear8bitLt_forward_thr6_4096tok_4096x16384", e0.elapsed_time(e1) * 1e3 / 5, ops=2.0 * m * n * k)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="quant,gemv,gemm4,int8")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    torch.manual_seed(0)
    which = a.only.split(",")
    if "quant" in which:
        bench_quant(a.iters)
    if "gemv" in which:
        bench_gemv(a.iters)
    if "gemm4" in which:
        bench_gemm4(a.iters)
    if "int8" in which:
        bench_int8(a.iters)
