"""GPU parity: 4-bit GEMV (K3) through the C-ABI vs the oracle.
Floating point -> tolerance gates, stated here (derived from the three rounding chains):
  bf16:  rel-L2 <= 2.5e-3 vs the fp64-exact product, and <= 5e-3 vs the reference-faithful T-arithmetic chain
  fp16:  rel-L2 <= 6e-4 vs exact;   fp32: rel-L2 <= 2e-6 vs exact
and the kernel must be no less accurate than the reference chain (w.r.t. exact)."""
import numpy as np
import pytest
import torch

from helpers import DT, bits_equal, from_bits, rel_l2, to_bits
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL_EXACT = {"bf16": 2.5e-3, "fp16": 6e-4, "fp32": 2e-6}
TOL_FAITHFUL = {"bf16": 5e-3, "fp16": 1.5e-3, "fp32": 2e-6}


@pytest.fixture(scope="module")
def F():
    assert torch.cuda.is_available()
    from bnb_b200 import functional
    return functional


def make_case(F, N, K, dtype, qtype="nf4", nested=True, blocksize=64, seed=0):
    torch.manual_seed(seed)
    W = (torch.randn(N, K) * 0.02).to(DT[dtype])
    x = torch.randn(1, K).to(DT[dtype])
    q, st = F.quantize_4bit(W.cuda(), blocksize=blocksize, compress_statistics=nested, quant_type=qtype)
    return W, x, q, st


def oracle_outputs(F, x, q, st, N, K, dtype):
    qn = q.cpu().numpy().ravel()
    if st.nested:
        absmax = orc.denest_absmax(st.absmax.cpu().numpy(), st.state2.absmax.cpu().numpy(),
                                   st.state2.code.cpu().numpy(), np.float32(st.offset.item()), st.state2.blocksize)
    else:
        absmax = st.absmax.cpu().numpy()
    code = st.code.cpu().numpy()
    xb = to_bits(x).ravel()
    exact = orc.gemm_4bit_exact(xb, dtype, qn, absmax, code, 1, N, K, st.blocksize)[0]
    faithful = orc.gemv_4bit(xb, dtype, qn, absmax, code, N, K, st.blocksize, 0)
    faithful = from_bits(faithful, dtype).double().numpy()
    return exact, faithful


@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("shape", [(4096, 4096), (11008, 4096), (4096, 11008)])
def test_gemv_config2_shapes(F, dtype, shape):
    """BASELINE config 2: Llama-2-7B shapes, NF4, blocksize 64, double-quantised absmax, batch 1."""
    N, K = shape
    W, x, q, st = make_case(F, N, K, dtype)
    y = F.gemv_4bit(x.cuda(), q.t(), state=st)
    assert y.shape == (1, N) and y.dtype == DT[dtype]
    exact, faithful = oracle_outputs(F, x, q, st, N, K, dtype)
    yk = y.double().cpu().numpy().ravel()
    err_exact = rel_l2(yk, exact)
    err_ref_chain = rel_l2(faithful, exact)
    assert err_exact <= TOL_EXACT[dtype], (err_exact,)
    assert rel_l2(yk, faithful) <= TOL_FAITHFUL[dtype]
    assert err_exact <= err_ref_chain * 1.05 + 1e-7, (err_exact, err_ref_chain)   # no less accurate than the reference
    # elementwise: output rounding (half an ulp of T, relative) + the accumulated code/x rounding noise, which
    # scales with the rms of the outputs (~5 sigma)
    rms = np.sqrt(np.mean(exact ** 2))
    rel, noise = {"bf16": (2.0 ** -8, 2.0 ** -7), "fp16": (2.0 ** -11, 2.0 ** -9), "fp32": (1e-5, 1e-5)}[dtype]
    assert np.all(np.abs(yk - exact) <= rel * np.abs(exact) + noise * rms)


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_fused_nested_equals_denested_path(F, dtype):
    """One-launch nested GEMV == de-nest (2 launches) + GEMV: same arithmetic, bit-identical outputs."""
    N, K = 1024, 2048
    W, x, q, st = make_case(F, N, K, dtype, seed=3)
    y_fused = F.gemv_4bit(x.cuda(), q.t(), state=st)
    old = F.FUSED_NESTED_GEMV
    try:
        F.FUSED_NESTED_GEMV = False
        y_ref_path = F.gemv_4bit(x.cuda(), q.t(), state=st)
    finally:
        F.FUSED_NESTED_GEMV = old
    assert torch.equal(y_fused.view(torch.int16), y_ref_path.view(torch.int16))


@pytest.mark.parametrize("qtype", ["nf4", "fp4"])
@pytest.mark.parametrize("blocksize", [64, 128, 512])
@pytest.mark.parametrize("nested", [False, True])
def test_gemv_variants(F, qtype, blocksize, nested):
    N, K = 520, 1536                               # N not a multiple of 16, K = 24 * 64
    W, x, q, st = make_case(F, N, K, "bf16", qtype, nested, blocksize, seed=7)
    y = F.gemv_4bit(x.cuda(), q.t(), state=st)
    exact, faithful = oracle_outputs(F, x, q, st, N, K, "bf16")
    assert rel_l2(y.double().cpu().numpy(), exact) <= TOL_EXACT["bf16"]


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("nested", [False, True])
@pytest.mark.parametrize("shape", [(16, 256), (5, 512), (300, 768), (33, 1280), (2064, 8192), (1024, 28672)])
def test_gemv_block_column_edges(F, dtype, nested, shape):
    """Block-column kernel edge cases: half chunks (K % 512 == 256), N below / not a multiple of the 16-row tile,
    fewer tiles than SMs, many chunks per row."""
    N, K = shape
    W, x, q, st = make_case(F, N, K, dtype, "nf4", nested, 64, seed=N + K)
    y = F.gemv_4bit(x.cuda(), q.t(), state=st)
    exact, faithful = oracle_outputs(F, x, q, st, N, K, dtype)
    yk = y.double().cpu().numpy().ravel()
    assert np.all(np.isfinite(yk))
    if N >= 256:          # rel-L2 is a statistical gate: meaningless for a handful of outputs
        assert rel_l2(yk, exact) <= TOL_EXACT[dtype]
    rms = np.sqrt(np.mean(exact ** 2))
    rel, noise = {"bf16": (2.0 ** -8, 2.0 ** -7), "fp16": (2.0 ** -11, 2.0 ** -9)}[dtype]
    assert np.all(np.abs(yk - exact) <= rel * np.abs(exact) + noise * rms)


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(28672, 8192), (8192, 28672), (8192, 8192), (1024, 8192)])
def test_gemv_config5_shapes_row_sample(F, dtype, shape):
    """BASELINE config 5 (Llama-3-70B linears) at FULL size on the device; the oracle checks a sample of 16-row
    tiles (first, last, around every 1/148th boundary where a CTA's tile range ends, and random ones) against the
    fp64-exact product and the reference-faithful chain -- same gates as config 2."""
    N, K = shape
    torch.manual_seed(N + K)
    W = (torch.randn(N, K, device="cuda") * 0.02).to(DT[dtype])
    x = torch.randn(1, K).to(DT[dtype])
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    del W
    y = F.gemv_4bit(x.cuda(), q.t(), state=st)
    torch.cuda.synchronize()
    tiles = N // 16
    rng = np.random.RandomState(N)
    pick = {0, 1, tiles - 1, tiles - 2}
    for b in (1, 37, 74, 111, 147):
        pick.update({min(tiles - 1, b * tiles // 148), max(0, b * tiles // 148 - 1)})
    pick.update(rng.randint(0, tiles, 16).tolist())
    rows = np.concatenate([np.arange(tl * 16, tl * 16 + 16) for tl in sorted(pick)])
    absmax = orc.denest_absmax(st.absmax.cpu().numpy(), st.state2.absmax.cpu().numpy(), st.state2.code.cpu().numpy(),
                               np.float32(st.offset.item()), st.state2.blocksize).reshape(N, K // 64)[rows].ravel()
    qn = q.cpu().numpy().reshape(N, K // 2)[rows].ravel()
    code = st.code.cpu().numpy()
    xb = to_bits(x).ravel()
    exact = orc.gemm_4bit_exact(xb, dtype, qn, absmax, code, 1, len(rows), K, 64)[0]
    faithful = from_bits(orc.gemv_4bit(xb, dtype, qn, absmax, code, len(rows), K, 64, 0), dtype).double().numpy()
    yk = y.double().cpu().numpy().ravel()[rows]
    err_exact, err_ref_chain = rel_l2(yk, exact), rel_l2(faithful, exact)
    assert err_exact <= TOL_EXACT[dtype], (err_exact,)
    assert rel_l2(yk, faithful) <= TOL_FAITHFUL[dtype]
    assert err_exact <= err_ref_chain * 1.05 + 1e-7, (err_exact, err_ref_chain)
    rms = np.sqrt(np.mean(exact ** 2))
    rel, noise = {"bf16": (2.0 ** -8, 2.0 ** -7), "fp16": (2.0 ** -11, 2.0 ** -9)}[dtype]
    assert np.all(np.abs(yk - exact) <= rel * np.abs(exact) + noise * rms)


def test_gemv_deterministic(F):
    """Partial sums meet in a fixed order: two launches give bit-identical outputs."""
    W, x, q, st = make_case(F, 4096, 4096, "bf16", seed=5)
    xc = x.cuda()
    y1 = F.gemv_4bit(xc, q.t(), state=st)
    y2 = F.gemv_4bit(xc, q.t(), state=st)
    assert torch.equal(y1.view(torch.int16), y2.view(torch.int16))


def test_gemv_3d_input_and_module_dispatch(F):
    from bnb_b200.nn import LinearNF4
    torch.manual_seed(11)
    lin = LinearNF4(1024, 768, bias=True, compute_dtype=torch.bfloat16)
    ref_w = lin.weight.data.clone()
    ref_b = lin.bias.data.clone()
    lin = lin.cuda()
    assert lin.weight.dtype == torch.uint8 and lin.weight.quant_state.nested
    x = torch.randn(1, 1, 1024, dtype=torch.bfloat16, device="cuda")
    with torch.no_grad():
        y = lin(x)                                   # batch 1 -> gemv path
    assert y.shape == (1, 1, 768)
    Wd = F.dequantize_4bit(lin.weight.data, lin.weight.quant_state).float()
    y_ref = x.float().reshape(1, -1) @ Wd.t() + ref_b.cuda().float()
    assert rel_l2(y.float().cpu().numpy(), y_ref.cpu().numpy()) < 5e-3
    xb = torch.randn(4, 7, 1024, dtype=torch.bfloat16, device="cuda")
    with torch.no_grad():
        yb = lin(xb)                                 # batch > 1 -> MatMul4Bit
    yb_ref = xb.float() @ Wd.t() + ref_b.cuda().float()
    assert yb.shape == (4, 7, 768)
    assert rel_l2(yb.float().cpu().numpy(), yb_ref.cpu().numpy()) < 5e-3
    # the reference's own test bar (tests_pvc/autograd.py:389-391): mean |out_bnb - out_torch| < 0.115
    y_fp = xb.float() @ ref_w.cuda().float().t() + ref_b.cuda().float()
    assert (yb.float() - y_fp).abs().mean().item() < 0.115


def test_gemv_generic_path_odd_shapes(F):
    """K not a multiple of 64 / fp32: the generic CUDA path; tail semantics of kernel_gemm.cpp:1312-1366."""
    import ctypes as ct
    torch.manual_seed(2)
    N, K, bs = 37, 96, 32
    W = torch.randn(N, K) * 0.05
    x = torch.randn(K)
    q_ref, am_ref = orc.quantize_blockwise(W.numpy().ravel(), "fp32", None, bs, "nf4")
    code = orc.nf4_table()
    out = torch.zeros(N, device="cuda")
    qd, amd, cd, xd = (torch.from_numpy(a).cuda() for a in (q_ref, am_ref, code, x.numpy()))
    F.lib.cbnb_set_stream(ct.c_void_p(torch.cuda.current_stream().cuda_stream))
    F.lib.cgemm_4bit_inference_naive_fp32(ct.c_int(N), ct.c_int(1), ct.c_int(K), ct.c_void_p(xd.data_ptr()),
                                          ct.c_void_p(qd.data_ptr()), ct.c_void_p(amd.data_ptr()),
                                          ct.c_void_p(cd.data_ptr()), ct.c_void_p(out.data_ptr()), ct.c_int(N),
                                          ct.c_int(K // 2), ct.c_int(N), ct.c_int(bs))
    torch.cuda.synchronize()
    assert F.lib.cbnb_last_error() == 0
    exact = orc.gemm_4bit_exact(x.numpy(), "fp32", q_ref, am_ref, code, 1, N, K, bs)[0]
    assert rel_l2(out.cpu().numpy(), exact) < 2e-6


def test_gemv_host_tables_hint_is_bit_identical(F):
    """cbnb_set_gemv_host_tables (NF4 lookup table from immediates) must not change a single output bit; a table that
    is NOT NF4 (FP4) must ignore the hint's fast path and still be exact."""
    import ctypes as ct
    for qtype in ("nf4", "fp4"):
        W, x, q, st = make_case(F, 1024, 4096, "bf16", qtype=qtype, seed=21)
        y_hint = F.gemv_4bit(x.cuda(), q.t(), state=st)               # the Python mirror passes the host tables
        assert getattr(st, "_tables_host", None) is not None
        saved, st._tables_host = st._tables_host, (None, None)
        y_plain = F.gemv_4bit(x.cuda(), q.t(), state=st)              # device-pointer tables only
        st._tables_host = saved
        assert torch.equal(y_hint.view(torch.int16), y_plain.view(torch.int16)), qtype


@pytest.mark.parametrize("impl", ["m", "t"])
def test_experimental_gemv_kernels_stay_parity_green(impl):
    """The TMEM-staged (BNB_B200_GEMV_IMPL=m) and the first-session MMA (=t) kernels are selectable experiments
    (DESIGN.md K3); they must keep producing the block-column kernel's accuracy.  Run in a subprocess: the selection
    is read once per process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch, numpy as np\n"
        f"sys.path[:0] = [{root!r}, {os.path.join(root, 'bitsandbytes-sycl_b200')!r}, {os.path.join(root, 'tests')!r}]\n"
        "from bnb_b200 import functional as F\n"
        "torch.manual_seed(5)\n"
        "for (N, K) in [(4096, 4096), (1024, 11008), (176, 8192)]:\n"
        "    W = (torch.randn(N, K) * 0.02).bfloat16().cuda()\n"
        "    x = torch.randn(1, K).bfloat16().cuda()\n"
        "    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type='nf4')\n"
        "    y = F.gemv_4bit(x, q.t(), state=st).double()\n"
        "    import copy; st32 = copy.copy(st); st32.dtype = torch.float32   # exact weights: code * absmax in fp32\n"
        "    ref = x.double() @ F.dequantize_4bit(q, st32).double().t()\n"
        "    err = float((y - ref).norm() / ref.norm())\n"
        "    assert err < 2.5e-3, (N, K, err)\n"
        "print('ok')\n")
    env = dict(os.environ, BNB_B200_GEMV_IMPL=impl)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("Ns,K", [((4096, 4096, 4096), 4096), ((11008, 11008), 4096), ((100, 4096, 24, 1000), 1024), ((512,), 11008)])
def test_gemv_multi_is_bit_identical_to_single_calls(F, dtype, Ns, K):
    """ADDITIVE cgemm_4bit_inference_nested_multi_*: the matrices of a decoder layer that share x (q/k/v, gate/up) in one
    launch.  Same arithmetic per output element -> every output bit equals the single-matrix call's."""
    torch.manual_seed(len(Ns) * 1000 + K)
    x = torch.randn(1, K).to(DT[dtype]).cuda()
    qs, sts = [], []
    for i, N in enumerate(Ns):
        W = (torch.randn(N, K) * (0.02 + 0.01 * i)).to(DT[dtype]).cuda()
        q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
        qs.append(q)
        sts.append(st)
    singles = [F.gemv_4bit(x, q.t(), state=st) for q, st in zip(qs, sts)]
    multi = F.gemv_4bit_multi(x, [q.t() for q in qs], sts)
    torch.cuda.synchronize()
    for a, b, N in zip(singles, multi, Ns):
        assert b.shape == (1, N)
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), N


def test_matmul_4bit_multi_matches_linear4bit_modules(F):
    """bnb_b200.matmul_4bit_multi over three LinearNF4 modules (q/k/v style, with bias) == the modules' own forward."""
    import bnb_b200
    torch.manual_seed(31)
    k = 1024
    lins = [bnb_b200.nn.LinearNF4(k, n, bias=True, compute_dtype=torch.bfloat16).cuda() for n in (1024, 256, 256)]
    x = torch.randn(1, k, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        ref = [lin(x) for lin in lins]
        got = bnb_b200.matmul_4bit_multi(x, [lin.weight.t() for lin in lins], [lin.weight.quant_state for lin in lins],
                                         biases=[lin.bias.to(torch.bfloat16) for lin in lins])
        x8 = torch.randn(8, k, device="cuda", dtype=torch.bfloat16)          # batch > 1: per-weight fallback
        ref8 = [lin(x8) for lin in lins]
        got8 = bnb_b200.matmul_4bit_multi(x8, [lin.weight.t() for lin in lins], [lin.weight.quant_state for lin in lins],
                                          biases=[lin.bias.to(torch.bfloat16) for lin in lins])
    for a, b in zip(ref + ref8, got + got8):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_peer_store_epilogue_and_barrier_emulated_on_one_gpu(F, dtype):
    """The multi-GPU data plane on ONE device: `peer_outs` are ordinary device buffers standing in for the peer-mapped
    copies of the output vector (the kernel cannot tell), so every store of the fused all-gather epilogue
    (cgemm_4bit_inference_nested_push_* and its multi-matrix form) is checked bit for bit; cbnb_peer_barrier runs with
    its "peer" slots pointing at its own signal words (each rank publishes to itself), which exercises the sequence
    counting and the bounded wait without a second process (B200_PROFILING.md: never spin on another launch)."""
    import ctypes as ct
    N, K = 1024, 4096
    W, x, q, st = make_case(F, N, K, dtype, seed=91)
    xc = x.cuda()
    ref = F.gemv_4bit(xc, q.t(), state=st)
    npeers = 3
    peers = [torch.zeros(1, N, dtype=DT[dtype], device="cuda") for _ in range(npeers)]
    out = torch.zeros(1, N, dtype=DT[dtype], device="cuda")
    F.gemv_4bit(xc, q.t(), out=out, state=st, peer_outs=[p.data_ptr() for p in peers])
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16))
    for p in peers:
        assert torch.equal(p.view(torch.int16), ref.view(torch.int16))
    # several matrices that share x, one launch, peer copies of every output slice
    W2, _, q2, st2 = make_case(F, 512, K, dtype, seed=92)
    ref2 = F.gemv_4bit(xc, q2.t(), state=st2)
    outs = [torch.zeros(1, N, dtype=DT[dtype], device="cuda"), torch.zeros(1, 512, dtype=DT[dtype], device="cuda")]
    pm = [[torch.zeros(1, n_, dtype=DT[dtype], device="cuda") for _ in range(2)] for n_ in (N, 512)]
    F.gemv_4bit_multi(xc, [q.t(), q2.t()], [st, st2], outs=outs, peer_outs=[[p.data_ptr() for p in row] for row in pm])
    torch.cuda.synchronize()
    for got, want, row in ((outs[0], ref, pm[0]), (outs[1], ref2, pm[1])):
        assert torch.equal(got.view(torch.int16), want.view(torch.int16))
        for p in row:
            assert torch.equal(p.view(torch.int16), want.view(torch.int16))
    # the PDL-chained barrier: two "peers" whose slots are this rank's own slots
    sig = torch.zeros(8, dtype=torch.int32, device="cuda")
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    arr = (ct.c_void_p * 2)(sig.data_ptr(), sig.data_ptr() + 4)
    F.lib.cbnb_set_stream(ct.c_void_p(torch.cuda.current_stream().cuda_stream))
    for it in range(1, 4):
        F.gemv_4bit(xc, q.t(), out=out, state=st)                     # the barrier is a link of the PDL chain after a GEMV
        F.lib.cbnb_peer_barrier(ct.c_void_p(counter.data_ptr()), ct.c_void_p(sig.data_ptr()), arr, ct.c_int32(2))
        y = F.gemv_4bit(xc, q.t(), state=st)                          # ... and in front of the next one
        torch.cuda.synchronize()
        assert int(counter.item()) == it and sig[:2].tolist() == [it, it]
        assert torch.equal(y.view(torch.int16), ref.view(torch.int16))
    assert F.lib.cbnb_last_error() == 0
