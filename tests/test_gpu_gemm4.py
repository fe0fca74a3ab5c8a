"""GPU parity: fused batch>1 4-bit GEMM (K4, tcgen05) through the C-ABI.
The kernel dequantises with the reference's arithmetic -- w = T(fp32 code * fp32 absmax), one rounding
(kernel_quant.cpp:1449-1450) -- so the oracle is: dequantised weight from the CPU restatement (bit-exact with the
K2 kernel), product in fp64.  Tolerance (stated): fp32 accumulation + ONE output rounding,
|y - y64| <= 2^-8 |y64| + 2^-9 rms(y64) for bf16, 2^-11 / 2^-12 for fp16."""
import numpy as np
import pytest
import torch

from helpers import DT, from_bits, to_bits
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    assert torch.cuda.is_available()
    from bnb_b200 import functional
    return functional


def reference64(F, x, q, st, dtype, bias=None):
    """fp64 product with the dequantised weight rounded to T exactly like dequantize_4bit."""
    Wd = F.dequantize_4bit(q, st).to(DT[dtype])            # K2 kernel: bit-exact with the oracle (test_gpu_blockwise)
    if Wd.numel() > (1 << 24):                             # config-4 sizes: the fp64 product itself runs on the device
        y = x.double().cuda() @ Wd.double().t()
        if bias is not None:
            y = y + bias.double()[None, :]
        return y.cpu().numpy()
    y = x.double().cpu().numpy() @ Wd.double().cpu().numpy().T
    if bias is not None:
        y = y + bias.double().cpu().numpy()[None, :]
    return y


CASES = [
    (16, 256, 512, "bf16", True, 64, False),
    (5, 130, 1024, "bf16", False, 64, True),       # batch not a multiple of 16, N not a multiple of 128, bias
    (33, 384, 4096, "fp16", True, 64, True),
    (256, 384, 2048, "bf16", True, 64, False),     # widest activation tile
    (32, 128, 8192, "bf16", True, 64, True),       # one row tile -> split-K + finalize
    (64, 4096, 14336, "bf16", True, 64, False),    # Llama-3-8B down projection: split-K
    (16, 1024, 4096, "fp16", False, 128, False),   # blocksize 128
    # BASELINE config 4 (Llama-3-8B MLP), the shapes and batches the kernel numbers are quoted on
    (16, 14336, 4096, "bf16", True, 64, False),
    (32, 14336, 4096, "bf16", True, 64, True),
    (128, 14336, 4096, "bf16", True, 64, False),
    (256, 14336, 4096, "bf16", True, 64, True),
    (16, 14336, 4096, "fp16", True, 64, True),
    (256, 14336, 4096, "fp16", True, 64, False),
    (128, 4096, 14336, "fp16", True, 64, False),
    (20, 14336, 4096, "bf16", True, 64, False),    # batch between the GEMV route (<= 8) and a full 32-row tile
    # routing edges of round 2's two kernels
    (48, 512, 4096, "bf16", True, 64, True),       # first width of the wide kernel (NB = 48, four dequant groups)
    (64, 14336, 4096, "bf16", True, 64, False),
    (192, 256, 4096, "bf16", True, 64, False),     # last width with a four-stage activation ring
    (200, 256, 2048, "fp16", False, 64, True),     # NB = 208: three stages, three groups
    (16, 256, 256, "bf16", True, 64, False),       # K too short for the small kernel (4 blocks): wide kernel at batch 16
    (24, 300, 320, "fp16", False, 64, True),       # 5 blocks, ragged N
    (100, 200, 1024, "fp16", True, 256, False),    # blocksize 256 in the wide kernel
    (32, 384, 1024, "fp16", False, 128, True),     # blocksize 128 in the small kernel, NB = 32
]


@pytest.mark.parametrize("batch,N,K,dtype,nested,blocksize,with_bias", CASES)
def test_gemm_4bit_vs_fp64(F, batch, N, K, dtype, nested, blocksize, with_bias):
    torch.manual_seed(batch * 7 + N)
    W = (torch.randn(N, K) * 0.02).to(DT[dtype])
    x = torch.randn(batch, K).to(DT[dtype])
    bias = (torch.randn(N) * 0.1).to(DT[dtype]).cuda() if with_bias else None
    q, st = F.quantize_4bit(W.cuda(), blocksize=blocksize, compress_statistics=nested, quant_type="nf4")
    y = F.gemm_4bit(x.cuda(), q.t(), st, bias=bias)
    assert y is not None, "fused kernel refused a supported shape"
    assert y.shape == (batch, N) and y.dtype == DT[dtype]
    y64 = reference64(F, x, q, st, dtype, bias)
    yk = y.double().cpu().numpy()
    assert np.all(np.isfinite(yk))
    y_ref = torch.nn.functional.linear(x.cuda(), F.dequantize_4bit(q, st).to(DT[dtype]), bias).double().cpu().numpy()
    if batch <= 32 and K >= 512:
        # small-batch route (k_gemm4_small): the MMA operand is the UNSCALED T(code[q]) and absmax multiplies fp32 block
        # sums -- code * absmax is never rounded to T.  Gate (stated): against the fp64 product with EXACT weights
        # (fp32 code * fp32 absmax), |y - exact| <= 2^-8 |exact| + 2^-7 rms for bf16 (2^-11 / 2^-9 for fp16) -- the gates
        # of the batch-1 GEMV, which has the same arithmetic -- and no less accurate than the reference's own
        # composition (dequantize_4bit -> T, then the library GEMM) with respect to that exact product.
        W32 = F.dequantize_4bit(q, F.QuantState(absmax=F._denest(st) if st.nested else st.absmax, shape=st.shape, code=st.code,
                                                blocksize=st.blocksize, quant_type=st.quant_type, dtype=torch.float32))
        exact = (x.double().cuda() @ W32.double().t())
        if bias is not None:
            exact = exact + bias.double()[None, :]
        exact = exact.cpu().numpy()
        rel, noise = {"bf16": (2.0 ** -8, 2.0 ** -7), "fp16": (2.0 ** -11, 2.0 ** -9)}[dtype]
        rms = np.sqrt(np.mean(exact ** 2))
        assert np.all(np.abs(yk - exact) <= rel * np.abs(exact) + noise * rms)
        err_k = np.linalg.norm(yk - exact) / np.linalg.norm(exact)
        err_ref = np.linalg.norm(y_ref - exact) / np.linalg.norm(exact)
        assert err_k <= err_ref * 1.05 + 1e-7, (err_k, err_ref)
        return
    rel, noise = {"bf16": (2.0 ** -8, 2.0 ** -9), "fp16": (2.0 ** -11, 2.0 ** -12)}[dtype]
    rms = np.sqrt(np.mean(y64 ** 2))
    assert np.all(np.abs(yk - y64) <= rel * np.abs(y64) + noise * rms)
    # and against the reference's own composition (dequantize_4bit + F.linear): same weights, library GEMM
    assert np.all(np.abs(yk - y_ref) <= 2 * rel * np.abs(y64) + 2 * noise * rms)


def test_gemm_4bit_deterministic_and_module_path(F):
    from bnb_b200.nn import LinearNF4
    torch.manual_seed(3)
    lin = LinearNF4(2048, 768, bias=True, compute_dtype=torch.bfloat16).cuda()
    x = torch.randn(4, 24, 2048, dtype=torch.bfloat16, device="cuda")
    with torch.no_grad():
        y1 = lin(x)
        y2 = lin(x)
    assert y1.shape == (4, 24, 768)
    assert torch.equal(y1.view(torch.int16), y2.view(torch.int16))
    Wd = F.dequantize_4bit(lin.weight.data, lin.weight.quant_state).to(torch.bfloat16)
    y_ref = torch.nn.functional.linear(x, Wd, lin.bias.to(torch.bfloat16))
    assert (y1.float() - y_ref.float()).abs().max().item() <= 2.0 ** -7 * y_ref.float().abs().max().item()


def test_gemm_4bit_refuses_what_it_cannot_do(F):
    """K not a multiple of 64 -> rc 1 -> Python falls back to the reference composition (still on the GPU)."""
    torch.manual_seed(0)
    W = (torch.randn(64, 96) * 0.02).bfloat16()
    q, st = F.quantize_4bit(W.cuda(), blocksize=64, quant_type="nf4")      # numel 6144 = 96 blocks
    x = torch.randn(8, 96).bfloat16().cuda()
    assert F.gemm_4bit(x, q.t(), st) is None
    import bnb_b200
    y = bnb_b200.matmul_4bit(x, q.t(), quant_state=st)
    assert y.shape == (8, 64) and torch.isfinite(y).all()


# ------------------------------------------------------------------------------------------------
# SURVEY 8f "next" rows on the 4-bit side: state-dict wire format and the backward caller
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("quant_type", ["nf4", "fp4"])
def test_linear4bit_state_dict_roundtrip(F, quant_type):
    """Linear4bit._save_to_state_dict (reference modules.py:436-445) writes the packed weight plus the QuantState
    components (absmax, nested state, `quant_state.bitsandbytes__<type>` JSON blob); Params4bit.from_prequantized
    (:281-311) rebuilds the parameter.  The reloaded layer must produce bit-identical outputs."""
    import bnb_b200
    torch.manual_seed(11)
    k, n = 512, 256
    lin = bnb_b200.nn.Linear4bit(k, n, bias=True, compute_dtype=torch.bfloat16, compress_statistics=True,
                                 quant_type=quant_type).cuda()
    x1 = torch.randn(1, k, device="cuda", dtype=torch.bfloat16)
    x8 = torch.randn(8, k, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        y1, y8 = lin(x1), lin(x8)
    sd = lin.state_dict()
    qs_keys = [key for key in sd if key.startswith("weight.")]
    assert any("quant_state.bitsandbytes__" + quant_type in key for key in qs_keys)
    assert "weight.absmax" in sd and "weight.nested_absmax" in sd and sd["weight"].dtype == torch.uint8
    stats = {key[len("weight."):]: v for key, v in sd.items() if key.startswith("weight.")}
    lin2 = bnb_b200.nn.Linear4bit(k, n, bias=True, compute_dtype=torch.bfloat16, compress_statistics=True,
                                  quant_type=quant_type)
    lin2.weight = bnb_b200.nn.Params4bit.from_prequantized(sd["weight"], stats, device="cuda")
    lin2.bias = torch.nn.Parameter(sd["bias"].clone())
    lin2 = lin2.cuda()
    with torch.no_grad():
        z1, z8 = lin2(x1), lin2(x8)
    assert torch.equal(y1, z1) and torch.equal(y8, z8)
    assert lin2.weight.quant_state == lin.weight.quant_state


def test_matmul4bit_backward(F):
    """MatMul4Bit.backward (reference _functions.py:520-540): grad_A = grad_out @ dequant(W), grad_bias = sum."""
    import bnb_b200
    torch.manual_seed(12)
    k, n, b = 512, 256, 24
    W = (torch.randn(n, k, device="cuda") * 0.02).to(torch.bfloat16)
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    Wd = F.dequantize_4bit(q, st).float()
    x = torch.randn(b, k, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    bias = torch.randn(n, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = bnb_b200.matmul_4bit(x, q.t(), quant_state=st, bias=bias)
    g = torch.randn(b, n, device="cuda", dtype=torch.bfloat16)
    y.backward(g)
    gA = g.float() @ Wd
    assert float((x.grad.float() - gA).norm() / gA.norm()) < 5e-3
    gb = g.float().sum(0)
    assert float((bias.grad.float() - gb).norm() / gb.norm()) < 1e-2


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("batch,N,K,with_bias", [(2, 4096, 4096, False), (5, 256, 1024, False), (8, 1024, 11008, False),
                                                 (5, 130, 1024, True)])   # last one: bias / ragged N -> stays on the tcgen05 kernel
def test_small_batch_rides_the_gemv_kernels(F, batch, N, K, dtype, with_bias):
    """batch 2..8 with a nested state: the batch is the MMA n dimension of the GEMV kernels (cgemm_4bit_inference_nested_*
    with n = batch).  Gate: rel-L2 vs the fp64 product with the EXACT (fp32 code x fp32 absmax) weights, the GEMV
    tolerances of test_gpu_gemv (2.5e-3 bf16 / 6e-4 fp16), per batch row."""
    import copy
    torch.manual_seed(batch * 100 + N)
    W = (torch.randn(N, K) * 0.02).to(DT[dtype]).cuda()
    x = torch.randn(batch, K).to(DT[dtype]).cuda()
    bias = torch.randn(N).to(DT[dtype]).cuda() if with_bias else None
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    y = F.gemm_4bit(x, q.t(), st, bias=bias)
    assert y is not None and y.shape == (batch, N) and y.dtype == DT[dtype]
    st32 = copy.copy(st)
    st32.dtype = torch.float32
    ref = x.double() @ F.dequantize_4bit(q, st32).double().t()
    if bias is not None:
        ref = ref + bias.double()
    tol = 2.5e-3 if dtype == "bf16" else 6e-4
    if with_bias:
        tol *= 1.6        # tcgen05 route: the operand is rounded to T first (reference arithmetic), see test_gemm_4bit_vs_fp64
    for b in range(batch):
        err = float((y[b].double() - ref[b]).norm() / ref[b].norm())
        assert err <= tol, (b, err)


def test_gemm_4bit_shape_checks_follow_the_reference(F):
    """The reference's batch>1 route is F.linear(A, dequantize_4bit(B, state).t()): a wrong activation width raises a
    shape error there, and an un-transposed packed weight computes A @ W.  The fused kernel must not silently read
    out of bounds or compute the other product: it refuses, and the reference-shaped route decides."""
    import bnb_b200
    torch.manual_seed(1)
    W = (torch.randn(128, 256) * 0.02).bfloat16()
    q, st = F.quantize_4bit(W.cuda(), blocksize=64, quant_type="nf4")
    x_bad = torch.randn(16, 192, device="cuda").bfloat16()
    assert F.gemm_4bit(x_bad, q.t(), st) is None
    with pytest.raises(RuntimeError):
        bnb_b200.matmul_4bit(x_bad, q.t(), quant_state=st)
    x_n = torch.randn(16, 128, device="cuda").bfloat16()          # A @ W (un-transposed B): [16,128] @ [128,256]
    assert F.gemm_4bit(x_n, q, st) is None
    y = bnb_b200.matmul_4bit(x_n, q, quant_state=st)
    ref = x_n.float() @ F.dequantize_4bit(q, st).float()
    assert y.shape == (16, 256) and (y.float() - ref).abs().max().item() <= 2.0 ** -6 * ref.abs().max().item()


@pytest.mark.parametrize("batch,N,K", [(16, 4096, 4096), (32, 14336, 4096), (64, 2048, 4096), (128, 4096, 4096), (256, 2048, 8192)])
def test_gemm_4bit_cold_caches(F, batch, N, K):
    """The fused GEMM with its weights NOT resident in L2 (a 256 MB fill between launches): round 1's two-CTAs-per-SM
    configuration returned wrong tiles in 37 of 40 such launches while every warm-cache test passed (tools/gemm4_stress.py).
    Both routes (batch <= 32: k_gemm4_small, above: k_gemm4_wide) must give the same bits on every launch, and the
    right ones."""
    torch.manual_seed(batch + N)
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(batch, K, device="cuda").bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    ref = (x.double() @ F.dequantize_4bit(q, st).double().t())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    first = None
    for it in range(12):
        flush.fill_(it)
        if it % 3 == 0:
            torch.cuda.synchronize()
        y = F.gemm_4bit(x, q.t(), st)
        assert y is not None
        if first is None:
            first = y.clone()
            assert float((y.double() - ref).norm() / ref.norm()) < 4e-3
        assert torch.equal(y.view(torch.int16), first.view(torch.int16)), f"launch {it} differs from launch 0"


@pytest.mark.parametrize("batch", [8, 32, 64, 200])
def test_gemm_4bit_strided_output_and_peer_stores(F, batch):
    """N-sharded form (cgemm_4bit_push_*): the result goes into a column slice of a wider [batch, ldo] buffer and, through
    peer_outs, into the same slice of other copies of that buffer.  One GPU stands in for the peers (a second and a third
    buffer on the same device); everything outside the slice must stay untouched, and the slice must carry the bits of the
    plain call."""
    torch.manual_seed(batch)
    N, K, ldo, off = 384, 2048, 1024, 256
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(batch, K, device="cuda").bfloat16()
    bias = (torch.randn(N, device="cuda") * 0.1).bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    plain = F.gemm_4bit(x, q.t(), st, bias=bias, out=torch.empty(batch, N, dtype=torch.bfloat16, device="cuda"))
    assert plain is not None
    bufs = [torch.full((batch, ldo), 7.0, dtype=torch.bfloat16, device="cuda") for _ in range(3)]
    peer_ptrs = [b.data_ptr() + off * b.element_size() for b in bufs[1:]]
    y = F.gemm_4bit(x, q.t(), st, bias=bias, out=bufs[0][:, off:off + N], peer_outs=peer_ptrs)
    assert y is not None
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b[:, off:off + N].contiguous().view(torch.int16), plain.view(torch.int16))
        assert bool((b[:, :off] == 7.0).all()) and bool((b[:, off + N:] == 7.0).all())
