"""GPU parity: the LLM.int8 path (K5-K7) through the C-ABI vs the oracle.
Bar: row/col stats, nnz counts, int8 codes, COO entries, layouts and int32 accumulators are bit-exact;
mm_dequant (fixed fp32 op order) is bit-exact as well."""
import numpy as np
import pytest
import torch

from helpers import bits_equal, rel_l2, to_bits
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    assert torch.cuda.is_available()
    from bnb_b200 import functional
    return functional


def make_acts(rows, cols, seed=0, outliers=True):
    torch.manual_seed(seed)
    A = torch.randn(rows, cols).half()
    if outliers:
        g = torch.Generator().manual_seed(seed + 1)
        for c in torch.randint(0, cols, (6,), generator=g).tolist():
            A[:, c] = 8.0 * torch.sign(torch.randn(rows, generator=g)).half()   # tests_pvc/test_matmulqlt.py:291-292 style
        A[rows // 2, cols // 3] = 6.0                                            # exactly the threshold
        A[0, 0] = -7.25
    return A


@pytest.mark.parametrize("shape", [(64, 256), (37, 300), (130, 1000), (512, 4096), (1, 8)])
@pytest.mark.parametrize("threshold", [0.0, 6.0])
def test_stats_and_double_quant_bit_exact(F, shape, threshold):
    rows, cols = shape
    A = make_acts(rows, cols, seed=rows, outliers=threshold > 0)
    An = A.numpy()
    rs, cs, nnz = F.get_colrow_absmax(A.cuda(), threshold=threshold)
    rs_ref, cs_ref, nnz_ref = orc.get_col_row_stats(An, threshold)
    assert np.array_equal(rs.cpu().numpy().view(np.uint32), rs_ref.view(np.uint32))
    assert np.array_equal(cs.cpu().numpy().view(np.uint32), cs_ref.view(np.uint32))
    if threshold > 0:
        assert np.array_equal(nnz.cpu().numpy(), np.cumsum(nnz_ref).astype(np.int32))
    else:
        assert nnz is None
    out_row, out_col, rs2, cs2, coo = F.double_quant(A.cuda(), threshold=threshold)
    ptr = None if nnz_ref is None else np.cumsum(nnz_ref).astype(np.int32)
    r_ref, c_ref, ri, ci, val = orc.double_rowcol_quant(An, rs_ref, cs_ref, ptr, threshold)
    assert np.array_equal(out_row.cpu().numpy(), r_ref)
    assert np.array_equal(out_col.cpu().numpy(), c_ref)
    if threshold > 0 and ri is not None and ri.size:
        assert coo is not None and coo.nnz == ri.size
        got = sorted(zip(coo.rowidx.cpu().tolist(), coo.colidx.cpu().tolist(),
                         coo.values.cpu().view(torch.int16).tolist()))
        want = sorted(zip(ri.tolist(), ci.tolist(), val.view(np.int16).tolist()))
        assert got == want                                   # as a set (reference order is atomic-arbitrary)
        assert torch.all(coo.rowidx[1:] >= coo.rowidx[:-1])  # sorted by row, like the reference's post-sort
        # outlier positions are 0 in the row-quantised matrix
        assert torch.all(out_row[coo.rowidx.long(), coo.colidx.long()] == 0)
    else:
        assert coo is None


def test_double_quant_zero_row_and_column(F):
    A = make_acts(32, 256, outliers=False)
    A[5, :] = 0
    A[:, 17] = 0
    out_row, out_col, rs, cs, _ = F.double_quant(A.cuda())
    assert rs[5].item() == 0.0 and cs[17].item() == 0.0
    assert torch.all(out_row[5] == 0) and torch.all(out_col[:, 17] == 0)     # 0 * inf = NaN -> 0 (defined)


@pytest.mark.parametrize("fmt", ["col32", "col_turing", "col_ampere"])
@pytest.mark.parametrize("shape", [(32, 32), (7, 33), (40, 100), (129, 65), (256, 512), (64, 288), (96, 1056), (2048, 1024)])   # the last four: tiled kernel
@pytest.mark.parametrize("transpose", [False, True])
def test_transforms_bit_exact(F, fmt, shape, transpose):
    rng = np.random.RandomState(shape[0])
    A = rng.randint(-128, 128, shape).astype(np.int8)
    out, S = F.transform(torch.from_numpy(A).cuda(), fmt, transpose=transpose)
    ref = orc.transform(A, fmt, transpose)
    assert out.numel() == ref.size
    assert np.array_equal(out.cpu().numpy().ravel(), ref)
    assert S[1] == fmt


@pytest.mark.parametrize("fmtB", ["col_turing", "col_ampere"])
@pytest.mark.parametrize("mnk", [(32, 32, 32), (19, 45, 96), (128, 256, 128), (200, 300, 528), (16, 40, 72)])
def test_igemmlt_reference_layouts_exact(F, fmtB, mnk):
    """cigemmlt_<fmt>_32 with the reference's operand layouts: exact int32, col32 output."""
    m, n, k = mnk
    rng = np.random.RandomState(m + n)
    A = rng.randint(-128, 128, (m, k)).astype(np.int8)
    B = rng.randint(-128, 128, (n, k)).astype(np.int8)
    C32A, SA = F.transform(torch.from_numpy(A).cuda(), "col32")
    CxB, SB = F.transform(torch.from_numpy(B).cuda(), fmtB)
    out, Sout = F.igemmlt(C32A, CxB, SA, SB)
    assert Sout[1] == "col32" and out.dtype == torch.int32
    ref = orc.igemmlt_32(orc.transform(A, "col32"), orc.transform(B, fmtB), m, n, k, fmtB)
    assert np.array_equal(out.cpu().numpy().ravel(), ref)
    idx = np.unique(rng.randint(0, k, 5)).astype(np.int32)
    got = F.extract_outliers(CxB, SB, torch.from_numpy(idx).cuda())
    assert np.array_equal(got.cpu().numpy(), B[:, idx])


@pytest.mark.parametrize("fmtB", ["col_turing", "col_ampere"])
@pytest.mark.parametrize("rowscale", [False, True])
@pytest.mark.parametrize("mnk", [(32, 32, 32), (19, 45, 96), (200, 300, 528)])
def test_igemmlt_int8_output_variants(F, fmtB, rowscale, mnk):
    """cigemmlt_<fmt>_8 / _8_rowscale (reference op_gemm.cpp:604-638, pythonInterface.cpp:303-316): saturated int8
    col32 output of the fp32-scaled accumulator, bit-exact against oracle.igemmlt_8 -- including saturation (small
    operands keep part of the products inside [-128, 127], the rest clip)."""
    import ctypes as ct
    m, n, k = mnk
    rng = np.random.RandomState(m * 3 + n)
    A = rng.randint(-3, 4, (m, k)).astype(np.int8)
    B = rng.randint(-3, 4, (n, k)).astype(np.int8)
    A[0], B[0] = 127, 127                                   # one saturating corner, both signs
    A[1] = -128
    scale = (rng.rand(m).astype(np.float32) * 2.0 + 0.01) if rowscale else None
    C32A, SA = F.transform(torch.from_numpy(A).cuda(), "col32")
    CxB, SB = F.transform(torch.from_numpy(B).cuda(), fmtB)
    ref = orc.igemmlt_8(orc.transform(A, "col32"), orc.transform(B, fmtB), m, n, k, fmtB, scale)
    if not rowscale:
        out, Sout = F.igemmlt(C32A, CxB, SA, SB, dtype=torch.int8)
        assert Sout[1] == "col32" and out.dtype == torch.int8
    else:
        out, Sout = F.get_transform_buffer((m, n), torch.int8, C32A.device, "col32", "row")
        rs = torch.from_numpy(scale).cuda()
        fmt = "turing" if fmtB == "col_turing" else "ampere"
        ldb = ((n + 7) // 8) * 8 * 32 if fmtB == "col_turing" else ((n + 31) // 32) * 32 * 32
        rc = getattr(F.lib, f"cigemmlt_{fmt}_8_rowscale")(ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), F.get_ptr(C32A), F.get_ptr(CxB),
                                                         F.get_ptr(out), F.get_ptr(rs), ct.c_int32(m * 32), ct.c_int32(ldb), ct.c_int32(m * 32))
        assert rc == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy().ravel()
    assert np.array_equal(got, ref)
    assert (np.abs(ref.astype(np.int32)) == 127).any() or (ref == -128).any()      # the case does saturate
    assert (np.abs(ref.astype(np.int32)) < 100).any()


@pytest.mark.parametrize("mnk", [(128, 256, 128), (1, 8, 16), (130, 260, 144), (384, 512, 1024), (1000, 1000, 1008),
                                 (64, 64, 40)])
def test_igemm_rowmajor_exact(F, mnk):
    """B200-native row-major tcgen05 kind::i8 GEMM (and the SIMT path for K % 16 != 0): exact int32."""
    m, n, k = mnk
    rng = np.random.RandomState(k)
    A = rng.randint(-128, 128, (m, k)).astype(np.int8)
    B = rng.randint(-128, 128, (n, k)).astype(np.int8)
    out, _ = F.igemmlt(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), ((m, k), "row"), ((n, k), "row"))
    ref = A.astype(np.int32) @ B.astype(np.int32).T
    assert np.array_equal(out.cpu().numpy(), ref)


def test_igemm_extreme_values_exact(F):
    """all -128 * -128 over K = 4096: 2^26 per output, far from overflow; +-127 / -128 mixes."""
    m, n, k = 128, 256, 4096
    A = np.full((m, k), -128, np.int8)
    B = np.full((n, k), -128, np.int8)
    B[1::2] = 127
    out, _ = F.igemmlt(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), ((m, k), "row"), ((n, k), "row"))
    ref = A.astype(np.int64) @ B.astype(np.int64).T
    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)


@pytest.mark.parametrize("shape", [(128, 256), (77, 100), (256, 1024)])
@pytest.mark.parametrize("bias", [False, True])
def test_mm_dequant_bit_exact(F, shape, bias):
    rows, cols = shape
    rng = np.random.RandomState(rows)
    C = rng.randint(-2 ** 22, 2 ** 22, (rows, cols)).astype(np.int32)
    rs = np.abs(rng.randn(rows)).astype(np.float32) * 3
    cs = np.abs(rng.randn(cols)).astype(np.float32)
    b = (rng.randn(cols)).astype(np.float16) if bias else None
    C32 = orc.transform(C, "col32")
    out = F.mm_dequant(torch.from_numpy(C32).cuda(), ((rows, cols), "col32"), torch.from_numpy(rs).cuda(),
                       torch.from_numpy(cs).cuda(), bias=None if b is None else torch.from_numpy(b).cuda())
    ref = orc.mm_dequant(C32, rs, cs, rows, cols, b, col32=True)
    assert bits_equal(out, ref.view(np.uint16))


@pytest.mark.parametrize("mnk", [(256, 512, 256), (130, 264, 144)])
def test_fused_igemm_dequant_bit_exact(F, mnk):
    """cigemm_rowmajor_dequant_fp16 == igemm (exact) -> mm_dequant (oracle), bit for bit."""
    m, n, k = mnk
    rng = np.random.RandomState(7)
    A = rng.randint(-127, 128, (m, k)).astype(np.int8)
    B = rng.randint(-127, 128, (n, k)).astype(np.int8)
    rs = (np.abs(rng.randn(m)) + 0.5).astype(np.float32)
    cs = (np.abs(rng.randn(n)) + 0.5).astype(np.float32)
    b = rng.randn(n).astype(np.float16)
    out = F.int8_linear_dequant(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(rs).cuda(),
                                torch.from_numpy(cs).cuda(), bias=torch.from_numpy(b).cuda())
    C = A.astype(np.int32) @ B.astype(np.int32).T
    ref = orc.mm_dequant(C, rs, cs, m, n, b, col32=False)
    assert bits_equal(out, ref.view(np.uint16))


@pytest.mark.parametrize("threshold", [0.0, 6.0])
@pytest.mark.parametrize("has_bias", [False, True])
def test_linear8bitlt_forward(F, threshold, has_bias):
    """Module-level check in the style of the reference's tests (tests_pvc/autograd.py:216-324,
    tolerances :277-280): <= 1.75 % of elements outside atol 0.01 / rtol 0.1 vs the fp16 torch result."""
    from bnb_b200.nn import Linear8bitLt
    torch.manual_seed(3)
    lin = Linear8bitLt(1024, 768, bias=has_bias, has_fp16_weights=False, threshold=threshold)
    torch.nn.init.xavier_uniform_(lin.weight)
    W = lin.weight.data.clone().half()
    b = lin.bias.data.clone().half() if has_bias else None
    lin = lin.cuda().half()
    assert lin.weight.dtype == torch.int8
    A = torch.randn(96, 1024, dtype=torch.float16, device="cuda")
    if threshold > 0:
        A[:, [5, 100, 777]] = 6.0                      # autograd.py:228-230
    out = lin(A)
    ref = torch.nn.functional.linear(A, W.cuda(), None if b is None else b.cuda())
    assert out.shape == ref.shape and out.dtype == torch.float16
    close = torch.isclose(out, ref, atol=0.01, rtol=0.1)
    assert (close == 0).sum().item() <= out.numel() * 0.0175
    close2 = torch.isclose(out, ref, atol=0.035, rtol=0.2)
    assert (close2 == 0).sum().item() <= out.numel() * 0.001
    out2 = lin(A)                                      # second call reuses the state
    assert torch.equal(out, out2)


def test_linear8bitlt_matches_reference_layout_path(F):
    """Native row-major fused path == the reference-shaped path (col32/col_turing transforms, igemmlt,
    mm_dequant) bit for bit."""
    from bnb_b200 import matmul, MatmulLtState
    torch.manual_seed(9)
    W = (torch.randn(512, 1024) * 0.03).half().cuda()
    A = torch.randn(64, 1024, dtype=torch.float16, device="cuda")
    A[:, 3] = 7.0
    bias = torch.randn(512, dtype=torch.float16, device="cuda")
    outs = []
    for fmt in ("row", "col_turing", "col_ampere"):
        st = MatmulLtState()
        st.formatB = fmt
        st.threshold = 6.0
        st.has_fp16_weights = False
        CB, _, SCB, _, _ = F.double_quant(W)
        st.CB, st.SCB = CB, SCB
        outs.append(matmul(A, CB, state=st, bias=bias))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_config3_full_size_checksum(F):
    """BASELINE config 3 at full size: 4096 tokens x 4096 -> 16384, exact int32.  Size-independent
    property: sum_j C[i, j] == A[i, :] . (sum_j B[j, :]) and sum_i C[i, j] == (sum_i A[i, :]) . B[j, :]
    in exact integer arithmetic, plus a 16-row slab against the CPU oracle."""
    m, k, n = 4096, 4096, 16384
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randint(-127, 128, (m, k), dtype=torch.int8, device="cuda", generator=g)
    B = torch.randint(-127, 128, (n, k), dtype=torch.int8, device="cuda", generator=g)
    C, _ = F.igemmlt(A, B, ((m, k), "row"), ((n, k), "row"))
    row_sum = C.sum(dim=1, dtype=torch.int64)
    col_sum = C.sum(dim=0, dtype=torch.int64)
    Bs = B.sum(dim=0, dtype=torch.int64).double()          # |.| <= 127*16384 < 2^21, products sum < 2^45: exact in fp64
    As = A.sum(dim=0, dtype=torch.int64).double()
    assert torch.equal(row_sum, (A.double() @ Bs).long())
    assert torch.equal(col_sum, (B.double() @ As).long())
    slab = orc.igemm_rowmajor(A[1000:1016].cpu().numpy(), B[5000:5128].cpu().numpy())
    assert np.array_equal(C[1000:1016, 5000:5128].cpu().numpy(), slab)


# ------------------------------------------------------------------------------------------------
# fused LLM.int8 inference forward (additive cint8_linear_fp16) vs the step-by-step reference orchestration
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_outlier_cols", [0, 3, 16, 23])
@pytest.mark.parametrize("with_bias", [False, True])
def test_int8_linear_fused_matches_stepwise(F, n_outlier_cols, with_bias):
    """Same quantised activations / row statistics (bit-exact) and the same fp16 output (<= 2 ulp: only the fp32
    summation order of the tiny outlier product differs, and that product is rounded to fp16 before it is added) as double_quant -> zero outlier columns -> int8 GEMM ->
    mm_dequant -> + subA @ subB (reference autograd/_functions.py:292-434)."""
    import bnb_b200
    torch.manual_seed(100 + n_outlier_cols)
    m, k, n = 200, 1024, 384
    A = torch.randn(m, k, device="cuda").half()
    cols = torch.randperm(k)[:n_outlier_cols].sort().values
    for c in cols.tolist():          # an outlier column: a few rows at +-8, the rest ordinary (as in real activations)
        rows = torch.randperm(m)[: max(1, m // 7)]
        A[rows.cuda(), c] = 8.0 * (1 if c % 2 else -1)
    lin = bnb_b200.nn.Linear8bitLt(k, n, bias=with_bias, has_fp16_weights=False, threshold=6.0).cuda().half()
    with torch.no_grad():
        old = F.FUSED_INT8_LINEAR
        try:
            F.FUSED_INT8_LINEAR = False
            y_step = lin(A)
            F.FUSED_INT8_LINEAR = True
            CB = lin.state.CB if lin.state.CB is not None else lin.state.CxB
            res = F.int8_linear_fused(A, CB, lin.state.SCB, bias=lin.bias, threshold=6.0, return_quantized=True)
            assert res is not None
            y_fused, CA, SCA, idx, count = res
            y_mod = lin(A)                       # the module takes the fused path too
        finally:
            F.FUSED_INT8_LINEAR = old
    torch.cuda.synchronize()
    assert int(count.item()) == n_outlier_cols
    assert torch.equal(idx[:n_outlier_cols].cpu().long(), cols)
    CA_ref, _, SCA_ref, _, coo = F.double_quant(A, threshold=6.0)
    CA_ref[:, cols.cuda()] = 0
    assert torch.equal(SCA, SCA_ref)
    assert torch.equal(CA, CA_ref)
    assert torch.equal(y_fused, y_mod)
    d = (y_fused.float() - y_step.float()).abs()
    # the outlier product U is rounded to fp16 before it is added: a different fp32 summation order can move half(U) by
    # one ulp OF U, which is more than an ulp of the result where the two terms cancel -> scale the ulp by max(|y|, |U|)
    U = (A[:, cols.cuda()].float() @ (CB[:, cols.cuda()].float() * lin.state.SCB.float().view(-1, 1) / 127.0).half().float().t()).abs()
    ulp = torch.maximum(torch.maximum(y_step.float().abs(), U), torch.tensor(2.0 ** -14, device="cuda")) * 2.0 ** -10
    assert bool((d <= 2.01 * ulp).all()), float((d / ulp).max())
    assert float((d > 0).float().mean()) < 0.02   # and almost everywhere bit-identical


def test_int8_linear_fused_exact_vs_oracle(F):
    """int32 accumulators are exact and the dequant chain is the oracle's: without outliers the fused output is
    bit-identical to oracle mm_dequant(igemm(CA, CB))."""
    torch.manual_seed(5)
    m, k, n = 96, 512, 256
    A = (torch.randn(m, k) * 0.5).half()
    W = (torch.randn(n, k) * 0.05).half()
    CB, _, SCB, _, _ = F.double_quant(W.cuda())
    bias = torch.randn(n).half()
    y, CA, SCA, idx, count = F.int8_linear_fused(A.cuda(), CB, SCB, bias=bias.cuda(), threshold=6.0, return_quantized=True)
    assert int(count.item()) == 0
    acc = orc.igemm_rowmajor(CA.cpu().numpy(), CB.cpu().numpy())
    ref = orc.mm_dequant(acc, SCA.cpu().numpy(), SCB.cpu().numpy(), m, n, bias.view(torch.int16).numpy().view(np.uint16), col32=False)
    assert np.array_equal(y.cpu().view(torch.int16).numpy().view(np.uint16), ref.view(np.uint16))


# ------------------------------------------------------------------------------------------------
# MatMul8bitLt.backward (reference autograd/_functions.py:436-483) -- SURVEY 8f item 2
# ------------------------------------------------------------------------------------------------
def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.mark.parametrize("has_fp16_weights", [True, False])
@pytest.mark.parametrize("threshold", [0.0, 6.0])
def test_matmul8bitlt_backward(F, has_fp16_weights, threshold):
    """Gradients of Linear8bitLt against the fp32 linear layer with the same weights: int8 quantisation noise only
    (the reference's own tests use statistical tolerances, tests_pvc/autograd.py:277-280, 389-391)."""
    import bnb_b200
    torch.manual_seed(7)
    m, k, n = 64, 256, 128
    W = (torch.randn(n, k) * 0.05).half()
    bias = (torch.randn(n) * 0.1).half()
    x = torch.randn(m, k).half()
    if threshold > 0:
        x[:, 5] = 8.0
        x[3, 77] = -9.0
    lin = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=has_fp16_weights, threshold=threshold)
    lin.weight.data.copy_(W)
    lin.bias.data.copy_(bias)
    lin = lin.cuda().half()
    lin.train()
    xg = x.cuda().clone().requires_grad_(True)
    y = lin(xg)
    g = torch.randn(m, n, device="cuda").half()
    y.backward(g)
    # fp32 reference
    xr = x.cuda().float().requires_grad_(True)
    Wr = W.cuda().float().requires_grad_(True)
    br = bias.cuda().float().requires_grad_(True)
    yr = xr @ Wr.t() + br
    yr.backward(g.float())
    assert _rel(y, yr) < 0.02
    assert _rel(xg.grad, xr.grad) < 0.03
    assert _rel(lin.bias.grad, br.grad) < 0.01
    if has_fp16_weights:
        assert lin.weight.grad is not None
        assert _rel(lin.weight.grad, Wr.grad) < 0.03
    else:
        assert lin.weight.grad is None   # frozen int8 weight (LoRA-style fine-tuning): only grad_A flows


def test_linear8bitlt_state_dict_roundtrip(F):
    """SCB + weight_format travel through state_dict and the reloaded module gives identical outputs (reference
    modules.py:725-796)."""
    import bnb_b200
    torch.manual_seed(3)
    k, n = 256, 128
    lin = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
    x = torch.randn(16, k, device="cuda").half()
    with torch.no_grad():
        y0 = lin(x)
    sd = lin.state_dict()
    assert "SCB" in sd and "weight_format" in sd and sd["weight"].dtype == torch.int8
    lin2 = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
    lin2.load_state_dict(sd)
    with torch.no_grad():
        y1 = lin2(x)
    assert torch.equal(y0, y1)


@pytest.mark.parametrize("n_outlier_cols", [0, 2, 5])
def test_int8_linear_fused_vs_oracle_orchestration(F, n_outlier_cols):
    """The fused forward against the ORACLE's restatement of the reference orchestration (oracle.llm_int8_forward,
    _functions.py:292-434): quantised activations, row statistics and outlier columns bit-exact; fp16 output within
    2 ulp of max(|y|, |outlier term|) (only the fp32 summation order of the <= 8-term outlier product differs) and
    bit-identical without outliers."""
    torch.manual_seed(40 + n_outlier_cols)
    m, k, n = 96, 512, 256
    A = torch.randn(m, k).half()
    cols = sorted(torch.randperm(k)[:n_outlier_cols].tolist())
    for c in cols:
        A[torch.randperm(m)[: m // 5], c] = 7.0 if c % 2 else -9.0
    W = (torch.randn(n, k) * 0.05).half()
    CB, _, SCB, _, _ = F.double_quant(W.cuda())
    bias = torch.randn(n).half()
    y, CA, SCA, idx, count = F.int8_linear_fused(A.cuda(), CB, SCB, bias=bias.cuda(), threshold=6.0, return_quantized=True)
    torch.cuda.synchronize()
    y_o, CA_o, SCA_o, idx_o = orc.llm_int8_forward(A.numpy(), CB.cpu().numpy(), SCB.cpu().numpy(), bias.numpy(), 6.0)
    assert int(count.item()) == n_outlier_cols and idx[:n_outlier_cols].cpu().tolist() == idx_o.tolist() == cols
    assert np.array_equal(CA.cpu().numpy(), CA_o)
    assert np.array_equal(SCA.cpu().numpy().view(np.uint32), SCA_o.view(np.uint32))
    yk = y.cpu().numpy()
    if n_outlier_cols == 0:
        assert np.array_equal(yk.view(np.uint16), y_o.view(np.uint16))
    else:
        subB = ((CB.cpu().numpy()[:, cols].astype(np.float32) * SCB.cpu().numpy()[:, None]) / np.float32(127.0)).astype(np.float16)
        U = np.abs(A.numpy()[:, cols].astype(np.float32) @ subB.astype(np.float32).T)
        ulp = np.maximum(np.maximum(np.abs(y_o.astype(np.float32)), U), 2.0 ** -14) * 2.0 ** -10
        d = np.abs(yk.astype(np.float32) - y_o.astype(np.float32))
        assert (d <= 2.01 * ulp).all(), float((d / ulp).max())
        assert (d > 0).mean() < 0.02


@pytest.mark.parametrize("k", [2048, 4096, 8192])
@pytest.mark.parametrize("m", [96, 4096])
@pytest.mark.parametrize("n_outlier_cols", [0, 4, 9, 23])
def test_int8_linear_fused_config3_paths_vs_oracle(F, k, m, n_outlier_cols):
    """BASELINE config 3 takes k_i8_row_onepass<16> (K = 4096); K <= 2048 takes <8>, K > 4096 the two-pass route
    (csrc/int8_fused.cu) -- each against the oracle's restatement of MatMul8bitLt.forward (reference
    autograd/_functions.py:292-434; kernel arithmetic kernel_quant.cpp:3292-3301, 3424, 3475).  m = 4096 is the
    config's token count; the weight has 128 output rows so that the oracle's integer GEMM stays a few seconds.
    Bar: CA, SCA, outlier column list bit-exact; fp16 output bit-identical without outliers, within 2 ulp of
    max(|y|, |outlier term|) with them (<= 8 columns ride in the GEMM epilogue, more in the follow-up kernel)."""
    torch.manual_seed(1000 + k + m + n_outlier_cols)
    n = 128 if m > 1000 else 256
    A = torch.randn(m, k).half()
    cols = sorted(torch.randperm(k)[:n_outlier_cols].tolist())
    for c in cols:
        A[torch.randperm(m)[: max(1, m // 5)], c] = 7.0 if c % 2 else -9.0
    if cols:
        A[0, cols[0]] = 6.0                      # exactly the threshold counts as an outlier (>=)
    W = (torch.randn(n, k) * 0.05).half()
    CB, _, SCB, _, _ = F.double_quant(W.cuda())
    bias = torch.randn(n).half()
    res = F.int8_linear_fused(A.cuda(), CB, SCB, bias=bias.cuda(), threshold=6.0, return_quantized=True)
    assert res is not None
    y, CA, SCA, idx, count = res
    torch.cuda.synchronize()
    y_o, CA_o, SCA_o, idx_o = orc.llm_int8_forward(A.numpy(), CB.cpu().numpy(), SCB.cpu().numpy(), bias.numpy(), 6.0)
    assert int(count.item()) == n_outlier_cols and idx[:n_outlier_cols].cpu().tolist() == idx_o.tolist() == cols
    assert np.array_equal(CA.cpu().numpy(), CA_o)
    assert np.array_equal(SCA.cpu().numpy().view(np.uint32), SCA_o.view(np.uint32))
    yk = y.cpu().numpy()
    if n_outlier_cols == 0:
        assert np.array_equal(yk.view(np.uint16), y_o.view(np.uint16))
    else:
        subB = ((CB.cpu().numpy()[:, cols].astype(np.float32) * SCB.cpu().numpy()[:, None]) / np.float32(127.0)).astype(np.float16)
        U = np.abs(A.numpy()[:, cols].astype(np.float32) @ subB.astype(np.float32).T)
        ulp = np.maximum(np.maximum(np.abs(y_o.astype(np.float32)), U), 2.0 ** -14) * 2.0 ** -10
        d = np.abs(yk.astype(np.float32) - y_o.astype(np.float32))
        assert (d <= 2.01 * ulp).all(), float((d / ulp).max())
        assert (d > 0).mean() < 0.02


@pytest.mark.parametrize("fmt", ["col32", "col_turing", "col_ampere"])
@pytest.mark.parametrize("shape", [(64, 64), (37, 300), (512, 4096), (96, 288), (32, 1056)])
def test_inverse_layout_transforms(F, fmt, shape):
    """ctransform_{col32,turing,ampere}2row (the reference's Python calls the last two at functional.py:2645-2647):
    row -> layout -> row is the identity, and the device inverse agrees with the host-side undo_layout_to_row."""
    rows, cols = shape
    rng = np.random.RandomState(rows * 7 + cols)
    A = torch.from_numpy(rng.randint(-128, 128, (rows, cols)).astype(np.int8)).cuda()
    buf, S = F.transform(A, fmt)
    back, S2 = F.transform(buf, "row", state=S)
    assert S2[1] == "row" and back.shape == (rows, cols)
    assert torch.equal(back, A)
    host = F.undo_layout_to_row(buf.cpu(), fmt, rows, cols)
    assert torch.equal(host, A.cpu())
    assert np.array_equal(buf.cpu().numpy().ravel(), orc.transform(A.cpu().numpy(), fmt))


@pytest.mark.parametrize("fmt", ["col_turing", "col_ampere"])
def test_linear8bitlt_loads_reference_format_checkpoint(F, fmt):
    """A state dict as upstream bitsandbytes / the reference writes it from `state.CxB` (nn/modules.py:725-796): int8
    weight in col_turing / col_ampere + `weight_format` code + SCB.  maybe_rearrange_weight (:635-654) must bring it
    back to row-major at load time; the loaded layer computes exactly what the row-format layer computes."""
    import bnb_b200
    from bnb_b200.utils import LINEAR_8BIT_WEIGHTS_FORMAT_MAPPING
    torch.manual_seed(21)
    k, n = 256, 96 if fmt == "col_turing" else 128       # whole 8- / 32-row tiles, as undo_layout requires
    lin = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
    x = torch.randn(24, k, device="cuda").half()
    x[:, 5] = 7.0
    with torch.no_grad():
        y0 = lin(x)
    sd = lin.state_dict()
    CB = sd["weight"]
    assert CB.dtype == torch.int8 and CB.shape == (n, k)
    CxB, _ = F.transform(CB.cuda(), fmt)
    sd_ref = {"weight": CxB.cpu(), "bias": sd["bias"].cpu(), "SCB": sd["SCB"].cpu(),
              "weight_format": torch.tensor(LINEAR_8BIT_WEIGHTS_FORMAT_MAPPING[fmt], dtype=torch.uint8)}
    lin2 = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
    lin2.load_state_dict(sd_ref)
    with torch.no_grad():
        y1 = lin2(x)
    assert torch.equal(y0, y1)


@pytest.mark.parametrize("has_fp16_weights", [False, True])
@pytest.mark.parametrize("threshold", [0.0, 6.0])
def test_matmul8bitlt_backward_vs_oracle(F, has_fp16_weights, threshold):
    """MatMul8bitLt.backward against the oracle's restatement of reference _functions.py:436-483 (oracle.llm_int8_backward):
    the int8 products are exact and mm_dequant has a fixed fp32 order, so grad_B and the int8-route grad_A are bit-exact
    where no 16-bit side product is added; the 16-bit GEMMs (frozen-weight grad_A, outlier columns of grad_B) agree to
    2 fp16 ulp of the magnitudes involved (fp32 summation order of the BLAS)."""
    import bnb_b200
    torch.manual_seed(17)
    m, k, n = 96, 256, 128
    W = (torch.randn(n, k) * 0.05).half()
    x = torch.randn(m, k).half()
    if threshold > 0:
        x[:, 9] = 8.0
        x[5, 100] = -7.5
    lin = bnb_b200.nn.Linear8bitLt(k, n, bias=False, has_fp16_weights=has_fp16_weights, threshold=threshold)
    lin.weight.data.copy_(W)
    lin = lin.cuda().half()
    lin.train()
    xg = x.cuda().clone().requires_grad_(True)
    y = lin(xg)
    g = torch.randn(m, n).half()
    y.backward(g.cuda())
    torch.cuda.synchronize()
    if has_fp16_weights:
        gA, gB = orc.llm_int8_backward(g.numpy(), x.numpy(), threshold, W_f16=W.numpy(), need_grad_B=True)
        got_B = lin.weight.grad.cpu().numpy()
        idx = [9, 100] if threshold > 0 else []
        clean = np.ones(k, bool)
        clean[idx] = False
        assert np.array_equal(got_B[:, clean].view(np.uint16), gB[:, clean].view(np.uint16))        # int8 route: bit-exact
        if idx:
            d = np.abs(got_B[:, idx].astype(np.float32) - gB[:, idx].astype(np.float32))
            mag = np.maximum(np.abs(gB[:, idx].astype(np.float32)), np.abs(g.numpy().astype(np.float32)).T @ np.abs(x.numpy()[:, idx].astype(np.float32)) * 2.0 ** -3)
            assert (d <= 2.01 * 2.0 ** -10 * np.maximum(mag, 2.0 ** -14)).all()
        assert np.array_equal(xg.grad.cpu().numpy().view(np.uint16), gA.view(np.uint16))             # int8 route: bit-exact
    else:
        CB = lin.state.CB if lin.state.CB is not None else lin.state.CxB
        gA, _ = orc.llm_int8_backward(g.numpy(), x.numpy(), threshold, CB=CB.cpu().numpy(), SCB=lin.state.SCB.cpu().numpy())
        got = xg.grad.cpu().numpy().astype(np.float32)
        scale = np.abs(g.numpy().astype(np.float32)) @ np.abs((CB.cpu().numpy().astype(np.float32) * lin.state.SCB.cpu().numpy()[:, None] / 127.0))
        assert (np.abs(got - gA.astype(np.float32)) <= 2.0 ** -10 * np.maximum(np.abs(gA.astype(np.float32)), scale * 2.0 ** -4) * 2.01 + 1e-7).all()
        assert lin.weight.grad is None
