"""CPU-only: pin the oracle (oracle/bnb_oracle.c) against the golden fixtures generated from the
reference itself (tests/golden/make_golden.py) and against internal consistency properties."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def dec_to_f32(s: str) -> np.float32:
    """Correctly-rounded decimal literal -> fp32 (what a C/C++ compiler does with `<s>f`).
    NOT np.float32(float(s)): several NF4 thresholds are exact fp32 ties when seen as doubles, so the
    double detour rounds them the other way."""
    from fractions import Fraction
    exact = Fraction(s)
    f = np.float32(float(s))
    cands = [np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))]
    best = min(cands, key=lambda c: (abs(Fraction(float(c)) - exact), int(c.view(np.uint32)) & 1))
    return np.float32(best)


@pytest.fixture(scope="module")
def tables():
    return np.load(os.path.join(GOLDEN, "ref_python_tables.npz"))


@pytest.fixture(scope="module")
def consts():
    return json.load(open(os.path.join(GOLDEN, "ref_kernel_constants.json")))


def test_nf4_table_matches_reference_kernel_and_python(tables, consts):
    t = orc.nf4_table()
    ref_kernel = np.array([dec_to_f32(v) for v in consts["nf4_table"]], np.float32)
    assert np.array_equal(t, ref_kernel)           # kernel_quant.cpp:650-703
    assert np.array_equal(t, tables["nf4"])        # functional.py:1035-1052


def test_fp4_table_matches_reference_kernel_and_python(tables, consts):
    t = orc.fp4_table()
    mags = np.array([dec_to_f32(v) for v in consts["fp4_dequant_by_low3bits"]], np.float32)
    assert np.array_equal(t[:8], mags)
    assert np.array_equal(t[8:], -mags) and np.signbit(t[8])
    # python get_4bit_type("fp4") = data / 12 (functional.py:1063): same fp32 values
    assert np.array_equal(np.abs(t), np.abs(tables["fp4"]))


def test_nf4_thresholds_are_the_reference_tree(consts):
    thr = np.array([dec_to_f32(v) for v in consts["nf4_thresholds_ascending"]], np.float32)
    L = orc.lib()
    for i, t in enumerate(thr):
        # strict '>' : exactly on the threshold stays below, next float up goes above
        assert L.orc_quantize_nf4_scalar(float(t)) == i
        assert L.orc_quantize_nf4_scalar(float(np.nextafter(t, np.float32(2)))) == i + 1
    assert L.orc_quantize_nf4_scalar(float("nan")) == 0
    # thresholds are the midpoints of adjacent table entries up to the decimal-literal rounding
    # (SURVEY Appendix B): within 1 ulp of fp32(midpoint)
    tab = orc.nf4_table().astype(np.float64)
    mid = ((tab[:-1] + tab[1:]) / 2).astype(np.float32)
    assert np.max(np.abs(mid.view(np.int32).astype(np.int64) - thr.view(np.int32).astype(np.int64))) <= 1


def test_fp4_thresholds_are_the_reference_tree(consts):
    thr = sorted(dec_to_f32(v) for v in consts["fp4_quant_thresholds_tree_order"])
    # ascending magnitude buckets -> low-3-bit codes (kernel_quant.cpp:570-593)
    codes = [0, 1, 6, 7, 4, 5, 2, 3]
    L = orc.lib()
    probe = [np.float32(0.0)] + [np.nextafter(t, np.float32(2)) for t in thr]
    for x, c in zip(probe, codes):
        assert L.orc_quantize_fp4_scalar(float(x)) == c
        assert L.orc_quantize_fp4_scalar(float(-x)) == (c | 8 if x > 0 else c)
    for t, c in zip(thr, codes):
        assert L.orc_quantize_fp4_scalar(float(t)) == c
    assert L.orc_quantize_fp4_scalar(float("nan")) == 0
    assert L.orc_quantize_fp4_scalar(-0.0) == 0


def test_mm_dequant_const(consts):
    C = np.array([[12345, -777]], np.int32)
    out = orc.mm_dequant(C, np.array([2.0], np.float32), np.array([3.0, 0.5], np.float32), 1, 2, col32=False)
    k = dec_to_f32(consts["mm_dequant_const"])
    exp = np.array([[np.float32(np.float32(np.float32(np.float32(12345) * k) * np.float32(2.0)) * np.float32(3.0)),
                     np.float32(np.float32(np.float32(np.float32(-777) * k) * np.float32(2.0)) * np.float32(0.5))]],
                   np.float32).astype(np.float16)
    assert np.array_equal(out.view(np.uint16), exp.view(np.uint16))


# ------------------------------------------------------------------ reference CPU ops (oracle/_ref)
@pytest.mark.parametrize("case,bs", [("bs64", 64), ("bs4096", 4096)])
def test_port_matches_reference_cpu_blockwise_golden(tables, case, bs):
    g = np.load(os.path.join(GOLDEN, "ref_cpu_blockwise.npz"))
    A, q_ref, am_ref, deq_ref = g[case + "_A"], g[case + "_q"], g[case + "_absmax"], g[case + "_deq"]
    q, am, code_after = orc.quantize_cpu_port(tables["dynamic_map"], A, bs)
    assert np.array_equal(am, am_ref)
    assert np.array_equal(code_after, g[case + "_code_after"]) and code_after[0] == -1.0
    assert np.array_equal(q, q_ref)                      # port of quantize_cpu is bit-exact
    deq = orc.dequantize_cpu_port(code_after, q, am, bs)
    assert np.array_equal(deq.view(np.uint32), deq_ref.view(np.uint32))


@pytest.mark.parametrize("case,bs", [("bs64", 64), ("bs4096", 4096)])
def test_device_8bit_restatement_vs_reference_cpu_golden(tables, case, bs):
    """kQuantizeBlockwise<General8bit> (reciprocal-multiply + pivot search) vs the reference's
    quantize_cpu (divide + nearest): absmax identical, codes agree >= 99.9%, and the
    8-bit dequantize is bit-identical given the same codes."""
    g = np.load(os.path.join(GOLDEN, "ref_cpu_blockwise.npz"))
    A, q_ref, am_ref = g[case + "_A"], g[case + "_q"], g[case + "_absmax"]
    code = tables["dynamic_map"]
    q, am = orc.quantize_blockwise(A, "fp32", code, bs, "8bit")
    assert np.array_equal(am, am_ref)
    agree = np.mean(q == q_ref)
    assert agree >= 0.999, agree
    # disagreements are neighbours only
    assert np.max(np.abs(q.astype(int) - q_ref.astype(int))) <= 1
    deq = orc.dequantize_blockwise(q_ref, am_ref, A.size, "fp32", g[case + "_code_after"], bs, "8bit")
    assert np.array_equal(deq.view(np.uint32), g[case + "_deq"].view(np.uint32))


@pytest.mark.skipif(not os.path.isdir("/root/reference/sycl"), reason="reference sources absent")
def test_live_reference_cpu_ops_agree_with_port(tables):
    rng = np.random.RandomState(7)
    A = (rng.randn(64 * 64 + 5) * 0.02).astype(np.float32)
    q1, a1, _ = orc.quantize_cpu_reference(tables["dynamic_map"], A, 64)
    q2, a2, _ = orc.quantize_cpu_port(tables["dynamic_map"], A, 64)
    assert np.array_equal(q1, q2) and np.array_equal(a1, a2)


# ------------------------------------------------------------------ blockwise 4-bit properties
@pytest.mark.parametrize("qtype", ["nf4", "fp4"])
@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 127, 4096 + 31])
def test_4bit_roundtrip_is_idempotent(qtype, n):
    rng = np.random.RandomState(n)
    A = rng.randn(n).astype(np.float32)
    q, am = orc.quantize_blockwise(A, "fp32", None, 64, qtype)
    assert q.size == (n + 1) // 2 and am.size == (n + 63) // 64
    d = orc.dequantize_blockwise(q, am, n, "fp32", None, 64, qtype)
    q2, am2 = orc.quantize_blockwise(d, "fp32", None, 64, qtype)
    d2 = orc.dequantize_blockwise(q2, am2, n, "fp32", None, 64, qtype)
    if qtype == "nf4":
        assert np.array_equal(d.view(np.uint32), d2.view(np.uint32))
    else:  # fp4 has a -0 code (0b1000) that re-quantizes to +0
        assert np.array_equal(d, d2)
    if n:
        # the block maximum is always representable (code +-1.0)
        blk = np.abs(A[:64]).max()
        assert am[0] == blk and np.abs(d[:64]).max() == blk


def test_nf4_odd_tail_and_zero_block():
    A = np.zeros(64 + 3, np.float32)
    A[64:] = [1.0, -0.5, 0.25]
    q, am = orc.quantize_blockwise(A, "fp32", None, 64, "nf4")
    assert am[0] == 0.0 and np.all(q[:32] == 0)       # 0 * inf = NaN -> code 0 (both nibbles)
    assert q[32] == (15 << 4) | 2 and q[33] == (10 << 4) | 7   # odd tail: low nibble = code(0.0) = 7
    d = orc.dequantize_blockwise(q, am, A.size, "fp32", None, 64, "nf4")
    assert np.all(d[:64] == 0) and np.all(np.signbit(d[:64]))  # -1.0 * 0 = -0.0


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_16bit_inputs_equal_fp32_of_same_values(dtype):
    import torch
    t = torch.randn(1000, dtype=torch.float16 if dtype == "fp16" else torch.bfloat16)
    bits = t.view(torch.int16).numpy().view(np.uint16)
    q1, a1 = orc.quantize_blockwise(bits, dtype, None, 64, "nf4")
    q2, a2 = orc.quantize_blockwise(t.float().numpy(), "fp32", None, 64, "nf4")
    assert np.array_equal(q1, q2) and np.array_equal(a1, a2)
    d = orc.dequantize_blockwise(q1, a1, 1000, dtype, None, 64, "nf4")
    d32 = orc.dequantize_blockwise(q1, a1, 1000, "fp32", None, 64, "nf4")
    exp = torch.from_numpy(d32).to(t.dtype).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(d, exp)        # our RNE conversions == torch's


# ------------------------------------------------------------------ layouts
@pytest.mark.parametrize("fmt", ["col32", "col_turing", "col_ampere"])
@pytest.mark.parametrize("shape", [(1, 1), (7, 33), (32, 32), (40, 100), (64, 256), (129, 65)])
def test_layouts_are_permutations_and_match_blas_utils(fmt, shape):
    rows, cols = shape
    L = orc.lib()
    f = orc.FORMATS[fmt]
    size = orc.layout_size(fmt, rows, cols)
    out_rows = size // (((cols + 31) // 32) * 32)
    seen = set()
    for r in range(rows):
        for c in range(cols):
            o = L.orc_layout_offset(f, rows, r, c)
            assert 0 <= o < size and o not in seen
            seen.add(o)
            assert o == L.orc_layout_offset_blasutils(f, out_rows * 32, r, c)   # blas_utils.h:263-325
    A = np.random.RandomState(0).randint(-128, 128, size=shape).astype(np.int8)
    T = orc.transform(A, fmt)
    assert np.array_equal(orc.untransform(T, fmt, rows, cols), A)
    # transposed variant == layout of A^T
    assert np.array_equal(orc.transform(A, fmt, transpose=True), orc.transform(np.ascontiguousarray(A.T), fmt))


@pytest.mark.parametrize("fmtB", ["col_turing", "col_ampere"])
def test_igemmlt_layouts_equal_rowmajor_int_matmul(fmtB):
    rng = np.random.RandomState(3)
    m, n, k = 19, 45, 96
    A = rng.randint(-128, 128, (m, k)).astype(np.int8)
    B = rng.randint(-128, 128, (n, k)).astype(np.int8)
    C = orc.igemmlt_32(orc.transform(A, "col32"), orc.transform(B, fmtB), m, n, k, fmtB)
    Crow = orc.untransform(C, "col32", m, n)
    exp = A.astype(np.int64) @ B.astype(np.int64).T
    assert np.array_equal(Crow, exp.astype(np.int32))
    assert np.array_equal(orc.igemm_rowmajor(A, B), exp.astype(np.int32))
    idx = np.array([3, 17, 64, 95], np.int32)
    assert np.array_equal(orc.extract_outliers(orc.transform(B, fmtB), idx, n, k, fmtB), B[:, idx])


# ------------------------------------------------------------------ int8 double quant
def test_double_quant_against_numpy():
    rng = np.random.RandomState(5)
    rows, cols = 37, 300
    A = rng.randn(rows, cols).astype(np.float16)
    A[3, 7] = 8.0
    A[20, 290] = -6.0
    A[20, 5] = 7.5
    rs, cs, nnz = orc.get_col_row_stats(A, 6.0)
    Af = np.abs(A.astype(np.float32))
    Az = np.where(Af >= 6.0, 0, Af)
    assert np.array_equal(rs, Az.max(1)) and np.array_equal(cs, Az.max(0))
    assert nnz.sum() == 3 and nnz[0] == 0
    ptr = np.cumsum(nnz).astype(np.int32)
    out_row, out_col, ri, ci, val = orc.double_rowcol_quant(A, rs, cs, ptr, 6.0)
    assert sorted(zip(ri.tolist(), ci.tolist())) == [(3, 7), (20, 5), (20, 290)]
    exp_row = np.rint(A.astype(np.float32) * (np.float32(127.0) / rs)[:, None]).astype(np.int32)
    exp_row[Af >= 6.0] = 0
    assert np.array_equal(out_row, exp_row.astype(np.int8))
    exp_col = np.clip(np.rint(A.astype(np.float32) * (np.float32(127.0) / cs)[None, :]), -128, 127)
    assert np.array_equal(out_col, exp_col.astype(np.int8))
    # threshold 0: plain absmax, no COO
    rs0, cs0, nnz0 = orc.get_col_row_stats(A, 0.0)
    assert nnz0 is None and np.array_equal(rs0, Af.max(1))


# ------------------------------------------------------------------ gemv chains
def test_gemv_chains_close_to_exact():
    import torch
    torch.manual_seed(0)
    N, K = 48, 1024
    W = (torch.randn(N, K) * 0.02)
    x = torch.randn(K).to(torch.bfloat16)
    q, am = orc.quantize_blockwise(W.numpy().ravel(), "fp32", None, 64, "nf4")
    code = orc.nf4_table()
    xb = x.view(torch.int16).numpy().view(np.uint16)
    exact = orc.gemm_4bit_exact(xb, "bf16", q, am, code, 1, N, K)[0]
    for mode, tol in ((0, 8e-3), (1, 5e-3)):
        y = orc.gemv_4bit(xb, "bf16", q, am, code, N, K, 64, mode)
        yf = torch.from_numpy(y.view(np.int16)).view(torch.bfloat16).double().numpy()
        rel = np.linalg.norm(yf - exact) / np.linalg.norm(exact)
        assert rel < tol, (mode, rel)
    y32 = orc.gemv_4bit(x.float().numpy(), "fp32", q, am, code, N, K, 64, 0)
    assert np.linalg.norm(y32 - exact) / np.linalg.norm(exact) < 1e-5


def test_llm_int8_forward_restatement_tracks_the_fp32_product():
    """oracle.llm_int8_forward (reference _functions.py:292-434 composed from the kernel restatements): outlier columns
    are found, zeroed in CA and carried in 16 bit; the result tracks the fp32 product within int8 quantisation noise."""
    rng = np.random.default_rng(3)
    m, k, n = 48, 256, 64
    A = rng.standard_normal((m, k)).astype(np.float16)
    A[:, 17] = 8.0
    A[5, 100] = -7.5
    W = (rng.standard_normal((n, k)) * 0.05).astype(np.float16)
    rs, cs, _ = orc.get_col_row_stats(W.T.copy().T, 0.0)        # row stats of W = SCB
    CB, _, _, _, _ = orc.double_rowcol_quant(W, rs, cs)
    bias = (rng.standard_normal(n) * 0.1).astype(np.float16)
    y, CA, SCA, idx = orc.llm_int8_forward(A, CB, rs, bias, 6.0)
    assert idx.tolist() == [17, 100]
    assert not CA[:, 17].any() and not CA[:, 100].any()
    ref = A.astype(np.float32) @ W.astype(np.float32).T + bias.astype(np.float32)
    err = np.linalg.norm(y.astype(np.float32) - ref) / np.linalg.norm(ref)
    assert err < 0.02, err
    # threshold 0: no decomposition, plain int8 path
    y0, CA0, _, idx0 = orc.llm_int8_forward(A, CB, rs, bias, 0.0)
    assert idx0.size == 0 and CA0[:, 17].any()


def test_scalar_trees_match_reference_executed_vectors(tables):
    """tests/golden/ref_device_trees.npz holds outputs of the reference's OWN scalar device functions (bodies cut out
    of kernel_quant.cpp:519-837 verbatim and compiled by make_golden.py) on ~260 k seeded inputs plus every decision
    threshold +- 1 ulp, +-0, denormals, inf and NaN.  The oracle's restatements must agree on every input -- this is
    what pins the NF4 / FP4 trees and dQuantize<0> beyond their constants."""
    import ctypes as ct
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mkg", os.path.join(GOLDEN, "make_golden.py"))
    mkg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mkg)
    x, code = mkg.tree_inputs()
    g = np.load(os.path.join(GOLDEN, "ref_device_trees.npz"))
    assert x.size == int(g["n_inputs"][0]) and int(np.bitwise_xor.reduce(x.view(np.uint32))) == int(g["x_crc"][0])
    L = orc.lib()
    L.orc_quantize_nf4_scalar.restype = ct.c_ubyte
    L.orc_quantize_fp4_scalar.restype = ct.c_ubyte
    L.orc_quantize_8bit_scalar.restype = ct.c_ubyte
    L.orc_quantize_nf4_scalar.argtypes = [ct.c_float]
    L.orc_quantize_fp4_scalar.argtypes = [ct.c_float]
    L.orc_quantize_8bit_scalar.argtypes = [ct.c_void_p, ct.c_float]
    step = 1 if os.environ.get("BNB_FULL_TREE_CHECK") else 7          # every 7th random input + ALL the specials
    n_special = x.size - 260000
    idx = np.concatenate([np.arange(0, 260000, step), np.arange(260000, x.size)])
    assert n_special > 1500
    q_nf4 = np.array([L.orc_quantize_nf4_scalar(float(v)) for v in x[idx]], np.uint8)
    q_fp4 = np.array([L.orc_quantize_fp4_scalar(float(v)) for v in x[idx]], np.uint8)
    assert np.array_equal(q_nf4, g["q_nf4"][idx])
    assert np.array_equal(q_fp4, g["q_fp4"][idx])
    finite = np.isfinite(x)
    pos = np.cumsum(finite) - 1                                       # index into the finite-only 8-bit fixture
    idx8 = idx[finite[idx]]
    cptr = code.ctypes.data_as(ct.c_void_p)
    q8 = np.array([L.orc_quantize_8bit_scalar(cptr, float(v)) for v in x[idx8]], np.uint8)
    assert np.array_equal(q8, g["q_8bit_finite"][pos[idx8]])
    assert np.array_equal(orc.nf4_table().view(np.uint32), g["deq_nf4"].view(np.uint32))
    fp4 = np.array([orc.fp4_table()[i] * np.float32(0.73) for i in range(16)], np.float32)
    # dDequantizeFP4Tree multiplies (c * absmax) * sign left to right: the table at absmax 1 times 0.73 is the same product
    assert np.array_equal(fp4.view(np.uint32), g["deq_fp4_absmax0p73"].view(np.uint32))


def test_llm_int8_backward_restatement_tracks_the_fp64_gradients():
    """oracle.llm_int8_backward (reference _functions.py:436-483) on both weight modes: int8 quantisation noise only
    against the exact gradients, and the outlier columns of grad_B carried in 16 bit."""
    rng = np.random.RandomState(3)
    m, k, n = 64, 128, 96
    g = rng.randn(m, n).astype(np.float16)
    x = rng.randn(m, k).astype(np.float16)
    x[:, 7] = 8.0
    W = (rng.randn(n, k) * 0.05).astype(np.float16)
    gA, gB = orc.llm_int8_backward(g, x, 6.0, W_f16=W, need_grad_B=True)
    ref_B = g.astype(np.float64).T @ x.astype(np.float64)
    ref_A = g.astype(np.float64) @ W.astype(np.float64)
    assert np.linalg.norm(gB - ref_B) / np.linalg.norm(ref_B) < 0.02
    assert np.linalg.norm(gA - ref_A) / np.linalg.norm(ref_A) < 0.02
    assert np.linalg.norm(gB[:, 7] - ref_B[:, 7]) / np.linalg.norm(ref_B[:, 7]) < 2e-3     # outlier column: 16-bit product
    rs, cs, _ = orc.get_col_row_stats(W, 0.0)
    CB, _, _, _, _ = orc.double_rowcol_quant(W, rs, cs)
    gA2, none = orc.llm_int8_backward(g, x, 6.0, CB=CB, SCB=rs)
    assert none is None and np.linalg.norm(gA2 - ref_A) / np.linalg.norm(ref_A) < 0.02
