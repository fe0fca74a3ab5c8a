"""ctypes/numpy front-end of the CPU oracle (oracle/bnb_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

16-bit floats travel as ``np.uint16`` bit patterns plus a dtype tag ("fp16" / "bf16"), so the
oracle never depends on numpy's (absent) bfloat16 support.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbnb_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")

DTYPES = {"fp32": 0, "fp16": 1, "bf16": 2}
QTYPES = {"8bit": 0, "fp4": 1, "nf4": 2}
FORMATS = {"col32": 0, "col_turing": 1, "col_ampere": 2}


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "bnb_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "_build/libbnb_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/sycl") and (force or not os.path.exists(_REF_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib() -> ct.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ct.CDLL(_LIB_PATH)
        _lib.orc_layout_offset.restype = ct.c_long
        _lib.orc_layout_offset_blasutils.restype = ct.c_long
        _lib.orc_layout_size.restype = ct.c_long
        _lib.orc_round_to_dtype.restype = ct.c_float
        _lib.orc_round_to_dtype.argtypes = [ct.c_float, ct.c_int]
        _lib.orc_quantize_nf4_scalar.restype = ct.c_ubyte
        _lib.orc_quantize_nf4_scalar.argtypes = [ct.c_float]
        _lib.orc_quantize_fp4_scalar.restype = ct.c_ubyte
        _lib.orc_quantize_fp4_scalar.argtypes = [ct.c_float]
        _lib.orc_layout_offset_blasutils.argtypes = [ct.c_int, ct.c_long, ct.c_int, ct.c_int]
    return _lib


def ref_available() -> bool:
    return os.path.exists(_REF_PATH) or os.path.isdir("/root/reference/sycl")


def ref_lib() -> ct.CDLL:
    """The reference's own cpu_ops.cpp compiled into oracle/_ref (kind == "reference")."""
    global _ref
    if _ref is None:
        build()
        _ref = ct.CDLL(_REF_PATH)
    return _ref


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle expects contiguous arrays"
    return a.ctypes.data_as(ct.c_void_p)


def _np_in(A: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "fp32":
        return np.ascontiguousarray(A, dtype=np.float32)
    if A.dtype == np.float16:
        A = A.view(np.uint16)
    return np.ascontiguousarray(A, dtype=np.uint16)


# ------------------------------------------------------------------------------------ tables
def nf4_table() -> np.ndarray:
    t = np.zeros(16, np.float32)
    lib().orc_nf4_table(_p(t))
    return t


def fp4_table() -> np.ndarray:
    t = np.zeros(16, np.float32)
    lib().orc_fp4_table(_p(t))
    return t


# ------------------------------------------------------------------------------------ a1 / a2
def quantize_blockwise(A: np.ndarray, dtype: str = "fp32", code: Optional[np.ndarray] = None,
                       blocksize: int = 64, qtype: str = "nf4") -> Tuple[np.ndarray, np.ndarray]:
    A = _np_in(A, dtype)
    n = A.size
    nblocks = (n + blocksize - 1) // blocksize
    absmax = np.zeros(nblocks, np.float32)
    out = np.zeros(n if qtype == "8bit" else (n + 1) // 2, np.uint8)
    if code is not None:
        code = np.ascontiguousarray(code, np.float32)
    lib().orc_quantize_blockwise(_p(code), _p(A), DTYPES[dtype], _p(absmax), _p(out),
                                 ct.c_int(blocksize), ct.c_long(n), QTYPES[qtype])
    return out, absmax


def dequantize_blockwise(Q: np.ndarray, absmax: np.ndarray, n: int, out_dtype: str = "fp32",
                         code: Optional[np.ndarray] = None, blocksize: int = 64,
                         qtype: str = "nf4") -> np.ndarray:
    Q = np.ascontiguousarray(Q, np.uint8)
    absmax = np.ascontiguousarray(absmax, np.float32)
    out = np.zeros(n, np.float32 if out_dtype == "fp32" else np.uint16)
    if code is not None:
        code = np.ascontiguousarray(code, np.float32)
    lib().orc_dequantize_blockwise(_p(code), _p(Q), _p(absmax), _p(out), DTYPES[out_dtype],
                                   ct.c_int(blocksize), ct.c_long(n), QTYPES[qtype])
    return out


def denest_absmax(qabsmax: np.ndarray, absmax2: np.ndarray, code2: np.ndarray, offset: float,
                  blocksize2: int = 256) -> np.ndarray:
    qabsmax = np.ascontiguousarray(qabsmax, np.uint8)
    out = np.zeros(qabsmax.size, np.float32)
    lib().orc_denest_absmax(_p(np.ascontiguousarray(code2, np.float32)), _p(qabsmax),
                            _p(np.ascontiguousarray(absmax2, np.float32)),
                            ct.c_float(float(offset)), _p(out), ct.c_int(blocksize2),
                            ct.c_long(qabsmax.size))
    return out


# ------------------------------------------------------------------------------------ a3 / a4
def gemv_4bit(A: np.ndarray, dtype: str, B: np.ndarray, absmax: np.ndarray, code: np.ndarray,
              N: int, K: int, blocksize: int = 64, mode: int = 0) -> np.ndarray:
    """mode 0: reference-faithful T-arithmetic chain; mode 1: fp32-product chain."""
    A = _np_in(A, dtype)
    out = np.zeros(N, np.float32 if dtype == "fp32" else np.uint16)
    lib().orc_gemv_4bit(ct.c_int(N), ct.c_int(K), _p(A), _p(np.ascontiguousarray(B, np.uint8)),
                        _p(np.ascontiguousarray(absmax, np.float32)),
                        _p(np.ascontiguousarray(code, np.float32)), _p(out), DTYPES[dtype],
                        ct.c_int((K + 1) // 2), ct.c_int(blocksize), ct.c_int(mode))
    return out


def gemm_4bit_exact(A: np.ndarray, dtype: str, B: np.ndarray, absmax: np.ndarray,
                    code: np.ndarray, batch: int, N: int, K: int, blocksize: int = 64) -> np.ndarray:
    A = _np_in(A, dtype)
    out = np.zeros((batch, N), np.float64)
    lib().orc_gemm_4bit_exact(ct.c_int(batch), ct.c_int(N), ct.c_int(K), _p(A), DTYPES[dtype],
                              _p(np.ascontiguousarray(B, np.uint8)),
                              _p(np.ascontiguousarray(absmax, np.float32)),
                              _p(np.ascontiguousarray(code, np.float32)), _p(out),
                              ct.c_int(blocksize))
    return out


def gemm_4bit_dequant_ref(A: np.ndarray, dtype: str, B: np.ndarray, absmax: np.ndarray,
                          code: np.ndarray, batch: int, N: int, K: int,
                          blocksize: int = 64) -> np.ndarray:
    A = _np_in(A, dtype)
    out = np.zeros((batch, N), np.float32 if dtype == "fp32" else np.uint16)
    lib().orc_gemm_4bit_dequant_ref(ct.c_int(batch), ct.c_int(N), ct.c_int(K), _p(A),
                                    DTYPES[dtype], _p(np.ascontiguousarray(B, np.uint8)),
                                    _p(np.ascontiguousarray(absmax, np.float32)),
                                    _p(np.ascontiguousarray(code, np.float32)), _p(out),
                                    ct.c_int(blocksize))
    return out


# ------------------------------------------------------------------------------------ a5 / a6
def get_col_row_stats(A_f16: np.ndarray, threshold: float = 0.0):
    """A_f16: [rows, cols] float16.  Returns (row_stats, col_stats, nnz_count_row | None);
    nnz_count_row is the raw per-(tile,row) count array (the caller cumsums it, functional.py:2432)."""
    A = _np_in(A_f16, "fp16")
    rows, cols = A.shape
    row_stats = np.full(rows, -50000.0, np.float32)
    col_stats = np.full(cols, -50000.0, np.float32)
    nnz = None
    if threshold > 0.0:
        col_tiles = (cols + 255) // 256
        tiled_rows = ((rows + 15) // 16) * 16
        nnz = np.zeros(tiled_rows * col_tiles + 1, np.int32)
    lib().orc_get_col_row_stats(_p(A), _p(row_stats), _p(col_stats), _p(nnz),
                                ct.c_float(threshold), ct.c_int(rows), ct.c_int(cols))
    return row_stats, col_stats, nnz


def double_rowcol_quant(A_f16: np.ndarray, row_stats: np.ndarray, col_stats: np.ndarray,
                        nnz_block_ptr: Optional[np.ndarray] = None, threshold: float = 0.0):
    A = _np_in(A_f16, "fp16")
    rows, cols = A.shape
    out_row = np.zeros((rows, cols), np.int8)
    out_col = np.zeros((rows, cols), np.int8)
    rowidx = colidx = val = None
    if threshold > 0.0 and nnz_block_ptr is not None:
        nnz = int(nnz_block_ptr[-1])
        rowidx = np.zeros(nnz, np.int32)
        colidx = np.zeros(nnz, np.int32)
        val = np.zeros(nnz, np.uint16)
        nnz_block_ptr = np.ascontiguousarray(nnz_block_ptr, np.int32)
    lib().orc_double_rowcol_quant(_p(A), _p(np.ascontiguousarray(row_stats, np.float32)),
                                  _p(np.ascontiguousarray(col_stats, np.float32)), _p(out_col),
                                  _p(out_row), _p(rowidx), _p(colidx), _p(val),
                                  _p(nnz_block_ptr) if rowidx is not None else None,
                                  ct.c_float(threshold if rowidx is not None else 0.0),
                                  ct.c_int(rows), ct.c_int(cols))
    return out_row, out_col, rowidx, colidx, val


# ------------------------------------------------------------------------------------ a7
def layout_size(fmt: str, rows: int, cols: int) -> int:
    return int(lib().orc_layout_size(FORMATS[fmt], ct.c_int(rows), ct.c_int(cols)))


def transform(A: np.ndarray, fmt: str, transpose: bool = False) -> np.ndarray:
    A = np.ascontiguousarray(A)
    assert A.dtype in (np.int8, np.int32)
    rows, cols = A.shape
    R, C = (cols, rows) if transpose else (rows, cols)
    out = np.zeros(layout_size(fmt, R, C), A.dtype)
    lib().orc_transform_row2fmt(_p(A), _p(out), ct.c_int(rows), ct.c_int(cols), FORMATS[fmt],
                                ct.c_int(1 if transpose else 0), ct.c_int(A.dtype.itemsize))
    return out


def untransform(Af: np.ndarray, fmt: str, rows: int, cols: int) -> np.ndarray:
    Af = np.ascontiguousarray(Af)
    out = np.zeros((rows, cols), Af.dtype)
    lib().orc_transform_fmt2row(_p(Af), _p(out), ct.c_int(rows), ct.c_int(cols), FORMATS[fmt],
                                ct.c_int(Af.dtype.itemsize))
    return out


# ------------------------------------------------------------------------------------ a8..a10
def igemmlt_32(A_col32: np.ndarray, B_fmt: np.ndarray, m: int, n: int, k: int, fmtB: str) -> np.ndarray:
    C = np.zeros(layout_size("col32", m, n), np.int32)
    lib().orc_igemmlt_32(ct.c_int(m), ct.c_int(n), ct.c_int(k),
                         _p(np.ascontiguousarray(A_col32, np.int8)),
                         _p(np.ascontiguousarray(B_fmt, np.int8)), _p(C), FORMATS[fmtB])
    return C


def igemmlt_8(A_col32: np.ndarray, B_fmt: np.ndarray, m: int, n: int, k: int, fmtB: str,
              row_scale: Optional[np.ndarray] = None) -> np.ndarray:
    """int8-output igemmlt (reference op_gemm.cpp:604-638 -> pythonInterface.cpp:303-316 cigemmlt_<fmt>_8 /
    _8_rowscale): the int32 accumulator is scaled in fp32 -- alpha = 1.0f, or the per-row vector row_scale[i]
    (matmul_desc pointer_mode = alpha vector, :624-632) -- rounded to nearest-even and saturated to int8; C is col32.
    The arithmetic lives in oneDNN / cublasLt (not in /root/reference): parity unpinned by the reference's tests;
    this is the documented semantics of an int8 D with an fp32 scale."""
    acc = untransform(igemmlt_32(A_col32, B_fmt, m, n, k, fmtB), "col32", m, n).astype(np.float32)
    alpha = np.ones((m, 1), np.float32) if row_scale is None else np.asarray(row_scale, np.float32).reshape(m, 1)
    q = np.clip(np.rint(acc * alpha), -128, 127).astype(np.int8)
    return transform(q, "col32")


def igemm_rowmajor(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    A = np.ascontiguousarray(A, np.int8)
    B = np.ascontiguousarray(B, np.int8)
    m, k = A.shape
    n = B.shape[0]
    C = np.zeros((m, n), np.int32)
    lib().orc_igemm_rowmajor(ct.c_int(m), ct.c_int(n), ct.c_int(k), _p(A), _p(B), _p(C))
    return C


def mm_dequant(C: np.ndarray, row_stats: np.ndarray, col_stats: np.ndarray, rows: int, cols: int,
               bias_f16: Optional[np.ndarray] = None, col32: bool = True) -> np.ndarray:
    out = np.zeros((rows, cols), np.uint16)
    b = None if bias_f16 is None else _np_in(bias_f16, "fp16")
    lib().orc_dequant_mm_int32_fp16(_p(np.ascontiguousarray(C, np.int32)),
                                    _p(np.ascontiguousarray(row_stats, np.float32)),
                                    _p(np.ascontiguousarray(col_stats, np.float32)), _p(out),
                                    _p(b), ct.c_int(rows), ct.c_int(cols),
                                    ct.c_int(1 if col32 else 0))
    return out.view(np.float16)


# ------------------------------------------------------------------------------------ a12 (orchestration)
def llm_int8_forward(A_f16: np.ndarray, CB: np.ndarray, SCB: np.ndarray, bias_f16: Optional[np.ndarray],
                     threshold: float):
    """CPU restatement of MatMul8bitLt.forward with has_fp16_weights=False (reference
    python_src_quants/autograd/_functions.py:292-434), composed from the per-kernel restatements above:
      :340      CA, CAt, SCA, SCAt, coo = double_quant(A, threshold)                  -> get_col_row_stats + double_rowcol_quant
      :369-372  idx = unique(coo.colidx)                                               (outlier feature columns)
      :381      subB = (extract_outliers(CxB)[..idx] * SCB.view(-1,1) / 127).t().to(fp16)
      :382-384  CA[:, idx] = 0 ; subA = A[:, idx]
      :404-414  out32 = igemmlt(CA, CB) ; output = mm_dequant(out32, SCA, SCB, bias)
      :431      output += matmul(subA, subB)       (fp16 operands, fp32 accumulation, fp16 result, fp16 add)
    Returns (y float16 [m, n], CA int8 [m, k] with outlier columns zeroed, SCA float32 [m], idx int32 ascending)."""
    A = np.ascontiguousarray(A_f16, np.float16)
    m, k = A.shape
    n = CB.shape[0]
    row_stats, col_stats, nnz = get_col_row_stats(A, threshold)
    if threshold > 0.0:
        ptr = np.cumsum(nnz, dtype=np.int64).astype(np.int32)      # functional.py:2432
        CA, _, _, colidx, _ = double_rowcol_quant(A, row_stats, col_stats, ptr, threshold)
        idx = np.unique(colidx).astype(np.int32) if colidx is not None and colidx.size else np.zeros(0, np.int32)
    else:
        CA, _, _, _, _ = double_rowcol_quant(A, row_stats, col_stats)
        idx = np.zeros(0, np.int32)
    CA = CA.copy()
    if idx.size:
        CA[:, idx] = 0
    acc = igemm_rowmajor(CA, CB)
    y = mm_dequant(acc, row_stats, np.ascontiguousarray(SCB, np.float32), m, n, bias_f16, col32=False)
    if idx.size:
        subA = A[:, idx].astype(np.float32)
        # fp32 product then one rounding to fp16, like the device expression (outliers * SCB / 127).to(fp16)
        subB = ((CB[:, idx].astype(np.float32) * np.asarray(SCB, np.float32)[:, None]) / np.float32(127.0)).astype(np.float16)
        side = (subA @ subB.astype(np.float32).T).astype(np.float16)      # fp32 accumulation, fp16 result
        y = (y.astype(np.float32) + side.astype(np.float32)).astype(np.float16)
    return y, CA, row_stats, idx


def llm_int8_backward(grad_f16: np.ndarray, A_f16: np.ndarray, threshold: float, CB: Optional[np.ndarray] = None,
                      SCB: Optional[np.ndarray] = None, W_f16: Optional[np.ndarray] = None, need_grad_B: bool = False):
    """CPU restatement of MatMul8bitLt.backward (reference python_src_quants/autograd/_functions.py:436-483), composed
    from the per-kernel restatements.  grad [m, n] fp16, A [m, k] fp16 (the forward's input).
      :454      Cgrad, Cgradt, SCgrad, SCgradt = double_quant(grad)                      (no threshold)
      :455-461  grad_B = mm_dequant(igemmlt(Cgradt^T, CAt^T), SCgradt, SCAt) (+ grad^T @ subA on the outlier columns):
                CAt / SCAt are the forward's column-quantised activations with the outlier columns zeroed (:382-383)
      :462-468  has_fp16_weights (CBt from double_quant(W)):  grad_A = mm_dequant(igemmlt(Cgrad, CBt^T), SCgrad, SCBt)
      :470-472  frozen int8 weight (CB, SCB):  grad_A = grad @ (CB.to(fp16) * (SCB / 127))     (16-bit GEMM)
    Returns (grad_A fp16 [m, k], grad_B fp16 [n, k] | None).  The int8 parts are exact; the two 16-bit GEMMs are
    restated with fp32 accumulation and one rounding (the BLAS's summation order is not specified)."""
    G = np.ascontiguousarray(grad_f16, np.float16)
    A = np.ascontiguousarray(A_f16, np.float16)
    m, n = G.shape
    k = A.shape[1]
    g_rs, g_cs, _ = get_col_row_stats(G, 0.0)
    Cgrad, Cgradt, _, _, _ = double_rowcol_quant(G, g_rs, g_cs)
    grad_B = None
    if need_grad_B:
        a_rs, a_cs, nnz = get_col_row_stats(A, threshold)
        if threshold > 0.0:
            ptr = np.cumsum(nnz, dtype=np.int64).astype(np.int32)
            _, CAt, _, colidx, _ = double_rowcol_quant(A, a_rs, a_cs, ptr, threshold)
            idx = np.unique(colidx).astype(np.int64) if colidx is not None and colidx.size else np.zeros(0, np.int64)
        else:
            _, CAt, _, _, _ = double_rowcol_quant(A, a_rs, a_cs)
            idx = np.zeros(0, np.int64)
        CAt = CAt.copy()
        if idx.size:
            CAt[:, idx] = 0
        acc = igemm_rowmajor(np.ascontiguousarray(Cgradt.T), np.ascontiguousarray(CAt.T))          # [n, k]
        grad_B = mm_dequant(acc, g_cs, a_cs, n, k, None, col32=False)
        if idx.size:
            side = (G.astype(np.float32).T @ A[:, idx].astype(np.float32)).astype(np.float16)
            grad_B[:, idx] = (grad_B[:, idx].astype(np.float32) + side.astype(np.float32)).astype(np.float16)
    if W_f16 is not None:            # has_fp16_weights: CBt = column-quantised W
        W = np.ascontiguousarray(W_f16, np.float16)
        w_rs, w_cs, _ = get_col_row_stats(W, 0.0)
        _, CBt, _, _, _ = double_rowcol_quant(W, w_rs, w_cs)
        acc = igemm_rowmajor(Cgrad, np.ascontiguousarray(CBt.T))                                   # [m, k]
        grad_A = mm_dequant(acc, g_rs, w_cs, m, k, None, col32=False)
    else:
        # CB.to(fp16).mul_(SCB.unsqueeze(1).mul(1/127)): the product is formed in fp32 and rounded to fp16 once
        Wd = (CB.astype(np.float32) * (np.asarray(SCB, np.float32)[:, None] * np.float32(1.0 / 127.0))).astype(np.float16)
        grad_A = (G.astype(np.float32) @ Wd.astype(np.float32)).astype(np.float16)
    return grad_A, grad_B


def extract_outliers(A_fmt: np.ndarray, idx: np.ndarray, rows: int, cols: int, fmt: str) -> np.ndarray:
    idx = np.ascontiguousarray(idx, np.int32)
    out = np.zeros((rows, idx.size), np.int8)
    lib().orc_extract_outliers(_p(np.ascontiguousarray(A_fmt, np.int8)), _p(idx), _p(out),
                               ct.c_int(idx.size), ct.c_int(rows), ct.c_int(cols), FORMATS[fmt])
    return out


# ------------------------------------------------------------------------------------ CPU 8-bit path
def quantize_cpu_port(code: np.ndarray, A: np.ndarray, blocksize: int):
    """Our restatement of sycl/cpu_ops.cpp quantize_cpu (kind == "port")."""
    code = np.array(code, np.float32)  # private copy: quantize_cpu mutates code[0]
    A = np.ascontiguousarray(A, np.float32).ravel()
    n = A.size
    absmax = np.zeros((n + blocksize - 1) // blocksize, np.float32)
    out = np.zeros(n, np.uint8)
    lib().orc_quantize_cpu(_p(code), _p(A), _p(absmax), _p(out), ct.c_longlong(blocksize),
                           ct.c_longlong(n))
    return out, absmax, code


def dequantize_cpu_port(code: np.ndarray, Q: np.ndarray, absmax: np.ndarray, blocksize: int):
    Q = np.ascontiguousarray(Q, np.uint8).ravel()
    out = np.zeros(Q.size, np.float32)
    lib().orc_dequantize_cpu(_p(np.ascontiguousarray(code, np.float32)), _p(Q),
                             _p(np.ascontiguousarray(absmax, np.float32)), _p(out),
                             ct.c_longlong(blocksize), ct.c_longlong(Q.size))
    return out


def quantize_cpu_reference(code: np.ndarray, A: np.ndarray, blocksize: int):
    """The reference's own quantize_cpu (oracle/_ref), one std::thread per block."""
    code = np.array(code, np.float32)
    A = np.ascontiguousarray(A, np.float32).ravel()
    n = A.size
    absmax = np.zeros((n + blocksize - 1) // blocksize, np.float32)
    out = np.zeros(n, np.uint8)
    ref_lib().cquantize_blockwise_cpu_fp32(_p(code), _p(A), _p(absmax), _p(out),
                                           ct.c_longlong(blocksize), ct.c_longlong(n))
    return out, absmax, code


def dequantize_cpu_reference(code: np.ndarray, Q: np.ndarray, absmax: np.ndarray, blocksize: int):
    Q = np.ascontiguousarray(Q, np.uint8).ravel()
    out = np.zeros(Q.size, np.float32)
    ref_lib().cdequantize_blockwise_cpu_fp32(_p(np.ascontiguousarray(code, np.float32)), _p(Q),
                                             _p(np.ascontiguousarray(absmax, np.float32)), _p(out),
                                             ct.c_longlong(blocksize), ct.c_longlong(Q.size))
    return out
