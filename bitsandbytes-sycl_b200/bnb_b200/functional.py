# Derived from python_src_quants/functional.py of abhilash1910/bitsandbytes-SYCL (itself bitsandbytes,
# Copyright (c) Facebook, Inc. and its affiliates, MIT license -- see the LICENSE file of that repository).
# This file keeps the reference's public interface (function names, argument meaning, asserts, QuantState.as_dict/from_dict wire format, create_dynamic_map) because it is
# the wire / API contract of the drop-in; "xpu" became "cuda" and everything below the interface calls the native
# sm_100a library.  It is a derived host-side shim, not from-scratch work -- the from-scratch work is csrc/.
"""Host-side mirror of the reference's python_src_quants/functional.py for the quantized-linear hot path.

Same function names, argument meaning and error behaviour as the reference (file:line cited per function,
relative to /root/reference/python_src_quants/functional.py); the device is "cuda" instead of "xpu" and
every op is a ctypes call into libbitsandbytes_b200.so (hand-written sm_100a kernels).  PyTorch is used
only for allocation, streams and the few host-side reductions the reference also does in torch
(absmax.mean(), cumsum, sort, unique).  There is no CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

import ctypes as ct
import os
from functools import reduce
from typing import Any, Dict, Optional, Tuple

import torch
from torch import Tensor

from .cextension import lib
from .utils import pack_dict_to_tensor, unpack_tensor_to_dict

name2qmap: Dict[str, Tensor] = {}

dtype2bytes = {torch.float32: 4, torch.float16: 2, torch.bfloat16: 2, torch.uint8: 1, torch.int8: 1}

# B200-native switches (additive; defaults pick the fast path, the reference-shaped path stays reachable)
FUSED_NESTED_GEMV = os.environ.get("BNB_B200_FUSED_NESTED_GEMV", "1") != "0"
FUSED_INT8_LINEAR = os.environ.get("BNB_B200_FUSED_INT8_LINEAR", "1") != "0"
FUSED_GEMM_4BIT = os.environ.get("BNB_B200_FUSED_GEMM_4BIT", "1") != "0"
INT8_LAYOUT = os.environ.get("BNB_B200_INT8_LAYOUT", "row")  # "row" (native) | "col_turing" | "col_ampere"


def prod(iterable):
    return reduce(lambda a, b: a * b, iterable, 1)


# ------------------------------------------------------------------------------------------------
# code books
# ------------------------------------------------------------------------------------------------
def create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    """Dynamic 8-bit quantisation map (reference :339-391): for every exponent e in [-(E-1), 0] a linear
    grid of fraction means scaled by 10**e, plus 0 and 1.0, sorted.  Restated, same float arithmetic
    (torch.linspace in fp32, python-float scaling) so the 256 values are bit-identical."""
    data = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    for i in range(max_exponent_bits):
        n_frac = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                     else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        edges = torch.linspace(0.1, 1, n_frac)
        means = (edges[:-1] + edges[1:]) / 2.0
        scale = 10 ** (-(max_exponent_bits - 1) + i)
        data += (scale * means).tolist()
        if signed:
            data += (-scale * means).tolist()
    if additional_items > 0:
        edges = torch.linspace(0.1, 1, additional_items + 1)
        means = (edges[:-1] + edges[1:]) / 2.0
        scale = 10 ** (-(max_exponent_bits - 1) + i)
        data += (scale * means).tolist()
        if signed:
            data += (-scale * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) == 2 ** total_bits
    data += [0] * (256 - len(data))
    data.sort()
    return Tensor(data)


_NF4 = [-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
        -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
        0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
        0.7229568362236023, 1.0]
_FP4 = [0, 0.0625, 8.0, 12.0, 4.0, 6.0, 2.0, 3.0, -0, -0.0625, -8.0, -12.0, -4.0, -6.0, -2.0, -3.0]


_code_cache: Dict[Tuple[str, str], Tensor] = {}


def get_4bit_type(typename, device=None, blocksize=64):
    """16-entry code of a 4-bit type, normalised to absmax 1 (reference :1020-1099).  The device tensor is
    built once per (type, device) and shared (read-only) -- no host->device copy per quantize call."""
    if device is None:
        device = "cuda"
    key = (typename, str(torch.device(device)) if not isinstance(device, torch.device) else str(device))
    cached = _code_cache.get(key)
    if cached is not None:
        return cached
    if typename == "nf4":
        data = _NF4
    elif typename == "fp4":
        data = _FP4
    elif typename == "int4":
        data = [7, 6, 5, 4, 3, 2, 1, 0, -0, -1, -2, -3, -4, -5, -6, -7]
    else:
        raise NotImplementedError(f"Typename {typename} not supported")
    data = torch.tensor(data, dtype=torch.float32, device=device)
    data.div_(data.abs().max())
    assert data.numel() == 16
    _code_cache[key] = data
    return data


def get_special_format_str():
    """Weight layout the int8 GEMM wants (reference :410-418 returns an IMMA layout by compute
    capability).  sm_100a's tcgen05 + TMA read K-major row-major int8 directly, so the native answer is
    "row"; BNB_B200_INT8_LAYOUT=col_turing|col_ampere keeps the reference's layouts (slower ABI path)."""
    return INT8_LAYOUT


# ------------------------------------------------------------------------------------------------
# call plumbing
# ------------------------------------------------------------------------------------------------
def is_on_gpu(tensors):
    """reference :421-439 with "xpu" -> "cuda"."""
    gpu_ids = set()
    for t in tensors:
        if t is None:
            continue
        if t.device.type != "cuda":
            raise TypeError(
                "All input tensors need to be on the same GPU, but found some tensors to not be on a GPU:\n"
                f" {[(t.shape, t.device) for t in tensors if t is not None]}")
        gpu_ids.add(t.device.index)
    if len(gpu_ids) > 1:
        raise TypeError(
            "Input tensors need to be on the same GPU, but found the following tensor and device combinations:\n"
            f" {[(t.shape, t.device) for t in tensors if t is not None]}")
    return True


def get_ptr(A: Optional[Tensor]) -> Optional[ct.c_void_p]:
    """reference :442-458."""
    if A is None:
        return None
    return ct.c_void_p(A.data_ptr())


def pre_call(device):
    """reference :461-464, plus: hand the caller's CURRENT stream to the library (the ABI has no slot)."""
    prev_device = torch.cuda.current_device()
    torch.cuda.set_device(device)
    lib.cbnb_set_stream(ct.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    return prev_device


def post_call(prev_device):
    """reference :467-468, plus: surface a latched CUDA error as an exception (fail loudly)."""
    err = lib.cbnb_last_error()
    torch.cuda.set_device(prev_device)
    if err != 0:
        raise RuntimeError(f"bnb_b200 native error {err}: {lib.cbnb_last_error_string().decode()}")


class CUBLAS_Context:
    """Per-device context cache (reference :137-163); kept for API compatibility."""
    _instance = None

    def __init__(self):
        raise RuntimeError("Call get_instance() instead")

    def initialize(self):
        self.context = {}

    @classmethod
    def get_instance(cls):
        if cls._instance is None:
            cls._instance = cls.__new__(cls)
            cls._instance.initialize()
        return cls._instance

    def get_context(self, device):
        if device.index not in self.context:
            prev = torch.cuda.current_device()
            torch.cuda.set_device(device)
            self.context[device.index] = ct.c_void_p(lib.get_context())
            torch.cuda.set_device(prev)
        return self.context[device.index]


# ------------------------------------------------------------------------------------------------
# QuantState (reference :625-798)
# ------------------------------------------------------------------------------------------------
class QuantState:
    """Container of everything needed to undo a blockwise quantisation: absmax (fp32, or uint8 when
    nested), the code, blocksize, original shape/dtype, and for nested statistics `offset` + `state2`."""

    valid_quant_types = ("fp4", "nf4")
    valid_qs_type_keys = [f"bitsandbytes__{x}" for x in valid_quant_types]
    valid_qs_keys = ["absmax", "quant_map", "nested_absmax", "nested_quant_map", "quant_state", "quant_type",
                     "blocksize", "dtype", "shape", "nested_blocksize", "nested_dtype", "nested_offset"]

    def __init__(self, absmax, shape=None, code=None, blocksize=None, quant_type=None, dtype=None, offset=None,
                 state2=None):
        self.absmax = absmax
        self.shape = shape
        self.code = code
        self.dtype = dtype
        self.blocksize = blocksize
        self.quant_type = quant_type
        self.offset = offset
        self.state2 = state2
        self.nested = state2 is not None

    def __get_item__(self, idx):
        nested = [self.offset, self.state2] if self.nested else None
        return [self.absmax, self.shape, self.dtype, self.blocksize, nested, self.quant_type][idx]

    @classmethod
    def from_dict(cls, qs_dict: Dict[str, Any], device: torch.device) -> "QuantState":
        """Unpack state-dict items (packed or unpacked) into a QuantState (reference :686-735)."""
        qs_key = [k for k, v in qs_dict.items() if "quant_state" in k and isinstance(v, torch.Tensor)]
        if not len(qs_key) and "quant_type" not in qs_dict:
            raise ValueError("Expected packed or unpacked quant_state items, found neither")
        elif len(qs_key) != 1 or qs_key[0].split(".")[-1] not in cls.valid_qs_type_keys:
            raise ValueError(
                f"There should be exactly one `quant_state` item with ending from {cls.valid_qs_type_keys}.\nDetected {qs_key}.")
        if len(qs_key) == 1:
            qs_dict.update(unpack_tensor_to_dict(qs_dict.pop(qs_key[0])))
        qs_dict = {k.split(".")[-1]: v for k, v in qs_dict.items()}
        assert set(qs_dict.keys()).issubset(cls.valid_qs_keys)
        if "nested_absmax" in qs_dict:
            offset = torch.tensor(float(qs_dict["nested_offset"])).to(device)
            state2 = cls(absmax=qs_dict["nested_absmax"].to(device), blocksize=qs_dict["nested_blocksize"],
                         code=qs_dict["nested_quant_map"].to(device), dtype=getattr(torch, qs_dict["nested_dtype"]))
        else:
            offset, state2 = None, None
        return cls(quant_type=qs_dict["quant_type"], absmax=qs_dict["absmax"].to(device),
                   blocksize=qs_dict["blocksize"], code=qs_dict["quant_map"].to(device),
                   dtype=getattr(torch, qs_dict["dtype"]),
                   shape=torch.Size(qs_dict["shape"]) if qs_dict["shape"] is not None else None,
                   offset=offset, state2=state2)

    def as_dict(self, packed=False):
        """Tensors + strings for serialisation (reference :737-767); packed=True folds the non-tensor items
        into one uint8 JSON tensor named quant_state.bitsandbytes__<type>."""
        qs_dict = {"quant_type": self.quant_type, "absmax": self.absmax, "blocksize": self.blocksize,
                   "quant_map": self.code, "dtype": str(self.dtype).strip("torch."), "shape": tuple(self.shape)}
        if self.nested:
            qs_dict.update({"nested_absmax": self.state2.absmax, "nested_blocksize": self.state2.blocksize,
                            "nested_quant_map": self.state2.code.clone(),
                            "nested_dtype": str(self.state2.dtype).strip("torch."),
                            "nested_offset": self.offset.item()})
        if not packed:
            return qs_dict
        packed_dict = {k: v for k, v in qs_dict.items() if isinstance(v, torch.Tensor)}
        non_tensor = {k: v for k, v in qs_dict.items() if not isinstance(v, torch.Tensor)}
        packed_dict["quant_state." + "bitsandbytes__" + self.quant_type] = pack_dict_to_tensor(non_tensor)
        return packed_dict

    def to(self, device):
        self.absmax = self.absmax.to(device)
        if self.code is not None:
            self.code = self.code.to(device)
        if self.nested:
            self.offset = self.offset.to(device)
            self.state2.absmax = self.state2.absmax.to(device)
            self.state2.code = self.state2.code.to(device)

    def __eq__(self, other):
        if not isinstance(other, QuantState):
            return False
        return (torch.allclose(self.absmax, other.absmax, atol=1e-6) and self.shape == other.shape
                and torch.allclose(self.code, other.code, atol=1e-6) and self.dtype == other.dtype
                and self.blocksize == other.blocksize and self.quant_type == other.quant_type
                and (self.offset == other.offset if self.offset is not None and other.offset is not None
                     else self.offset is other.offset)
                and (self.state2 == other.state2 if self.state2 is not None and other.state2 is not None
                     else self.state2 is other.state2))


# ------------------------------------------------------------------------------------------------
# blockwise 8-bit quantize / dequantize (reference :801-1017)
# ------------------------------------------------------------------------------------------------
_BLOCKSIZES = [4096, 2048, 1024, 512, 256, 128, 64]
_SUFFIX = {torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}


def _require_cuda(A: Tensor, what: str):
    if A.device.type != "cuda":
        raise NotImplementedError(f"Device type not supported for {what}: {A.device.type} (bnb_b200 has no CPU path)")


def quantize_blockwise(A: Tensor, code: Optional[Tensor] = None, absmax: Optional[Tensor] = None,
                       out: Optional[Tensor] = None, blocksize=4096, nested=False) -> Tuple[Tensor, QuantState]:
    """8-bit blockwise quantisation with a 256-entry code (reference :801-912) -> cquantize_blockwise_<T>."""
    _require_cuda(A, "blockwise quantization")
    if code is None:
        if "dynamic" not in name2qmap:
            name2qmap["dynamic"] = create_dynamic_map().to(A.device)
        code = name2qmap["dynamic"]
    n = A.numel()
    # the kernel writes every block's absmax and every code byte: no zero fill (the reference allocates with torch.zeros)
    if absmax is None:
        absmax = torch.empty(((n + blocksize - 1) // blocksize,), device=A.device, dtype=torch.float32)
    if out is None:
        out = torch.empty_like(A, dtype=torch.uint8)
    assert blocksize in _BLOCKSIZES
    if A.dtype not in _SUFFIX:
        raise ValueError(f"Blockwise quantization only supports 16/32-bit floats, but got {A.dtype}")
    A = A.contiguous()
    code = code.to(A.device)
    prev = pre_call(A.device)
    is_on_gpu([code, A, out, absmax])
    getattr(lib, f"cquantize_blockwise_{_SUFFIX[A.dtype]}")(
        get_ptr(code), get_ptr(A), get_ptr(absmax), get_ptr(out), ct.c_int32(blocksize), ct.c_int(n))
    post_call(prev)
    if nested:
        offset = absmax.mean()
        absmax -= offset
        qabsmax, state2 = quantize_blockwise(absmax, blocksize=blocksize, nested=False)
        state = QuantState(absmax=qabsmax, code=code, blocksize=blocksize, dtype=A.dtype, offset=offset, state2=state2)
    else:
        state = QuantState(absmax=absmax, code=code, blocksize=blocksize, dtype=A.dtype)
    return out, state


def _denest(quant_state: QuantState) -> Tensor:
    """absmax = dequantize_blockwise(qabsmax, state2); absmax += offset (reference :1346-1350, :1982-1984)."""
    absmax = dequantize_blockwise(quant_state.absmax, quant_state.state2)
    absmax += quant_state.offset
    if absmax.dtype != torch.float32:
        absmax = absmax.float()
    return absmax


def dequantize_blockwise(A: Tensor, quant_state: Optional[QuantState] = None, absmax: Optional[Tensor] = None,
                         code: Optional[Tensor] = None, out: Optional[Tensor] = None, blocksize: int = 4096,
                         nested=False) -> Tensor:
    """Inverse of quantize_blockwise (reference :915-1017) -> cdequantize_blockwise_<T>."""
    assert quant_state is not None or absmax is not None
    _require_cuda(A, "blockwise dequantization")
    if code is None and quant_state is None:
        if "dynamic" not in name2qmap:
            name2qmap["dynamic"] = create_dynamic_map().to(A.device)
        code = name2qmap["dynamic"]
    if quant_state is None:
        quant_state = QuantState(absmax=absmax, code=code, blocksize=blocksize, dtype=torch.float32)
    absmax = quant_state.absmax
    if quant_state.nested:
        absmax = _denest(quant_state)
    if out is None:
        out = torch.empty(A.shape, dtype=quant_state.dtype, device=A.device)
    if quant_state.blocksize not in _BLOCKSIZES:
        raise ValueError(f"The blockwise of {quant_state.blocksize} is not supported. Supported values: {_BLOCKSIZES}")
    if out.dtype not in _SUFFIX:
        raise ValueError(f"Blockwise quantization only supports 16/32-bit floats, but got {out.dtype}")
    qcode = quant_state.code.to(A.device)
    prev = pre_call(A.device)
    is_on_gpu([A, absmax, out, qcode])
    getattr(lib, f"cdequantize_blockwise_{_SUFFIX[out.dtype]}")(
        get_ptr(qcode), get_ptr(A), get_ptr(absmax), get_ptr(out), ct.c_int(quant_state.blocksize), ct.c_int(A.numel()))
    post_call(prev)
    return out


# ------------------------------------------------------------------------------------------------
# 4-bit quantize / dequantize (reference :1102-1424)
# ------------------------------------------------------------------------------------------------
def quantize_fp4(A, absmax=None, out=None, blocksize=64, compress_statistics=False, quant_storage=torch.uint8):
    return quantize_4bit(A, absmax, out, blocksize, compress_statistics, "fp4", quant_storage)


def quantize_nf4(A, absmax=None, out=None, blocksize=64, compress_statistics=False, quant_storage=torch.uint8):
    return quantize_4bit(A, absmax, out, blocksize, compress_statistics, "nf4", quant_storage)


def quantize_4bit(A: Tensor, absmax: Optional[Tensor] = None, out: Optional[Tensor] = None, blocksize=64,
                  compress_statistics=False, quant_type="fp4", quant_storage=torch.uint8) -> Tuple[Tensor, QuantState]:
    """FP4 / NF4 blockwise quantisation, two codes per byte, optional nested (double-quantised) absmax
    (reference :1124-1268) -> cquantize_blockwise_<T>_<fp4|nf4>."""
    _require_cuda(A, "FP4 quantization")
    if quant_type not in ["fp4", "nf4"]:
        raise NotImplementedError(f"4-bit quantization data type {quant_type} is not implemented.")
    n = A.numel()
    input_shape = A.shape
    if absmax is None:
        absmax = torch.empty(((n + blocksize - 1) // blocksize,), device=A.device, dtype=torch.float32)
    if out is None:
        mod = dtype2bytes[quant_storage] * 2
        # every packed byte is written when the storage is bytes and n is even; otherwise the tail must read as zero
        alloc = torch.empty if (quant_storage == torch.uint8 and n % 2 == 0) else torch.zeros
        out = alloc(((n + 1) // mod, 1), dtype=quant_storage, device=A.device)
    assert blocksize in _BLOCKSIZES
    if A.dtype not in _SUFFIX:
        raise ValueError(f"Blockwise quantization only supports 16/32-bit floats, but got {A.dtype}")
    A = A.contiguous()
    prev = pre_call(A.device)
    is_on_gpu([A, out, absmax])
    getattr(lib, f"cquantize_blockwise_{_SUFFIX[A.dtype]}_{quant_type}")(
        get_ptr(None), get_ptr(A), get_ptr(absmax), get_ptr(out), ct.c_int32(blocksize), ct.c_int(n))
    post_call(prev)
    code = get_4bit_type(quant_type, device=A.device)
    if compress_statistics:
        offset = absmax.mean()
        absmax -= offset
        qabsmax, state2 = quantize_blockwise(absmax, blocksize=256)
        del absmax
        state = QuantState(absmax=qabsmax, shape=input_shape, dtype=A.dtype, blocksize=blocksize, code=code,
                           quant_type=quant_type, offset=offset, state2=state2)
    else:
        state = QuantState(absmax=absmax, shape=input_shape, dtype=A.dtype, blocksize=blocksize, code=code,
                           quant_type=quant_type)
    return out, state


def dequantize_fp4(A, quant_state=None, absmax=None, out=None, blocksize=64):
    return dequantize_4bit(A, quant_state, absmax, out, blocksize, "fp4")


def dequantize_nf4(A, quant_state=None, absmax=None, out=None, blocksize=64):
    return dequantize_4bit(A, quant_state, absmax, out, blocksize, "nf4")


def dequantize_4bit(A: Tensor, quant_state: Optional[QuantState] = None, absmax: Optional[Tensor] = None,
                    out: Optional[Tensor] = None, blocksize: int = 64, quant_type="fp4") -> Tensor:
    """Inverse of quantize_4bit (reference :1291-1424) -> cdequantize_blockwise_<T>_<fp4|nf4>."""
    if blocksize not in _BLOCKSIZES:
        raise ValueError(f"The blockwise of {blocksize} is not supported. Supported values: {_BLOCKSIZES}")
    if quant_type not in ["fp4", "nf4"]:
        raise NotImplementedError(f"4-bit quantization data type {quant_type} is not implemented.")
    _require_cuda(A, "FP4 dequantization")
    if quant_state is None:
        assert absmax is not None and out is not None
        quant_state = QuantState(absmax=absmax, shape=out.shape, dtype=out.dtype, blocksize=blocksize,
                                 quant_type=quant_type)
    else:
        absmax = quant_state.absmax
    if quant_state.nested:
        absmax = _denest(quant_state)
    if out is None:
        out = torch.empty(quant_state.shape, dtype=quant_state.dtype, device=A.device)
    if out.dtype not in _SUFFIX:
        raise ValueError(f"Blockwise quantization only supports 16/32-bit floats, but got {out.dtype}")
    n = out.numel()
    prev = pre_call(A.device)
    is_on_gpu([A, absmax, out])
    getattr(lib, f"cdequantize_blockwise_{_SUFFIX[out.dtype]}_{quant_state.quant_type}")(
        get_ptr(None), get_ptr(A), get_ptr(absmax), get_ptr(out), ct.c_int(quant_state.blocksize), ct.c_int(n))
    post_call(prev)
    if A.shape[0] == 1:  # is_transposed (reference :1420-1422)
        return out.t()
    return out


# ------------------------------------------------------------------------------------------------
# 4-bit GEMV / GEMM (reference :1961-2060; batch > 1: autograd/_functions.py:490-518)
# ------------------------------------------------------------------------------------------------
class GemvSync(ct.Structure):
    """bnb_gemv_sync_t of include/bnb_b200.h: cross-GPU signal / wait folded into the GEMV kernels."""
    _fields_ = [("sig_local", ct.c_void_p), ("sig_peer", ct.c_void_p * 7), ("epoch", ct.c_void_p),
                ("cta_counter", ct.c_void_p), ("gidx", ct.c_int), ("ngroups", ct.c_int), ("do_signal", ct.c_int),
                ("do_wait", ct.c_int)]


def gemv_4bit(A: Tensor, B: Tensor, out: Optional[Tensor] = None, transposed_A=False, transposed_B=False, state=None,
              peer_outs=None, sync: Optional[GemvSync] = None):
    """out[.., N] = A[.., K] @ dequant(B)[N, K]^T for a single activation row.
    Reference: de-nest absmax (2 launches) then cgemm_4bit_inference_naive_<T>.  Here, when the state is
    nested and the fused path is on, ONE launch (cgemm_4bit_inference_nested_<T>) reads the uint8 absmax
    directly; the de-nest arithmetic inside the kernel is the same fl(fl(code2[q]*absmax2)+offset).
    ADDITIVE `peer_outs`: list of peer-mapped device addresses (ints) -- the kernel stores the result there too
    (N-sharded linear on one NVLink box: the output all-gather fused into the epilogue, parallel.py)."""
    if state is None:
        raise ValueError("state cannot None. gem_4bit( ) requires the state from quantize_4bit( )")
    if A.numel() != A.shape[-1]:
        raise ValueError('Dimensions of A are invalid. Must be a vector with the leading dimensions of "1", e.g. [1, 1, 2048]')
    _require_cuda(A, "gemv_4bit")
    Bshape = state.shape
    bout = Bshape[0]
    if out is None:
        if len(A.shape) == 3:
            out = torch.empty(size=(A.shape[0], A.shape[1], bout), dtype=A.dtype, device=A.device)
        else:
            out = torch.empty(size=(A.shape[0], bout), dtype=A.dtype, device=A.device)
    m, n, k = Bshape[0], 1, Bshape[1]
    lda = ldc = Bshape[0]
    ldb = (A.shape[-1] + 1) // 2
    if A.dtype not in _SUFFIX:
        raise NotImplementedError(f"Matmul not implemented for data type {A.dtype}")
    if B.dtype not in [torch.uint8, torch.bfloat16, torch.float16, torch.float32]:
        raise NotImplementedError(f"Matmul not implemented for data type {B.dtype}")
    A = A.contiguous()
    code = state.code.to(A.device)
    fused = (FUSED_NESTED_GEMV and state.nested and A.dtype in (torch.float16, torch.bfloat16)
             and k % 64 == 0 and state.blocksize % 64 == 0 and A.shape[-1] == k
             and A.data_ptr() % 16 == 0 and B.data_ptr() % 8 == 0)     # what the native fast path needs; else the reference-shaped route
    if fused:
        s2 = state.state2
        offset = getattr(state, "_offset_host", None)
        if offset is None:  # one host read per weight, cached: the scalar torch computed at quantize time
            offset = state._offset_host = float(state.offset)
        tabs = getattr(state, "_tables_host", None)
        if tabs is None:    # host copies of the two code tables, once per weight (like the offset scalar above)
            if code.numel() == 16 and s2.code.numel() == 256:
                tabs = ((ct.c_float * 16)(*code.float().cpu().tolist()), (ct.c_float * 256)(*s2.code.float().cpu().tolist()))
            else:
                tabs = (None, None)
            state._tables_host = tabs
        prev = pre_call(A.device)
        is_on_gpu([B, A, out, state.absmax, s2.absmax, s2.code, code])
        lib.cbnb_set_gemv_host_tables(tabs[0], tabs[1])
        if peer_outs:
            arr = (ct.c_void_p * len(peer_outs))(*[ct.c_void_p(int(p)) for p in peer_outs])
            getattr(lib, f"cgemm_4bit_inference_nested_push_{_SUFFIX[A.dtype]}")(
                ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), get_ptr(A), get_ptr(B), get_ptr(state.absmax),
                get_ptr(s2.absmax), get_ptr(s2.code), ct.c_float(offset), get_ptr(code), get_ptr(out),
                ct.c_int32(lda), ct.c_int32(ldb), ct.c_int32(ldc), ct.c_int32(state.blocksize), ct.c_int32(s2.blocksize),
                arr, ct.c_int32(len(peer_outs)), ct.byref(sync) if sync is not None else None)
            post_call(prev)
            return out
        getattr(lib, f"cgemm_4bit_inference_nested_{_SUFFIX[A.dtype]}")(
            ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), get_ptr(A), get_ptr(B), get_ptr(state.absmax),
            get_ptr(s2.absmax), get_ptr(s2.code), ct.c_float(offset), get_ptr(code), get_ptr(out),
            ct.c_int32(lda), ct.c_int32(ldb), ct.c_int32(ldc), ct.c_int32(state.blocksize), ct.c_int32(s2.blocksize))
        post_call(prev)
        return out
    if peer_outs:
        raise NotImplementedError("peer_outs needs the fused nested GEMV (nested absmax, fp16/bf16, blocksize 64)")
    absmax = state.absmax
    if state.nested:
        absmax = dequantize_blockwise(state.absmax, state.state2)
        absmax += state.offset
    prev = pre_call(A.device)
    is_on_gpu([B, A, out, absmax, code])
    getattr(lib, f"cgemm_4bit_inference_naive_{_SUFFIX[A.dtype]}")(
        ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), get_ptr(A), get_ptr(B), get_ptr(absmax), get_ptr(code),
        get_ptr(out), ct.c_int32(lda), ct.c_int32(ldb), ct.c_int32(ldc), ct.c_int32(state.blocksize))
    post_call(prev)
    return out


def gemv_4bit_multi(A: Tensor, Bs, states, outs=None, peer_outs=None):
    """ADDITIVE: out_i = A @ dequant(B_i)^T for up to four nested-absmax 4-bit weights that share the activation row A
    (q/k/v, gate/up of a decoder layer) in ONE launch.  Bit-identical to gemv_4bit per matrix; falls back to the single
    calls for shapes the native path does not take.  Bs: packed weights as passed to gemv_4bit (already .t()-viewed)."""
    n = len(Bs)
    if outs is None:
        outs = [None] * n
    s0 = states[0]
    ok = (1 <= n <= 4 and A.numel() == A.shape[-1] and A.dtype in (torch.float16, torch.bfloat16)
          and all(st.nested and st.blocksize == 64 and st.shape[1] == s0.shape[1] and st.quant_type == s0.quant_type
                  and st.state2.blocksize == s0.state2.blocksize for st in states)
          and s0.shape[1] % 256 == 0 and A.shape[-1] == s0.shape[1])
    if ok:
        key = tuple(id(st) for st in states)
        cache = getattr(s0, "_multi_code2_checked", None)
        if cache is None or cache[0] != key:   # the shared dynamic map: checked once per group of states
            cache = s0._multi_code2_checked = (key, all(torch.equal(st.state2.code, s0.state2.code) for st in states[1:]))
        ok = cache[1]
    if not ok:
        return [gemv_4bit(A, B, out=o, state=st, peer_outs=(peer_outs[i] if peer_outs else None))
                for i, (B, st, o) in enumerate(zip(Bs, states, outs))]
    _require_cuda(A, "gemv_4bit_multi")
    A = A.contiguous()
    k = s0.shape[1]
    outs = [o if o is not None else torch.empty((*A.shape[:-1], st.shape[0]), dtype=A.dtype, device=A.device)
            for o, st in zip(outs, states)]
    code = s0.code.to(A.device)
    offs = []
    for st in states:
        off = getattr(st, "_offset_host", None)
        if off is None:
            off = st._offset_host = float(st.offset)
        offs.append(off)
    tabs = getattr(s0, "_tables_host", None)
    if tabs is None:
        if code.numel() == 16 and s0.state2.code.numel() == 256:
            tabs = ((ct.c_float * 16)(*code.float().cpu().tolist()), (ct.c_float * 256)(*s0.state2.code.float().cpu().tolist()))
        else:
            tabs = (None, None)
        s0._tables_host = tabs
    vp = ct.c_void_p
    prev = pre_call(A.device)
    is_on_gpu([A, code, s0.state2.code] + list(Bs) + list(outs) + [st.absmax for st in states] + [st.state2.absmax for st in states])
    lib.cbnb_set_gemv_host_tables(tabs[0], tabs[1])
    common = (ct.c_int32(n), (ct.c_int32 * n)(*[st.shape[0] for st in states]), ct.c_int32(k), get_ptr(A),
              (vp * n)(*[B.data_ptr() for B in Bs]), (vp * n)(*[st.absmax.data_ptr() for st in states]),
              (vp * n)(*[st.state2.absmax.data_ptr() for st in states]), get_ptr(s0.state2.code), (ct.c_float * n)(*offs),
              get_ptr(code), (vp * n)(*[o.data_ptr() for o in outs]), ct.c_int32(64), ct.c_int32(s0.state2.blocksize))
    if peer_outs:   # N-sharded: [matrix][peer] peer-mapped addresses of the output slices
        npeers = len(peer_outs[0])
        flat = [int(p) for po in peer_outs for p in po]
        rc = getattr(lib, f"cgemm_4bit_inference_nested_multi_push_{_SUFFIX[A.dtype]}")(
            *common, (vp * len(flat))(*flat), ct.c_int32(npeers))
    else:
        rc = getattr(lib, f"cgemm_4bit_inference_nested_multi_{_SUFFIX[A.dtype]}")(*common)
    post_call(prev)
    if rc != 0:
        return [gemv_4bit(A, B, out=o, state=st, peer_outs=(peer_outs[i] if peer_outs else None))
                for i, (B, st, o) in enumerate(zip(Bs, states, outs))]
    return outs


def gemm_4bit(A: Tensor, B: Tensor, state: QuantState, bias: Optional[Tensor] = None,
              out: Optional[Tensor] = None, peer_outs: Optional[List[int]] = None) -> Optional[Tensor]:
    """ADDITIVE: fused batch>1 4-bit GEMM, out[b, N] = A[b, K] @ T(dequant(B))^T (+bias) without ever
    materialising the dequantised weight (tcgen05 kind::f16, TMEM accumulators).  Returns None when the
    native kernel does not take the shape -- the caller then runs the reference route
    (dequantize_4bit + F.linear, autograd/_functions.py:507).

    N-sharded stacks (bnb_b200/parallel.py): `out` may be a [batch, N] column slice of a wider row-major buffer (row stride
    out.stride(0)), and `peer_outs` the device addresses of the same slice in the peers' copies of that buffer (NVLink peer
    mappings): the epilogue stores into all of them (cgemm_4bit_push_*)."""
    if not FUSED_GEMM_4BIT or A.dtype not in (torch.float16, torch.bfloat16):
        return None
    if state.dtype not in (A.dtype, torch.float32):
        return None     # reference: dequantize to state.dtype, then .to(A.dtype) -- two roundings the fused kernel does not do
    if len(state.shape) != 2 or B.dim() != 2 or B.shape[0] != 1:
        return None     # un-transposed packed weight: the reference computes A @ W there, not A @ W^T
    N, K = state.shape
    if A.shape[-1] != K:
        return None     # wrong activation width: the reference route raises the shape error of F.linear
    A2 = A.reshape(-1, A.shape[-1]).contiguous()
    batch = A2.shape[0]
    strided = out is not None and (out.dim() != 2 or tuple(out.shape) != (batch, N) or out.stride(1) != 1 or out.stride(0) != N)
    if strided and (out.dim() != 2 or tuple(out.shape) != (batch, N) or out.stride(1) != 1 or out.dtype != A.dtype):
        raise ValueError("gemm_4bit: out must be a [batch, N] view with unit column stride and the dtype of A")
    push = strided or bool(peer_outs)
    if (not push and 2 <= batch <= 8 and bias is None and state.nested and FUSED_NESTED_GEMV and K % 64 == 0 and N % 16 == 0
            and state.blocksize % 64 == 0 and A2.shape[1] == K and B.dtype == torch.uint8):
        # batch 2..8: the batch rides in the MMA's n dimension of the LUT + mma.sync GEMV kernels (nested absmax read
        # directly, fp32 absmax applied to fp32 partial sums): ~1.8x faster than the tcgen05 kernel at these sizes
        # (18.5 vs 34 us on 14336x4096), and no de-nested absmax copy.
        if out is None:
            out = torch.empty((batch, N), dtype=A.dtype, device=A.device)
        s2 = state.state2
        offset = getattr(state, "_offset_host", None)
        if offset is None:
            offset = state._offset_host = float(state.offset)
        code = state.code.to(A.device)
        prev = pre_call(A.device)
        is_on_gpu([A2, B, state.absmax, s2.absmax, s2.code, code, out])
        getattr(lib, f"cgemm_4bit_inference_nested_{_SUFFIX[A.dtype]}")(
            ct.c_int32(N), ct.c_int32(batch), ct.c_int32(K), get_ptr(A2), get_ptr(B), get_ptr(state.absmax),
            get_ptr(s2.absmax), get_ptr(s2.code), ct.c_float(offset), get_ptr(code), get_ptr(out),
            ct.c_int32(N), ct.c_int32((K + 1) // 2), ct.c_int32(N), ct.c_int32(state.blocksize), ct.c_int32(s2.blocksize))
        post_call(prev)
        return out.reshape(*A.shape[:-1], N)
    if state.nested:
        # the de-nested fp32 absmax (two launches, reference :1346-1350) is a constant of the frozen weight: computed
        # once and kept with the state (+ 1/16 byte per weight) instead of once per forward
        absmax = getattr(state, "_absmax_f32", None)
        if absmax is None or absmax.device != A.device:
            absmax = state._absmax_f32 = _denest(state)
    else:
        absmax = state.absmax
    if out is None:
        out = torch.empty((batch, N), dtype=A.dtype, device=A.device)
    code = state.code.to(A.device)
    b = None if bias is None else bias.to(A.dtype).contiguous()
    prev = pre_call(A.device)
    is_on_gpu([A2, B, absmax, code, out, b])
    if push:
        peer_outs = list(peer_outs or [])
        if len(peer_outs) > 7:
            raise ValueError("gemm_4bit: at most 7 peers")
        parr = (ct.c_void_p * max(len(peer_outs), 1))(*[ct.c_void_p(int(p)) for p in peer_outs])
        rc = getattr(lib, f"cgemm_4bit_push_{_SUFFIX[A.dtype]}")(
            ct.c_int32(batch), ct.c_int32(N), ct.c_int32(K), get_ptr(A2), get_ptr(B), get_ptr(absmax), get_ptr(code),
            get_ptr(b), get_ptr(out), ct.c_int32(state.blocksize), ct.c_long(out.stride(0)), parr, ct.c_int32(len(peer_outs)))
    else:
        rc = getattr(lib, f"cgemm_4bit_{_SUFFIX[A.dtype]}")(
            ct.c_int32(batch), ct.c_int32(N), ct.c_int32(K), get_ptr(A2), get_ptr(B), get_ptr(absmax), get_ptr(code),
            get_ptr(b), get_ptr(out), ct.c_int32(state.blocksize))
    post_call(prev)
    if rc != 0:
        return None
    return out if push else out.reshape(*A.shape[:-1], N)


# ------------------------------------------------------------------------------------------------
# LLM.int8: statistics, double quant, layouts, igemmlt, mm_dequant, outliers (reference :2260-2653, :2914)
# ------------------------------------------------------------------------------------------------
def get_transform_buffer(shape, dtype, device, to_order, from_order="row", transpose=False):
    """Zeroed buffer + (shape, order) state for a layout (reference :482-518)."""
    dims = len(shape)
    if dims == 2:
        rows = shape[0]
    elif dims == 3:
        rows = shape[0] * shape[1]
    cols = shape[-1]
    state = (shape, to_order)
    if transpose:
        rows, cols = cols, rows
        state = (shape[::-1], to_order)
    rows_in, cols_in = rows, cols
    if to_order in ("row", "col"):
        return torch.zeros(shape, dtype=dtype, device=device), state
    elif to_order == "col32":
        cols = 32 * ((cols + 31) // 32)
    elif to_order == "col_turing":
        cols = 32 * ((cols + 31) // 32)
        rows = 8 * ((rows + 7) // 8)
    elif to_order == "col_ampere":
        cols = 32 * ((cols + 31) // 32)
        rows = 32 * ((rows + 31) // 32)
    else:
        raise NotImplementedError(f"To_order not supported: {to_order}")
    # without padding every byte is written by the kernel that fills the buffer: skip the zero fill (a 16 MB memset costs
    # as much as a third of the 4096 x 4096 transform itself)
    alloc = torch.empty if (rows, cols) == (rows_in, cols_in) else torch.zeros
    return alloc((rows, cols), dtype=dtype, device=device), state


def get_colrow_absmax(A, row_stats=None, col_stats=None, nnz_block_ptr=None, threshold=0.0):
    """Row / column absmax of an fp16 matrix, outliers (|x| >= threshold) excluded and counted per
    16x256 tile row (reference :2400-2435) -> cget_col_row_stats, then cumsum on the host side."""
    assert A.dtype == torch.float16
    _require_cuda(A, "get_colrow_absmax")
    device = A.device
    cols = A.shape[-1]
    rows = A.shape[0] * A.shape[1] if len(A.shape) == 3 else A.shape[0]
    col_tiles = (cols + 255) // 256
    tiled_rows = ((rows + 15) // 16) * 16
    if row_stats is None:
        row_stats = torch.empty((rows,), dtype=torch.float32, device=device).fill_(-50000.0)
    if col_stats is None:
        col_stats = torch.empty((cols,), dtype=torch.float32, device=device).fill_(-50000.0)
    if nnz_block_ptr is None and threshold > 0.0:
        nnz_block_ptr = torch.zeros(((tiled_rows * col_tiles) + 1,), dtype=torch.int32, device=device)
    A = A.contiguous()
    prev = pre_call(device)
    is_on_gpu([A, row_stats, col_stats, nnz_block_ptr])
    lib.cget_col_row_stats(get_ptr(A), get_ptr(row_stats), get_ptr(col_stats), get_ptr(nnz_block_ptr),
                           ct.c_float(threshold), ct.c_int32(rows), ct.c_int32(cols))
    post_call(prev)
    if threshold > 0.0:
        nnz_block_ptr.cumsum_(0)
    return row_stats, col_stats, nnz_block_ptr


class COOSparseTensor:
    """reference :2438-2454."""

    def __init__(self, rows, cols, nnz, rowidx, colidx, values):
        assert rowidx.dtype == torch.int32 and colidx.dtype == torch.int32 and values.dtype == torch.float16
        assert values.numel() == nnz and rowidx.numel() == nnz and colidx.numel() == nnz
        self.rows, self.cols, self.nnz = rows, cols, nnz
        self.rowidx, self.colidx, self.values = rowidx, colidx, values


def coo_zeros(rows, cols, nnz, device, dtype=torch.half):
    """reference :2510-2514."""
    return COOSparseTensor(rows, cols, nnz, torch.zeros((nnz,), dtype=torch.int32, device=device),
                           torch.zeros((nnz,), dtype=torch.int32, device=device),
                           torch.zeros((nnz,), dtype=dtype, device=device))


def double_quant(A, col_stats=None, row_stats=None, out_col=None, out_row=None, threshold=0.0):
    """Row-wise and column-wise int8 quantisation of an fp16 matrix in one pass, outliers split off
    into a COO tensor (reference :2517-2604) -> cdouble_rowcol_quant.
    Returns (out_row, out_col, row_stats, col_stats, coo_tensor | None)."""
    device = A.device
    assert A.dtype == torch.half
    _require_cuda(A, "double_quant")
    cols = A.shape[-1]
    rows = A.shape[0] * A.shape[1] if len(A.shape) == 3 else A.shape[0]
    nnz_row_ptr = None
    if row_stats is None or col_stats is None:
        row_stats, col_stats, nnz_row_ptr = get_colrow_absmax(A, threshold=threshold)
    if out_col is None:      # every element of both outputs is written by the kernel
        out_col = torch.empty(A.shape, device=device, dtype=torch.int8)
    if out_row is None:
        out_row = torch.empty(A.shape, device=device, dtype=torch.int8)
    A = A.contiguous()
    coo_tensor = None
    args = [get_ptr(A), get_ptr(row_stats), get_ptr(col_stats), get_ptr(out_col), get_ptr(out_row)]
    nnz = 0
    if threshold > 0.0 and nnz_row_ptr is not None:
        nnz = int(nnz_row_ptr[-1].item())  # host sync, as in the reference (:2546)
    prev = pre_call(device)
    is_on_gpu([A, col_stats, row_stats, out_col, out_row])
    if nnz > 0:
        coo_tensor = coo_zeros(A.shape[0], A.shape[1], nnz, device)
        lib.cdouble_rowcol_quant(*args, get_ptr(coo_tensor.rowidx), get_ptr(coo_tensor.colidx),
                                 get_ptr(coo_tensor.values), get_ptr(nnz_row_ptr), ct.c_float(threshold),
                                 ct.c_int32(rows), ct.c_int32(cols))
        post_call(prev)
        val, idx = torch.sort(coo_tensor.rowidx, stable=True)
        coo_tensor.rowidx = val
        coo_tensor.colidx = coo_tensor.colidx[idx]
        coo_tensor.values = coo_tensor.values[idx]
    else:
        lib.cdouble_rowcol_quant(*args, None, None, None, None, ct.c_float(0.0 if threshold > 0.0 else threshold),
                                 ct.c_int32(rows), ct.c_int32(cols))
        post_call(prev)
    return out_row, out_col, row_stats, col_stats, coo_tensor


def transform(A, to_order, from_order="row", out=None, transpose=False, state=None, ld=None):
    """int8 layout transform row -> col32 / col_turing / col_ampere (reference :2607-2653).
    to_order == "row" from "row" is the identity (B200-native weight layout)."""
    if state is None:
        state = (A.shape, from_order)
    else:
        from_order = state[1]
    if to_order == "row" and from_order == "row" and not transpose:
        return A, (state[0], "row")
    _require_cuda(A, "transform")
    if out is None:
        out, new_state = get_transform_buffer(state[0], A.dtype, A.device, to_order, state[1], transpose)
    else:
        new_state = (state[0], to_order)
    shape = state[0]
    if len(shape) == 2:
        dim1, dim2 = ct.c_int32(shape[0]), ct.c_int32(shape[1])
    else:
        dim1, dim2 = ct.c_int32(shape[0] * shape[1]), ct.c_int32(shape[2])
    names = {"col32": "col32", "col_turing": "turing", "col_ampere": "ampere"}
    if to_order == "row" and from_order in names and not transpose:
        # inverse layouts (reference :2645-2647 calls ctransform_turing2row / ctransform_ampere2row)
        sym = f"ctransform_{names[from_order]}2row"
    elif to_order in names and from_order == "row":
        sym = f"ctransform_row2{names[to_order]}{'T' if transpose else ''}"
    else:
        raise NotImplementedError(f"Transform function not implemented: From {from_order} to {to_order}")
    A = A.contiguous()
    prev = pre_call(A.device)
    is_on_gpu([A, out])
    getattr(lib, sym)(get_ptr(A), get_ptr(out), dim1, dim2)
    post_call(prev)
    return out, new_state


def undo_layout_to_row(weight: Tensor, weight_format: str, rows: Optional[int] = None, cols: Optional[int] = None) -> Tensor:
    """An int8 weight stored in col32 / col_turing / col_ampere (a checkpoint written from `state.CxB`, reference
    nn/modules.py:725-796) back to row-major [rows, cols]: the inverse of kTransformRowToFormat's index maps
    (kernel_quant.cpp:3673-3675, 3740-3755, 3822-3832 == blas_utils.h:244-346), as views + one in-tile gather, on
    whatever device the tensor lives (state dicts are usually loaded on the CPU).  Plays the role of the reference's
    undo_layout(weight, get_tile_inds(...)) (autograd/_functions.py:89-104)."""
    R, C = weight.shape[-2], weight.shape[-1]
    flat = weight.reshape(-1)
    assert C % 32 == 0, "formatted weights are padded to 32 columns"
    if weight_format == "col32":
        out = flat.view(C // 32, R, 32).permute(1, 0, 2).reshape(R, C)
    elif weight_format == "col_turing":
        assert R % 8 == 0, "col_turing pads rows to 8"
        r8 = torch.arange(8, device=weight.device).view(8, 1)
        c32 = torch.arange(32, device=weight.device).view(1, 32)
        w = torch.where(r8 % 2 == 1, 128 + (r8 - 1) * 2, r8 * 2) + (c32 // 4) * 16 + (c32 % 4)     # offset inside the 8x32 tile
        out = flat.view(C // 32, R // 8, 256)[:, :, w.reshape(-1)].view(C // 32, R // 8, 8, 32).permute(1, 2, 0, 3).reshape(R, C)
    elif weight_format == "col_ampere":
        assert R % 32 == 0, "col_ampere pads rows to 32"
        lr = torch.arange(32, device=weight.device)
        ar = ((lr % 8) // 2) * 8 + (lr // 8) * 2 + (lr % 2)                                        # tile row of logical row lr
        out = flat.view(C // 32, R // 32, 32, 32)[:, :, ar, :].permute(1, 2, 0, 3).reshape(R, C)
    else:
        raise ValueError(f"Unrecognized weights format {weight_format}")
    rows = R if rows is None else rows
    cols = C if cols is None else cols
    return out[:rows, :cols].contiguous()


def igemmlt(A, B, SA, SB, out=None, Sout=None, dtype=torch.int32):
    """int8 GEMM out = A @ B^T with int32 (or int8) output (reference :2260-2352).
    Reference layouts: A col32, B col_turing|col_ampere, out col32 -> cigemmlt_<fmt>_<32|8>.
    B200-native (additive): SA[1] == SB[1] == "row" -> cigemm_rowmajor_32, out row-major."""
    shapeA, shapeB = SA[0], SB[0]
    dimsA, dimsB = len(shapeA), len(shapeB)
    assert dimsB == 2, "Only two dimensional matrices are supported for argument B"
    if dimsA == 2:
        m = shapeA[0]
    elif dimsA == 3:
        m = shapeA[0] * shapeA[1]
    rows = n = shapeB[0]
    assert prod(list(shapeA)) > 0, f"Input tensor dimensions need to be > 0: {shapeA}"
    if shapeA[0] == 0 and dimsA == 2:
        return torch.empty((0, shapeB[0]), device=A.device, dtype=torch.float16)
    elif shapeA[1] == 0 and dimsA == 3:
        return torch.empty(tuple(shapeA[:2] + [shapeB[0]]), device=A.device, dtype=torch.float16)
    native = SA[1] == "row" and SB[1] == "row"
    out_order = "row" if native else "col32"
    if dimsA == 2 and out is None:
        out, Sout = get_transform_buffer((shapeA[0], shapeB[0]), dtype, A.device, out_order, "row")
    elif dimsA == 3 and out is None:
        out, Sout = get_transform_buffer((shapeA[0], shapeA[1], shapeB[0]), dtype, A.device, out_order, "row")
    assert A.device.type == "cuda" and B.device.type == "cuda"
    assert A.dtype == torch.int8 and B.dtype == torch.int8
    assert out.dtype == dtype
    assert shapeA[-1] == shapeB[-1], (
        f"Matmullt only supports A @ B^T. Inner matrix dimensions do not match: A @ B = {shapeA} @ {shapeB}")
    k = shapeA[-1]
    prev = pre_call(A.device)
    is_on_gpu([A, B, out])
    if native:
        assert dtype == torch.int32, "row-major igemm produces int32"
        has_error = lib.cigemm_rowmajor_32(ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), get_ptr(A), get_ptr(B), get_ptr(out))
    else:
        assert SA[1] == "col32"
        assert SB[1] in ["col_turing", "col_ampere"]
        assert Sout[1] == "col32"
        formatB = SB[1]
        lda = ct.c_int32(m * 32)
        ldb = ct.c_int32(((rows + 7) // 8) * 8 * 32 if formatB == "col_turing" else ((rows + 31) // 32) * 32 * 32)
        ldc = ct.c_int32(m * 32)
        fmt = "turing" if formatB == "col_turing" else "ampere"
        fn = getattr(lib, f"cigemmlt_{fmt}_{'32' if dtype == torch.int32 else '8'}")
        has_error = fn(ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), get_ptr(A), get_ptr(B), get_ptr(out), get_ptr(None),
                       lda, ldb, ldc)
    post_call(prev)
    if has_error == 1:
        raise NotImplementedError("igemmlt not available (probably built with NO_CUBLASLT)")
    if has_error:
        raise Exception("cublasLt ran into an error!")
    return out, Sout


def mm_dequant(A, quant_state, row_stats, col_stats, out=None, new_row_stats=None, new_col_stats=None, bias=None):
    """int32 (col32) -> fp16 row-major: out = half(((c * 6.200012e-05) * row) * col + bias)
    (reference :2355-2397) -> cdequant_mm_int32_fp16."""
    assert A.dtype == torch.int32
    if bias is not None:
        assert bias.dtype == torch.float16
    out_shape = quant_state[0]
    if len(out_shape) == 3:
        out_shape = (out_shape[0] * out_shape[1], out_shape[2])
    if out is None:
        out = torch.empty(out_shape, dtype=torch.float16, device=A.device)
    if new_row_stats is None:
        new_row_stats = torch.empty(out_shape[0], dtype=torch.float32, device=A.device)
    if new_col_stats is None:
        new_col_stats = torch.empty(out_shape[1], dtype=torch.float32, device=A.device)
    assert new_row_stats.shape[0] == row_stats.shape[0], f"{new_row_stats.shape} vs {row_stats.shape}"
    assert new_col_stats.shape[0] == col_stats.shape[0], f"{new_col_stats.shape} vs {col_stats.shape}"
    if quant_state[1] == "row":  # B200-native row-major accumulators: same arithmetic, torch-free kernel path
        A = transform_int32_row_to_col32(A, out_shape)
    prev = pre_call(A.device)
    is_on_gpu([A, row_stats, col_stats, out, new_row_stats, new_col_stats, bias])
    lib.cdequant_mm_int32_fp16(get_ptr(A), get_ptr(row_stats), get_ptr(col_stats), get_ptr(out), get_ptr(new_row_stats),
                               get_ptr(new_col_stats), get_ptr(bias), ct.c_int32(out_shape[0]), ct.c_int32(out_shape[1]))
    post_call(prev)
    return out


def transform_int32_row_to_col32(A: Tensor, shape) -> Tensor:
    """row-major int32 [rows, cols] -> col32 (pure index permutation done with torch views; only used when
    a caller feeds row-major accumulators to the reference-shaped mm_dequant)."""
    rows, cols = shape
    padded = 32 * ((cols + 31) // 32)
    buf = torch.zeros((rows, padded), dtype=A.dtype, device=A.device)
    buf[:, :cols] = A.reshape(rows, cols)
    return buf.reshape(rows, padded // 32, 32).permute(1, 0, 2).contiguous()


def int8_linear_dequant(CA: Tensor, CB: Tensor, SCA: Tensor, SCB: Tensor, bias: Optional[Tensor] = None,
                        out: Optional[Tensor] = None) -> Tensor:
    """ADDITIVE: igemmlt + mm_dequant in ONE kernel (tcgen05 kind::i8, dequant in the TMEM epilogue):
    out[m, n] = half(((sum_k CA[m,k]*CB[n,k]) * 6.200012e-05 * SCA[m]) * SCB[n] + bias[n]); CA, CB row-major
    int8.  Bit-identical to igemmlt -> mm_dequant, minus the 400 MB int32 round trip."""
    assert CA.dtype == torch.int8 and CB.dtype == torch.int8
    m, k = CA.shape
    n = CB.shape[0]
    assert CB.shape[1] == k
    if bias is not None:
        assert bias.dtype == torch.float16
    if out is None:
        out = torch.empty((m, n), dtype=torch.float16, device=CA.device)
    prev = pre_call(CA.device)
    is_on_gpu([CA, CB, SCA, SCB, bias, out])
    rc = lib.cigemm_rowmajor_dequant_fp16(ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), get_ptr(CA), get_ptr(CB),
                                          get_ptr(SCA), get_ptr(SCB), get_ptr(bias), get_ptr(out))
    post_call(prev)
    if rc != 0:
        raise Exception("cublasLt ran into an error!")
    return out


def int8_linear_fused(A: Tensor, CB: Tensor, SCB: Tensor, bias: Optional[Tensor] = None, threshold: float = 6.0,
                      out: Optional[Tensor] = None, return_quantized: bool = False):
    """ADDITIVE: the whole LLM.int8 inference forward (MatMul8bitLt.forward, reference _functions.py:292-434, with
    has_fp16_weights=False and threshold > 0) in one native call: no host synchronisation, no torch.unique / sort /
    index kernels.  A fp16 [m, k]; CB int8 [n, k] row-major; SCB fp32 [n].  Returns out fp16 [m, n], or None when
    the native path does not take the shape (the caller then runs the step-by-step route).  With
    return_quantized=True also returns (CA, SCA, idx, count): quantised activations (outlier columns zeroed), row
    statistics, ascending outlier column list and its device-side length."""
    if A.dtype != torch.float16 or CB.dtype != torch.int8 or threshold <= 0.0:
        return None
    A2 = A.reshape(-1, A.shape[-1]).contiguous()
    m, k = A2.shape
    n = CB.shape[0]
    if k % 16 != 0 or CB.shape[1] != k or not CB.is_contiguous():
        return None
    # workspace: ONE allocation per call from torch's caching allocator (stream-ordered reuse, safe across streams,
    # threads and CUDA-graph capture -- a module-global scratch shared by every layer was a race); the flags, the
    # counter and the index list must start at zero and sit together so that one small fill clears them
    dev = A.device
    idx_cap = max(k, 16)
    sizes = [("colflag", k, torch.uint8), ("count", 1, torch.int32), ("idx", idx_cap, torch.int32), ("pos", k, torch.int16),
             ("SCA", m, torch.float32), ("subA", m * 16, torch.float16), ("subB", n * 16, torch.float16), ("CA", m * k, torch.int8)]
    offs, off = {}, 0
    for name, cnt, dt in sizes:
        offs[name] = off
        off += (cnt * torch.empty((), dtype=dt).element_size() + 255) // 256 * 256
    raw = torch.empty(off, dtype=torch.uint8, device=dev)
    raw[:offs["pos"]].zero_()
    ws = {name: raw[offs[name]:offs[name] + cnt * torch.empty((), dtype=dt).element_size()].view(dt) for name, cnt, dt in sizes}
    ws["CA"] = ws["CA"].view(m, k)
    ws["subA"] = ws["subA"].view(m, 16)
    ws["subB"] = ws["subB"].view(n, 16)
    if bias is not None and bias.dtype != torch.float16:
        return None
    if out is None:
        out = torch.empty((m, n), dtype=torch.float16, device=A.device)
    SCB32 = SCB if SCB.dtype == torch.float32 else SCB.float()
    prev = pre_call(A.device)
    is_on_gpu([A2, CB, SCB32, bias, out])
    rc = lib.cint8_linear_fp16(get_ptr(A2), get_ptr(CB), get_ptr(SCB32), get_ptr(bias), get_ptr(out), ct.c_float(threshold),
                               ct.c_int32(m), ct.c_int32(n), ct.c_int32(k), get_ptr(ws["CA"]), get_ptr(ws["SCA"]),
                               get_ptr(ws["colflag"]), get_ptr(ws["pos"]), get_ptr(ws["idx"]), ct.c_int32(ws["idx"].numel()),
                               get_ptr(ws["count"]), get_ptr(ws["subA"]), get_ptr(ws["subB"]))
    post_call(prev)
    if rc == 1:
        return None
    if rc != 0:
        raise Exception("cublasLt ran into an error!")
    out = out.reshape(*A.shape[:-1], n)
    if return_quantized:
        return out, ws["CA"], ws["SCA"], ws["idx"], ws["count"]
    return out


def extract_outliers(A, SA, idx):
    """Gather the outlier columns idx out of the formatted int8 weight (reference :2914-2936)."""
    shapeA, formatA = SA[0], SA[1]
    assert formatA in ["col_turing", "col_ampere", "row"]
    _require_cuda(A, "extract_outliers")
    if formatA == "row":
        return A[:, idx.long()].contiguous()
    out = torch.zeros((shapeA[0], idx.numel()), dtype=torch.int8, device=A.device)
    prev = pre_call(A.device)
    fn = lib.cextractOutliers_turing if formatA == "col_turing" else lib.cextractOutliers_ampere
    fn(get_ptr(A), get_ptr(idx), get_ptr(out), ct.c_int32(idx.numel()), ct.c_int32(shapeA[0]), ct.c_int32(shapeA[1]))
    post_call(prev)
    return out
