# Derived from python_src_quants/autograd/_functions.py of abhilash1910/bitsandbytes-SYCL (itself bitsandbytes,
# Copyright (c) Facebook, Inc. and its affiliates, MIT license -- see the LICENSE file of that repository).
# This file keeps the reference's public interface (MatMul4Bit / MatMul8bitLt / MatmulLtState, their dispatch rules and forward/backward step order) because it is
# the wire / API contract of the drop-in; "xpu" became "cuda" and everything below the interface calls the native
# sm_100a library.  It is a derived host-side shim, not from-scratch work -- the from-scratch work is csrc/.
"""Forward paths of the reference's python_src_quants/autograd/_functions.py, re-hosted on the B200 kernels:
MatMul4Bit (:486-540), MatMul8bitLt (:288-483), MatmulLtState (:246-285), matmul (:543-554),
matmul_4bit (:557-577).  Dispatch rules are the reference's; what runs underneath is native:
  * batch 1  -> F.gemv_4bit            (one fused launch, nested absmax consumed in-kernel)
  * batch >1 -> F.gemm_4bit            (fused dequant + tcgen05 GEMM; reference: dequantize_4bit + F.linear)
  * int8     -> double_quant -> ONE fused tcgen05 kind::i8 GEMM with the mm_dequant epilogue
                (reference: transform + igemmlt + mm_dequant = 3 launches and a 400 MB int32 round trip),
                outliers (threshold > 0) via the 16-bit side GEMM exactly as the reference does.
Backward: MatMul4Bit keeps the reference's dequant + matmul backward; MatMul8bitLt.backward follows the
reference's step order (:436-483) on the row-major int8 GEMM.
"""
import warnings
from dataclasses import dataclass
from functools import reduce
from typing import Optional
from warnings import warn

import torch

from .. import functional as F


def prod(iterable):
    return reduce(lambda a, b: a * b, iterable, 1)


def supports_igemmlt(device: torch.device) -> bool:
    """reference :218-228 gates on compute capability >= 7.5; here the kernels are sm_100a only."""
    return torch.cuda.get_device_capability(device=device)[0] >= 10


@dataclass
class MatmulLtState:
    """reference :246-285."""
    _tile_indices: Optional[torch.Tensor] = None
    force_no_igemmlt: bool = False
    CB = None
    CxB = None
    SB = None
    SCB = None
    CxBt = None
    SBt = None
    CBt = None
    subB = None
    outlier_pool = None
    has_accumulated_gradients = False
    threshold = 0.0
    idx = None
    is_training = True
    has_fp16_weights = True
    memory_efficient_backward = False
    use_pool = False
    formatB = F.get_special_format_str()

    def reset_grads(self):
        self.CB = None
        self.CxB = None
        self.SB = None
        self.SCB = None
        self.CxBt = None
        self.SBt = None
        self.CBt = None


class MatMul8bitLt(torch.autograd.Function):
    """LLM.int8 forward (reference :292-434).  Steps kept in the reference's order:
    1. double_quant(A, threshold)  2. quantise / fetch B  3. int8 GEMM + dequant  4. outlier side-GEMM."""

    @staticmethod
    def forward(ctx, A, B, out=None, bias=None, state=MatmulLtState):
        ctx.is_empty = False
        if prod(A.shape) == 0:
            ctx.is_empty = True
            ctx.A, ctx.B, ctx.bias = A, B, bias
            if A.shape[-1] == B.shape[0]:
                return torch.empty(A.shape[:-1] + B.shape[1:], dtype=A.dtype, device=A.device)
            return torch.empty(A.shape[:-1] + B.shape[:1], dtype=A.dtype, device=A.device)

        formatB = state.formatB
        input_shape = A.shape
        if A.dtype != torch.float16:
            warnings.warn(f"MatMul8bitLt: inputs will be cast from {A.dtype} to float16 during quantization")

        # 0. inference fast path (additive): the whole forward in one native call, no host synchronisation.  Taken when
        # nothing needs a gradient, the weight is already int8 row-major (CB / SCB) and threshold > 0.
        if (F.FUSED_INT8_LINEAR and not any(ctx.needs_input_grad[:2]) and not state.has_fp16_weights and state.threshold > 0.0
                and A.dtype == torch.float16 and state.SCB is not None and (bias is None or bias.dtype == torch.float16)):
            CBrow = state.CB if state.CB is not None else (state.CxB if (state.SB and state.SB[1] == "row") else None)
            if CBrow is not None and CBrow.dtype == torch.int8 and CBrow.dim() == 2:
                fused = F.int8_linear_fused(A, CBrow, state.SCB, bias=bias, threshold=state.threshold)
                if fused is not None:
                    ctx.state = state
                    ctx.formatB = formatB
                    ctx.grad_shape = input_shape
                    ctx.dtype_A, ctx.dtype_B, ctx.dtype_bias = A.dtype, B.dtype, None if bias is None else bias.dtype
                    ctx.tensors = [None, None, None]
                    ctx.tensor_states = (None, None)
                    return fused

        # 1. quantise A (row- and column-wise int8 + outlier COO)
        if len(A.shape) == 3:
            A = A.reshape(-1, A.shape[-1])
        CA, CAt, SCA, SCAt, coo_tensorA = F.double_quant(A.to(torch.float16), threshold=state.threshold)

        subA = None
        if state.threshold > 0.0 and coo_tensorA is not None:
            if state.has_fp16_weights:
                idx = torch.unique(coo_tensorA.colidx).long()
                CA[:, idx] = 0
                CAt[:, idx] = 0
                subA = A[:, idx]
                state.subB = B[:, idx].t().contiguous()
                state.idx = idx
            elif state.CxB is None:
                state.CxB, state.SB = F.transform(state.CB, to_order=formatB)
        elif not state.has_fp16_weights and state.CxB is None:
            state.CxB, state.SB = F.transform(state.CB, to_order=formatB)

        # 2. quantise B when it is still a 16-bit weight
        if state.has_fp16_weights:
            has_grad = getattr(B, "grad", None) is not None
            if not B.is_contiguous() and B.shape[0] == B.stride(1):
                B = B.contiguous()
            if (state.is_training and not has_grad) or state.CxB is None:
                state.reset_grads()
                CB, state.CBt, state.SCB, state.SCBt, _ = F.double_quant(B.to(torch.float16))
                state.CxB, state.SB = F.transform(CB, to_order=formatB)

        if coo_tensorA is not None and not state.has_fp16_weights:
            # outlier columns: dequantised weight slice for the 16-bit side GEMM (reference :369-384)
            outlier_idx = torch.unique(coo_tensorA.colidx)
            state.idx = outlier_idx
            outliers = F.extract_outliers(state.CxB, state.SB, state.idx.int())
            state.subB = (outliers * state.SCB.view(-1, 1) / 127.0).t().contiguous().to(A.dtype)
            CA[:, state.idx.long()] = 0
            CAt[:, state.idx.long()] = 0
            subA = A[:, state.idx.long()]

        shapeB = state.SB[0] if state.SB else B.shape
        if len(input_shape) == 3:
            output_shape = (input_shape[0], input_shape[1], shapeB[0])
        else:
            output_shape = (input_shape[0], shapeB[0])

        # 3. int8 GEMM + dequant
        fused_bias = bias if (bias is None or bias.dtype == torch.float16) else None
        if state.SB[1] == "row":
            output = F.int8_linear_dequant(CA, state.CxB, SCA, state.SCB, bias=fused_bias)
        else:
            C32A, SA = F.transform(CA, "col32")
            out32, Sout32 = F.igemmlt(C32A, state.CxB, SA, state.SB)
            output = F.mm_dequant(out32, Sout32, SCA, state.SCB, bias=fused_bias)
        output = output.to(A.dtype)
        if bias is not None and fused_bias is None:
            output = output.add_(bias)

        # 4. mixed-precision decomposition: outlier columns in 16 bit
        if coo_tensorA is not None and subA is not None:
            output += torch.matmul(subA, state.subB)

        ctx.state = state
        ctx.formatB = formatB
        ctx.grad_shape = input_shape
        ctx.dtype_A, ctx.dtype_B, ctx.dtype_bias = A.dtype, B.dtype, None if bias is None else bias.dtype
        if any(ctx.needs_input_grad[:2]):
            ctx.tensors = (CAt, subA, A)
            ctx.tensor_states = (SCAt, state.idx)
        else:
            ctx.tensors = [None, None, A]
            ctx.tensor_states = (None, None)
            ctx.save_for_backward(None, None)
        clone_func = torch.clone if len(output_shape) == 3 else lambda x: x
        return clone_func(output.view(output_shape))

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.is_empty:
            bias_grad = None if ctx.bias is None else torch.zeros_like(ctx.bias)
            return torch.zeros_like(ctx.A), torch.zeros_like(ctx.B), None, bias_grad, None
        # reference :436-483.  Same quantities, but the two int8 products run on the row-major tcgen05 GEMM
        # (C[i, j] = sum_k A[i, k] * B[j, k]) instead of col32 / col_turing operands: the operand that the reference
        # re-lays out with transform(..., transpose=True) is simply transposed here.
        req_gradA, req_gradB, _, req_gradBias, _ = ctx.needs_input_grad
        CAt, subA, A = ctx.tensors
        SCAt, idx = ctx.tensor_states
        state = ctx.state
        grad_A = grad_B = grad_bias = None
        if req_gradBias:
            grad_bias = grad_output.sum(0, dtype=ctx.dtype_bias)
        if len(grad_output.shape) == 3:
            grad_output = grad_output.reshape(-1, grad_output.shape[-1]).contiguous()
        Cgrad, Cgradt, SCgrad, SCgradt, _ = F.double_quant(grad_output.to(torch.float16))
        if req_gradB:
            # grad_B[j, c] = sum_i grad[i, j] * A[i, c], both quantised column-wise (statistics over the token axis)
            grad_B = F.int8_linear_dequant(Cgradt.t().contiguous(), CAt.t().contiguous(), SCgradt, SCAt)
            if state.threshold > 0.0 and subA is not None:
                grad_B[:, idx] += torch.matmul(grad_output.t(), subA)
        if req_gradA:
            if state.CBt is not None:
                # grad_A[i, c] = sum_j grad[i, j] * B[j, c]: grad quantised row-wise, B column-wise
                if state.CxBt is None:
                    state.CxBt, state.SBt = state.CBt.t().contiguous(), (tuple(state.CBt.t().shape), "row")
                grad_A = F.int8_linear_dequant(Cgrad, state.CxBt, SCgrad, state.SCBt).view(ctx.grad_shape).to(ctx.dtype_A)
            else:
                CBrow = state.CB if state.CB is not None else (state.CxB if (state.SB and state.SB[1] == "row") else None)
                if CBrow is None:
                    raise Exception("State must contain either CBt or CB or CxB matrix for backward")
                CB = CBrow.to(ctx.dtype_A, copy=True).mul_(state.SCB.unsqueeze(1).mul(1.0 / 127.0))
                grad_A = torch.matmul(grad_output.to(ctx.dtype_A), CB).view(ctx.grad_shape).to(ctx.dtype_A)
        return grad_A, grad_B, None, grad_bias, None


class MatMul4Bit(torch.autograd.Function):
    """reference :486-540.  Forward: fused 4-bit GEMM when the native kernel takes the shape, otherwise
    the reference's own route (dequantize_4bit on the GPU + F.linear)."""

    @staticmethod
    def forward(ctx, A, B, out=None, bias=None, quant_state: Optional[F.QuantState] = None):
        ctx.is_empty = False
        if prod(A.shape) == 0:
            ctx.is_empty = True
            ctx.A, ctx.B, ctx.bias = A, B, bias
            B_shape = quant_state.shape
            if A.shape[-1] == B_shape[0]:
                return torch.empty(A.shape[:-1] + B_shape[1:], dtype=A.dtype, device=A.device)
            return torch.empty(A.shape[:-1] + B_shape[:1], dtype=A.dtype, device=A.device)

        # the fused kernel is the forward in training too: its operand is bit-identical to dequantize_4bit's output
        output = F.gemm_4bit(A, B, quant_state, bias=bias)
        if output is None:
            output = torch.nn.functional.linear(A, F.dequantize_4bit(B, quant_state).to(A.dtype).t(), bias)

        ctx.state = quant_state
        ctx.dtype_A, ctx.dtype_B, ctx.dtype_bias = A.dtype, B.dtype, None if bias is None else bias.dtype
        ctx.tensors = (None, B) if any(ctx.needs_input_grad[:2]) else (None, None)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.is_empty:
            bias_grad = None if ctx.bias is None else torch.zeros_like(ctx.bias)
            return torch.zeros_like(ctx.A), torch.zeros_like(ctx.B), None, bias_grad, None
        req_gradA, _, _, req_gradBias, _ = ctx.needs_input_grad
        _, B = ctx.tensors
        grad_A, grad_B, grad_bias = None, None, None
        if req_gradBias:
            grad_bias = grad_output.sum(0, dtype=ctx.dtype_bias)
        if req_gradA:
            grad_A = torch.matmul(grad_output, F.dequantize_4bit(B, ctx.state).to(grad_output.dtype).t())
        return grad_A, grad_B, None, grad_bias, None


def matmul(A: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None,
           state: Optional[MatmulLtState] = None, threshold=0.0, bias=None):
    """reference :543-554."""
    state = state or MatmulLtState()
    if threshold > 0.0:
        state.threshold = threshold
    return MatMul8bitLt.apply(A, B, out, bias, state)


def matmul_4bit(A: torch.Tensor, B: torch.Tensor, quant_state: F.QuantState, out: Optional[torch.Tensor] = None,
                bias=None):
    """reference :557-577: GEMV iff A is a single row, needs no grad and K % blocksize == 0."""
    assert quant_state is not None
    if A.numel() == A.shape[-1] and A.requires_grad == False:  # noqa: E712
        if A.shape[-1] % quant_state.blocksize != 0:
            warn(f"Some matrices hidden dimension is not a multiple of {quant_state.blocksize} and efficient "
                 f"inference kernels are not supported for these (slow). Matrix input size found: {A.shape}")
            return MatMul4Bit.apply(A, B, out, bias, quant_state)
        out = F.gemv_4bit(A, B.t(), out, state=quant_state)
        if bias is not None:
            out += bias
        return out
    return MatMul4Bit.apply(A, B, out, bias, quant_state)


def matmul_4bit_multi(A: torch.Tensor, Bs, quant_states, outs=None, biases=None):
    """ADDITIVE (inference): matmul_4bit for several 4-bit weights that share the activation row A -- the q/k/v or
    gate/up projections of a decoder layer -- in one GEMV launch (functional.gemv_4bit_multi).  Falls back to one
    matmul_4bit per weight whenever matmul_4bit itself would not take the GEMV route (batch > 1, gradients, K not a
    multiple of the blocksize)."""
    n = len(Bs)
    outs = list(outs) if outs is not None else [None] * n
    biases = list(biases) if biases is not None else [None] * n
    if A.numel() == A.shape[-1] and A.requires_grad == False and all(A.shape[-1] % st.blocksize == 0 for st in quant_states):  # noqa: E712
        res = F.gemv_4bit_multi(A, [B.t() for B in Bs], quant_states, outs=outs)
        for r, b in zip(res, biases):
            if b is not None:
                r += b
        return res
    return [matmul_4bit(A, B, st, out=o, bias=b) for B, st, o, b in zip(Bs, quant_states, outs, biases)]
