"""GPU parity: blockwise quantize / dequantize (K1/K2) through the C-ABI vs the CPU oracle.
Bar: bit-exact codes, absmax and dequantised values (integer / single-rounding fp work)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import DT, adversarial_block_values, bits_equal, from_bits, to_bits
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def F():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from bnb_b200 import functional
    return functional


def nf4_thresholds():
    c = json.load(open(os.path.join(GOLDEN, "ref_kernel_constants.json")))
    return [float(v) for v in c["nf4_thresholds_ascending"]]


def fp4_thresholds():
    c = json.load(open(os.path.join(GOLDEN, "ref_kernel_constants.json")))
    t = sorted(float(v) for v in c["fp4_quant_thresholds_tree_order"])
    return [-x for x in reversed(t)] + t


def test_lut_quantiser_equals_reference_tree_for_every_float(F):
    """All 2^32 fp32 bit patterns: shared-memory LUT quantiser == dQuantizeNF4 / dQuantizeFP4 trees."""
    assert F.lib.cbnb_selftest_quant_lut(2) == 0
    assert F.lib.cbnb_selftest_quant_lut(1) == 0


@pytest.mark.parametrize("qtype", ["nf4", "fp4"])
@pytest.mark.parametrize("dtype", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("blocksize", [64, 128, 256, 512, 1024, 2048, 4096])
def test_quantize_4bit_bit_exact(F, qtype, dtype, blocksize):
    torch.manual_seed(blocksize)
    n = blocksize * 37 + 13                      # ragged tail, odd n
    A = (torch.randn(n) * 0.05).to(DT[dtype])
    q, state = F.quantize_4bit(A.cuda(), blocksize=blocksize, quant_type=qtype)
    q_ref, am_ref = orc.quantize_blockwise(to_bits(A), dtype, None, blocksize, qtype)
    assert q.shape == ((n + 1) // 2, 1)
    assert np.array_equal(state.absmax.cpu().numpy().view(np.uint32), am_ref.view(np.uint32))
    assert np.array_equal(q.cpu().numpy().ravel(), q_ref)
    # dequantize: every output dtype, bit-exact
    for od in ("fp32", "fp16", "bf16"):
        out = torch.empty(n, dtype=DT[od], device="cuda")
        F.dequantize_4bit(q, absmax=state.absmax, out=out, blocksize=blocksize, quant_type=qtype)
        ref = orc.dequantize_blockwise(q_ref, am_ref, n, od, None, blocksize, qtype)
        assert bits_equal(out, ref), (od,)


@pytest.mark.parametrize("qtype", ["nf4", "fp4"])
def test_quantize_4bit_adversarial_values(F, qtype):
    thr = nf4_thresholds() if qtype == "nf4" else fp4_thresholds()
    A = adversarial_block_values(64, 64, thr, seed=3).ravel()
    A[64 * 9 + 5] = np.float32("inf")
    A[64 * 10 + 7] = np.float32("nan")
    A[64 * 11:64 * 12] = np.float32(1e-45)       # denormal absmax -> inv = inf
    t = torch.from_numpy(A)
    q, st = F.quantize_4bit(t.cuda(), blocksize=64, quant_type=qtype)
    q_ref, am_ref = orc.quantize_blockwise(A, "fp32", None, 64, qtype)
    assert np.array_equal(st.absmax.cpu().numpy().view(np.uint32), am_ref.view(np.uint32))
    assert np.array_equal(q.cpu().numpy().ravel(), q_ref)


@pytest.mark.parametrize("n", [1, 2, 3, 63, 64, 65, 127, 128, 129, 1000, 4097])
def test_ragged_sizes_and_unaligned_views(F, n):
    torch.manual_seed(n)
    base = torch.randn(n + 3).cuda()
    for off in (0, 1):                           # off=1: 4-byte aligned only -> scalar path
        A = base[off:off + n]
        q, st = F.quantize_4bit(A, blocksize=64, quant_type="nf4")
        q_ref, am_ref = orc.quantize_blockwise(A.cpu().numpy(), "fp32", None, 64, "nf4")
        assert np.array_equal(q.cpu().numpy().ravel(), q_ref)
        assert np.array_equal(st.absmax.cpu().numpy(), am_ref)
        out = F.dequantize_4bit(q, st)
        ref = orc.dequantize_blockwise(q_ref, am_ref, n, "fp32", None, 64, "nf4")
        assert bits_equal(out.reshape(-1), ref)


def test_empty_input(F):
    q, st = F.quantize_4bit(torch.empty(0, device="cuda"), blocksize=64, quant_type="nf4")
    assert q.numel() == 0 and st.absmax.numel() == 0


@pytest.mark.parametrize("dtype", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("blocksize", [64, 256, 4096])
def test_quantize_blockwise_8bit_bit_exact(F, dtype, blocksize):
    torch.manual_seed(1)
    n = blocksize * 21 + 5
    A = torch.randn(n).to(DT[dtype])
    code = F.create_dynamic_map()
    q, st = F.quantize_blockwise(A.cuda(), blocksize=blocksize)
    q_ref, am_ref = orc.quantize_blockwise(to_bits(A), dtype, code.numpy(), blocksize, "8bit")
    assert np.array_equal(st.absmax.cpu().numpy().view(np.uint32), am_ref.view(np.uint32))
    assert np.array_equal(q.cpu().numpy().ravel(), q_ref)
    for od in ("fp32", "fp16", "bf16"):
        out = torch.empty(n, dtype=DT[od], device="cuda")
        F.dequantize_blockwise(q, absmax=st.absmax, code=code.cuda(), out=out, blocksize=blocksize)
        ref = orc.dequantize_blockwise(q_ref, am_ref, n, od, code.numpy(), blocksize, "8bit")
        assert bits_equal(out, ref)


@pytest.mark.parametrize("case,bs", [("bs64", 64), ("bs4096", 4096)])
def test_8bit_against_reference_cpu_golden(F, case, bs):
    """The reference's own quantize_cpu / dequantize_cpu outputs (tests/golden/ref_cpu_blockwise.npz):
    absmax identical, codes agree >= 99.9 % (divide vs reciprocal-multiply), dequantize of the
    reference's codes bit-identical."""
    g = np.load(os.path.join(GOLDEN, "ref_cpu_blockwise.npz"))
    A, q_ref, am_ref, deq_ref = g[case + "_A"], g[case + "_q"], g[case + "_absmax"], g[case + "_deq"]
    q, st = F.quantize_blockwise(torch.from_numpy(A).cuda(), blocksize=bs)
    assert np.array_equal(st.absmax.cpu().numpy(), am_ref)
    qa = q.cpu().numpy()
    assert np.mean(qa == q_ref) >= 0.999 and np.max(np.abs(qa.astype(int) - q_ref.astype(int))) <= 1
    out = F.dequantize_blockwise(torch.from_numpy(q_ref).cuda(), absmax=torch.from_numpy(am_ref).cuda(),
                                 code=torch.from_numpy(g[case + "_code_after"]).cuda(), blocksize=bs)
    assert bits_equal(out, deq_ref)


@pytest.mark.parametrize("qtype", ["nf4", "fp4"])
@pytest.mark.parametrize("dtype", ["fp16", "bf16", "fp32"])
def test_nested_absmax_bit_exact(F, qtype, dtype):
    """compress_statistics=True: offset comes from torch's device mean (fed to the oracle as-is, SURVEY
    Appendix B); uint8 qabsmax, absmax2 and the de-nested dequantised weight must be bit-exact."""
    torch.manual_seed(5)
    W = (torch.randn(256, 1024) * 0.02).to(DT[dtype])
    q, st = F.quantize_4bit(W.cuda(), blocksize=64, compress_statistics=True, quant_type=qtype)
    assert st.nested and st.absmax.dtype == torch.uint8 and st.state2.blocksize == 256
    q_ref, am_ref = orc.quantize_blockwise(to_bits(W).ravel(), dtype, None, 64, qtype)
    assert np.array_equal(q.cpu().numpy().ravel(), q_ref)
    offset = np.float32(st.offset.item())
    am_centered = (am_ref - offset).astype(np.float32)
    code2 = F.create_dynamic_map().numpy()
    qam_ref, am2_ref = orc.quantize_blockwise(am_centered, "fp32", code2, 256, "8bit")
    assert np.array_equal(st.absmax.cpu().numpy(), qam_ref)
    assert np.array_equal(st.state2.absmax.cpu().numpy().view(np.uint32), am2_ref.view(np.uint32))
    deq = F.dequantize_4bit(q, st)
    am_denested = orc.denest_absmax(qam_ref, am2_ref, code2, offset, 256)
    ref = orc.dequantize_blockwise(q_ref, am_denested, W.numel(), dtype, None, 64, qtype)
    assert deq.shape == W.shape and bits_equal(deq, ref)


def test_config1_full_size_roundtrip_properties(F):
    """BASELINE config 1 at full size (4096x4096 fp32, NF4, blocksize 64): size-independent properties.
    (a) dequant(quant(x)) is a fixed point of quantisation (idempotence, bit-exact);
    (b) per-block maximum magnitude is reproduced exactly (codes 0 / 15 are -1 / +1);
    (c) checksum of codes on a 1/64 sample of blocks equals the oracle's."""
    torch.manual_seed(0)
    W = torch.randn(4096, 4096)
    Wg = W.cuda()
    q, st = F.quantize_4bit(Wg, blocksize=64, quant_type="nf4")
    d = F.dequantize_4bit(q, st)
    q2, st2 = F.quantize_4bit(d, blocksize=64, quant_type="nf4")
    d2 = F.dequantize_4bit(q2, st2)
    assert torch.equal(d.view(torch.int32), d2.view(torch.int32))
    assert torch.equal(q, q2)
    blk = Wg.reshape(-1, 64).abs().amax(1)
    assert torch.equal(st.absmax, blk) and torch.equal(d.reshape(-1, 64).abs().amax(1), blk)
    rows = slice(0, 64)                           # 64 rows = 4096 blocks, oracle runs in ms
    q_ref, am_ref = orc.quantize_blockwise(W[rows].numpy().ravel(), "fp32", None, 64, "nf4")
    assert np.array_equal(q.cpu().numpy().ravel()[:q_ref.size], q_ref)
    assert np.array_equal(st.absmax.cpu().numpy()[:am_ref.size], am_ref)
