#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the authoring container (where /root/reference exists):
    python tests/golden/make_golden.py

What it captures (nothing here is typed in by hand):
  * ref_python_tables.npz  -- `create_dynamic_map()` and `get_4bit_type("nf4"|"fp4")` executed from the
    reference's own python_src_quants/functional.py source (functions are cut out with `ast` because the
    module itself imports intel_extension_for_pytorch and cannot be imported here).
  * ref_kernel_constants.json -- the literal constants parsed out of sycl/sycl_code/kernel_quant.cpp:
    dDequantizeNF4 leaves (:650-703), dQuantizeNF4 thresholds (:705-756), dQuantizeFP4 thresholds
    (:547-594), dDequantizeFP4Tree leaves (:520-545), MM_DEQUANT_CONST (:3846).
  * ref_cpu_blockwise.npz -- inputs and outputs of the reference's compiled sycl/cpu_ops.cpp
    (oracle/_ref/libref_cpu.so): quantize_cpu / dequantize_cpu on seeded data at blocksize 64 and 4096,
    including a ragged tail and a run of zeros.
The fixtures travel to the GPU box; /root/reference does not.
"""
import ast
import json
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def cut_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
    return out


def python_tables():
    funcs = cut_functions(os.path.join(REF, "python_src_quants/functional.py"),
                          {"create_dynamic_map", "get_4bit_type", "create_fp8_map", "create_linear_map"})
    ns = {"torch": torch, "Tensor": torch.Tensor, "itertools": __import__("itertools"), "np": np}
    for name, code in funcs.items():
        exec(code, ns)
    dyn = ns["create_dynamic_map"]().numpy().astype(np.float32)
    nf4 = ns["get_4bit_type"]("nf4", device="cpu").numpy().astype(np.float32)
    fp4 = ns["get_4bit_type"]("fp4", device="cpu").numpy().astype(np.float32)
    np.savez(os.path.join(HERE, "ref_python_tables.npz"), dynamic_map=dyn, nf4=nf4, fp4=fp4)
    print("dynamic map: n=%d first=%r last=%r" % (dyn.size, dyn[0], dyn[-1]))


def body_of(src, signature_regex):
    m = re.search(signature_regex, src)
    assert m, signature_regex
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return src[i:j + 1]


def kernel_constants():
    src = open(os.path.join(REF, "sycl/sycl_code/kernel_quant.cpp")).read()
    num = r"(-?\d+\.\d+(?:e[-+]?\d+)?)f"
    deq_nf4 = body_of(src, r"float dDequantizeNF4\(unsigned char val\)")
    q_nf4 = body_of(src, r"unsigned char dQuantizeNF4\(float x\)")
    q_fp4 = body_of(src, r"unsigned char dQuantizeFP4\(float x\)")
    deq_fp4 = body_of(src, r"float dDequantizeFP4Tree\(unsigned char val, float absmax\)")
    mmc = re.search(r"#define MM_DEQUANT_CONST (\S+?)f", src).group(1)
    leaves = [(m.group(1)) for m in re.finditer(r"return " + num, deq_nf4)]
    # leaves appear in tree order 1111,1110,...,1000 then 0111 ... 0000 -> index = 15 - position
    nf4_by_index = [None] * 16
    for pos, v in enumerate(leaves):
        nf4_by_index[15 - pos] = v
    thresholds = sorted({m.group(1) for m in re.finditer(r"x > " + num, q_nf4)}, key=float)
    fp4_thr = [m.group(1) for m in re.finditer(r"x > " + num, q_fp4)]
    # dDequantizeFP4Tree leaves: comment holds the nibble, e.g. "return 0.25000000f*absmax*sign; // 1111"
    fp4_leaves = {}
    for m in re.finditer(r"return " + num + r"\*absmax\*sign; // 1(\d\d\d)", deq_fp4):
        fp4_leaves[int(m.group(2), 2)] = m.group(1)
    out = {
        "nf4_table": nf4_by_index,
        "nf4_thresholds_ascending": thresholds,
        "fp4_quant_thresholds_tree_order": fp4_thr,
        "fp4_dequant_by_low3bits": [fp4_leaves[i] for i in range(8)],
        "mm_dequant_const": mmc,
        "source": "sycl/sycl_code/kernel_quant.cpp",
    }
    assert len(leaves) == 16 and len(thresholds) == 15 and len(fp4_thr) == 7, (len(leaves), len(thresholds), len(fp4_thr))
    json.dump(out, open(os.path.join(HERE, "ref_kernel_constants.json"), "w"), indent=1)
    print("kernel constants ok:", out["mm_dequant_const"])


def cpu_blockwise():
    from oracle import oracle as orc
    tables = np.load(os.path.join(HERE, "ref_python_tables.npz"))
    code = tables["dynamic_map"]
    rng = np.random.RandomState(1234)
    cases = {}
    for name, n, bs in (("bs64", 64 * 40 + 37, 64), ("bs4096", 4096 * 2 + 1000, 4096)):
        A = rng.randn(n).astype(np.float32)
        # NOTE: no all-zero block here -- the reference's quantize_cpu divides by absmax == 0 and its
        # BinSearch then indexes out of bounds on the NaN (observed SIGSEGV), so that input is undefined
        # for the reference; zero runs INSIDE a non-zero block are fine.
        A[64:100] = 0.0
        A[200] = 3.5
        A[201] = -3.5
        q, absmax, code_after = orc.quantize_cpu_reference(code, A, bs)
        deq = orc.dequantize_cpu_reference(code_after, q, absmax, bs)
        cases[name + "_A"] = A
        cases[name + "_q"] = q
        cases[name + "_absmax"] = absmax
        cases[name + "_deq"] = deq
        cases[name + "_code_after"] = code_after
    np.savez_compressed(os.path.join(HERE, "ref_cpu_blockwise.npz"), **cases)
    print("cpu blockwise fixtures:", {k: v.shape for k, v in cases.items()})


if __name__ == "__main__":
    assert os.path.isdir(REF), "run this where /root/reference exists"
    python_tables()
    kernel_constants()
    cpu_blockwise()
