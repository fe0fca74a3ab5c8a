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
  * ref_device_trees.npz -- outputs of the reference's scalar DEVICE functions (dQuantizeNF4, dQuantizeFP4,
    dQuantize<0>, dDequantizeNF4, dDequantizeFP4Tree; kernel_quant.cpp:519-837) whose bodies are cut out of the
    reference source verbatim and compiled with g++ in a temporary directory, on ~260 k seeded inputs plus every
    decision threshold +- 1 ulp.  Inputs are regenerated from the seed by the test (tree_inputs()).
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


def tree_inputs():
    """The inputs of ref_device_trees.npz, regenerated identically by tests/test_oracle.py: seeded values in and around
    [-1, 1] plus, for every decision threshold of the three quantisers, the threshold itself and its two neighbours."""
    tables = np.load(os.path.join(HERE, "ref_python_tables.npz"))
    consts = json.load(open(os.path.join(HERE, "ref_kernel_constants.json")))
    rng = np.random.RandomState(20241)
    x = np.concatenate([rng.uniform(-1.05, 1.05, 200000), rng.randn(40000) * 0.3, rng.uniform(-0.01, 0.01, 20000)]).astype(np.float32)
    code = tables["dynamic_map"].astype(np.float32)
    thr = [np.float32(v) for v in consts["nf4_thresholds_ascending"]]
    thr += [np.float32(v) for v in consts["fp4_quant_thresholds_tree_order"]] + [-np.float32(v) for v in consts["fp4_quant_thresholds_tree_order"]]
    thr += list(code) + list(((code[1:].astype(np.float64) + code[:-1]) * 0.5).astype(np.float32))     # 8-bit pivots and midpoints
    thr = np.array(thr, np.float32)
    special = np.concatenate([thr, np.nextafter(thr, np.float32(2)), np.nextafter(thr, np.float32(-2)),
                              np.array([0.0, -0.0, 1.0, -1.0, 1e-42, -3e-41, 1.5, -1.5, np.inf, -np.inf, np.nan], np.float32)])
    return np.concatenate([x, special.astype(np.float32)]), code


def device_trees():
    """Reference-EXECUTED golden vectors for the scalar device functions: the bodies of dDequantizeFP4Tree, dQuantizeFP4,
    dDequantizeNF4, dQuantizeNF4 and dQuantize<0> (kernel_quant.cpp:519-837, plain scalar C++) are cut out of the
    reference source verbatim, compiled with g++ under /tmp (never written into this repo) and run on tree_inputs()."""
    import ctypes as ct
    import subprocess
    import tempfile
    src = open(os.path.join(REF, "sycl/sycl_code/kernel_quant.cpp")).read()
    sigs = [r"float dDequantizeFP4Tree\(unsigned char val, float absmax\)", r"unsigned char dQuantizeFP4\(float x\)",
            r"float dDequantizeNF4\(unsigned char val\)", r"unsigned char dQuantizeNF4\(float x\)",
            r"template <int STOCHASTIC>\s*unsigned char dQuantize\(float\* smem_code, const float rand, float x\)"]
    parts = []
    for sg in sigs:
        m = re.search(sg, src)
        assert m, sg
        parts.append(m.group(0) + "\n" + body_of(src, sg))
    shim = ("#include <cmath>\nnamespace sycl { static inline float fabs(float v) { return std::fabs(v); } }\n" + "\n".join(parts) +
            "\nextern \"C\" {\n"
            "void run_nf4(const float* x, unsigned char* q, long n) { for (long i = 0; i < n; i++) q[i] = dQuantizeNF4(x[i]); }\n"
            "void run_fp4(const float* x, unsigned char* q, long n) { for (long i = 0; i < n; i++) q[i] = dQuantizeFP4(x[i]); }\n"
            "void run_q8(float* code, const float* x, unsigned char* q, long n) { for (long i = 0; i < n; i++) q[i] = dQuantize<0>(code, 0.0f, x[i]); }\n"
            "void run_deq(float* nf4, float* fp4, float absmax) { for (int i = 0; i < 16; i++) { nf4[i] = dDequantizeNF4((unsigned char)i); fp4[i] = dDequantizeFP4Tree((unsigned char)i, absmax); } }\n"
            "}\n")
    with tempfile.TemporaryDirectory(prefix="bnb_ref_trees_") as td:
        cpp, so = os.path.join(td, "trees.cpp"), os.path.join(td, "trees.so")
        open(cpp, "w").write(shim)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp])
        lib = ct.CDLL(so)
        x, code = tree_inputs()
        n = x.size
        out = {}
        fp = lambda a: a.ctypes.data_as(ct.c_void_p)
        for name, fn in (("nf4", lib.run_nf4), ("fp4", lib.run_fp4)):
            q = np.zeros(n, np.uint8)
            fn(fp(x), fp(q), ct.c_long(n))
            out["q_" + name] = q
        q = np.zeros(n, np.uint8)
        code_c = np.ascontiguousarray(code)
        finite = np.isfinite(x)                           # the 8-bit search is only defined for finite inputs
        xf = np.ascontiguousarray(x[finite])
        qf = np.zeros(xf.size, np.uint8)
        lib.run_q8(fp(code_c), fp(xf), fp(qf), ct.c_long(xf.size))
        out["q_8bit_finite"] = qf
        nf4 = np.zeros(16, np.float32)
        fp4 = np.zeros(16, np.float32)
        lib.run_deq(fp(nf4), fp(fp4), ct.c_float(0.73))
        out["deq_nf4"], out["deq_fp4_absmax0p73"] = nf4, fp4
    out["n_inputs"] = np.array([n], np.int64)
    out["x_crc"] = np.array([int(np.bitwise_xor.reduce(x.view(np.uint32)))], np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_device_trees.npz"), **out)
    print("device tree fixtures:", {k: (v.shape, v.dtype) for k, v in out.items()}, "bytes", os.path.getsize(os.path.join(HERE, "ref_device_trees.npz")))


def loader_and_call_sites():
    """ref_abi_call_sites.json: every `lib.<symbol>` the reference's hot-path Python functions call
    (python_src_quants/functional.py) and every attribute the reference LOADER touches on the CDLL
    (python_src_quants/cextension.py:79-85, 103) -- what an unmodified reference needs from a replacement .so."""
    src = open(os.path.join(REF, "python_src_quants/functional.py")).read()
    hot = {"quantize_blockwise", "dequantize_blockwise", "quantize_4bit", "dequantize_4bit", "gemv_4bit", "get_colrow_absmax",
           "double_quant", "transform", "igemmlt", "mm_dequant", "extract_outliers"}
    calls = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in hot:
            calls[node.name] = sorted(set(re.findall(r"\blib\.(\w+)", ast.get_source_segment(src, node))))
    lsrc = open(os.path.join(REF, "python_src_quants/cextension.py")).read()
    loader = sorted(set(re.findall(r"\blib\.(\w+)\.restype", lsrc)) | set(re.findall(r'hasattr\(dll, "(\w+)"\)', lsrc)))
    json.dump({"functional_call_sites": calls, "loader_attributes": loader,
               "source": ["python_src_quants/functional.py", "python_src_quants/cextension.py"]},
              open(os.path.join(HERE, "ref_abi_call_sites.json"), "w"), indent=1)
    print("loader touches:", loader)


def checkpoint_fixture():
    """ref_quantstate_packed.npz: a nested NF4 QuantState serialised by the REFERENCE's own QuantState.as_dict(packed=True)
    (python_src_quants/functional.py:625-798, class cut out with ast; pack_dict_to_tensor from utils.py:169-183) -- the
    wire format of Linear4bit._save_to_state_dict (nn/modules.py:436-445)."""
    fsrc = open(os.path.join(REF, "python_src_quants/functional.py")).read()
    usrc = open(os.path.join(REF, "python_src_quants/utils.py")).read()
    ns = {"torch": torch, "Tensor": torch.Tensor, "Dict": __import__("typing").Dict, "Any": __import__("typing").Any, "json": json}
    for node in ast.parse(usrc).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("pack_dict_to_tensor", "unpack_tensor_to_dict"):
            exec(ast.get_source_segment(usrc, node), ns)
    for node in ast.parse(fsrc).body:
        if isinstance(node, ast.ClassDef) and node.name == "QuantState":
            exec(ast.get_source_segment(fsrc, node), ns)
    QS = ns["QuantState"]
    tables = np.load(os.path.join(HERE, "ref_python_tables.npz"))
    g = torch.Generator().manual_seed(77)
    N, K = 64, 256
    nblocks = N * K // 64
    s2 = QS(absmax=torch.rand(nblocks // 256, generator=g) * 0.01 + 0.001, blocksize=256, code=torch.from_numpy(tables["dynamic_map"].copy()),
            dtype=torch.float32)
    st = QS(absmax=torch.randint(0, 256, (nblocks,), generator=g, dtype=torch.uint8), shape=torch.Size((N, K)),
            code=torch.from_numpy(tables["nf4"].copy()), blocksize=64, quant_type="nf4", dtype=torch.bfloat16,
            offset=torch.tensor(0.017303466796875), state2=s2)
    packed = st.as_dict(packed=True)
    out = {k: v.numpy() for k, v in packed.items()}
    out["weight"] = torch.randint(0, 256, (N * K // 2, 1), generator=g, dtype=torch.uint8).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_quantstate_packed.npz"), **out)
    # round trip through the reference's own from_dict as a self-check
    back = QS.from_dict({k: v.clone() for k, v in packed.items()}, device=torch.device("cpu"))
    assert back == st and back.offset.item() == st.offset.item()
    print("quantstate fixture keys:", sorted(out))


if __name__ == "__main__":
    assert os.path.isdir(REF), "run this where /root/reference exists"
    checkpoint_fixture()
    loader_and_call_sites()
    python_tables()
    kernel_constants()
    cpu_blockwise()
    device_trees()
