"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol include/bnb_b200.h
declares (no compute calls), and the host-side mirror matches the golden fixtures."""
import ctypes as ct
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bnb_b200.h")).read()
    return re.findall(r"BNB_B200_API\s+[\w\s\*]+?\b(\w+)\s*\(", src)


def test_header_declares_the_reference_hot_path_abi():
    syms = set(declared_symbols())
    # the 41 hot-path symbols of SURVEY.md section 8b (minus the two CPU-only ones, which need no GPU library)
    want = set()
    for t in ("fp32", "fp16", "bf16"):
        for q in ("", "_fp4", "_nf4"):
            want.add(f"cquantize_blockwise_{t}{q}")
            want.add(f"cdequantize_blockwise_{t}{q}")
        want.add(f"cgemm_4bit_inference_naive_{t}")
    want |= {"cget_col_row_stats", "cdouble_rowcol_quant", "cdequant_mm_int32_fp16", "get_context",
             "cextractOutliers_turing", "cextractOutliers_ampere"}
    for f in ("col32", "turing", "ampere"):
        want |= {f"ctransform_row2{f}", f"ctransform_row2{f}T"}
    for f in ("turing", "ampere"):
        want |= {f"cigemmlt_{f}_32", f"cigemmlt_{f}_8", f"cigemmlt_{f}_8_rowscale"}
    assert len(want) == 39
    assert want <= syms, sorted(want - syms)


def test_library_exports_every_declared_symbol():
    from bnb_b200.cextension import LIB_PATH, lib
    assert os.path.exists(LIB_PATH)
    dll = ct.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(dll, s)]
    assert not missing, missing
    assert lib.cbnb_version().startswith(b"bnb_b200 sm_100a")


def test_library_is_sm100a_tensor_path():
    """cuobjdump (when present) must show tcgen05 / TMA SASS in the shipped library."""
    import shutil
    import subprocess
    from bnb_b200.cextension import LIB_PATH
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCIMMA", "UTCHMMA", "UTMALDG", "LDTM", "STTM", "HMMA"):
        assert mnemonic in sass, mnemonic


def test_stream_setter_roundtrip_without_gpu():
    from bnb_b200.cextension import lib
    lib.cbnb_set_stream(ct.c_void_p(0x1234))
    assert lib.cbnb_get_stream() == 0x1234
    lib.cbnb_set_stream(None)
    assert lib.cbnb_get_stream() in (None, 0)
    assert lib.cbnb_last_error() == 0


def test_python_codebooks_match_reference_golden():
    from bnb_b200 import functional as F
    g = np.load(os.path.join(GOLDEN, "ref_python_tables.npz"))
    assert np.array_equal(F.create_dynamic_map().numpy(), g["dynamic_map"])
    assert np.array_equal(F.get_4bit_type("nf4", device="cpu").numpy(), g["nf4"])
    assert np.array_equal(F.get_4bit_type("fp4", device="cpu").numpy(), g["fp4"])


def test_kernel_constants_match_reference_source():
    """The literals in csrc/codebooks.cuh are the literals of the reference kernel source."""
    consts = json.load(open(os.path.join(GOLDEN, "ref_kernel_constants.json")))
    src = open(os.path.join(ROOT, "bitsandbytes-sycl_b200", "csrc", "codebooks.cuh")).read()

    def macro(name):
        m = re.search(r"#define " + name + r"\s*\\?\s*\{(.*?)\}", src, re.S)
        return [x.strip().rstrip("f") for x in m.group(1).replace("\\", "").split(",") if x.strip()]

    assert [float(x) for x in macro("BNB_NF4_TABLE")] == [float(x) for x in consts["nf4_table"]]
    assert macro("BNB_NF4_THRESHOLDS") == consts["nf4_thresholds_ascending"]
    assert sorted(macro("BNB_FP4_THRESHOLDS"), key=float) == sorted(consts["fp4_quant_thresholds_tree_order"], key=float)
    assert [float(x) for x in macro("BNB_FP4_MAGNITUDES")] == [float(x) for x in consts["fp4_dequant_by_low3bits"]]
    int8_src = open(os.path.join(ROOT, "bitsandbytes-sycl_b200", "csrc", "int8_quant.cu")).read()
    assert consts["mm_dequant_const"] + "f" in int8_src


def test_quantstate_dict_roundtrip_cpu():
    from bnb_b200.functional import QuantState
    s2 = QuantState(absmax=torch.rand(4), blocksize=256, code=torch.rand(256), dtype=torch.float32)
    qs = QuantState(absmax=torch.randint(0, 255, (1024,), dtype=torch.uint8), shape=torch.Size([256, 256]),
                    code=torch.rand(16), blocksize=64, quant_type="nf4", dtype=torch.bfloat16,
                    offset=torch.tensor(0.25), state2=s2)
    packed = qs.as_dict(packed=True)
    assert "quant_state.bitsandbytes__nf4" in packed and all(isinstance(v, torch.Tensor) for v in packed.values())
    back = QuantState.from_dict(dict(packed), device=torch.device("cpu"))
    assert back == qs and back.nested and back.state2.blocksize == 256 and back.dtype == torch.bfloat16


def test_no_cpu_fallback():
    from bnb_b200 import functional as F
    with pytest.raises(NotImplementedError):
        F.quantize_4bit(torch.randn(64), quant_type="nf4")
    with pytest.raises(NotImplementedError):
        F.quantize_blockwise(torch.randn(4096))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bitsandbytes-sycl_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), os.path.join(d, f)


def test_unmodified_reference_loader_and_call_sites_bind():
    """tests/golden/ref_abi_call_sites.json (cut out of the reference by make_golden.py) lists what the reference's own
    loader touches on the CDLL (cextension.py:79-85: restype of get_context / get_cusparse / cget_managed_ptr, and
    hasattr(get_context) at :103) and every lib.<symbol> its hot-path functions call.  Replay both against the .so:
    the loader sequence must not raise, and every call site except the two CPU-fallback symbols must resolve."""
    from bnb_b200.cextension import LIB_PATH
    g = json.load(open(os.path.join(GOLDEN, "ref_abi_call_sites.json")))
    dll = ct.cdll.LoadLibrary(LIB_PATH)
    assert hasattr(dll, "get_context")                      # "only a CUDA-built library exposes this"
    for name in g["loader_attributes"]:                     # SYCLBNBNativeLibrary.__init__
        getattr(dll, name).restype = ct.c_void_p
    assert set(g["loader_attributes"]) >= {"get_context", "get_cusparse", "cget_managed_ptr"}
    cpu_only = {"cquantize_blockwise_cpu_fp32", "cdequantize_blockwise_cpu_fp32"}   # no CPU path here, by design (header note)
    missing = sorted({s for fn, syms in g["functional_call_sites"].items() for s in syms if s not in cpu_only and not hasattr(dll, s)})
    assert not missing, missing
    dll.get_cusparse.restype = ct.c_void_p
    dll.cget_managed_ptr.restype = ct.c_void_p
    dll.cget_managed_ptr.argtypes = [ct.c_size_t]
    assert dll.get_cusparse() is None and dll.cget_managed_ptr(64) is None       # off the hot path: NULL, no allocation


def test_quantstate_wire_format_matches_reference_serialisation():
    """tests/golden/ref_quantstate_packed.npz was written by the reference's OWN QuantState.as_dict(packed=True)
    (functional.py:686-767, class executed from the reference source by make_golden.py).  This repo's QuantState must
    read it (from_dict) and write the same bytes back (as_dict) -- tensors and the JSON blob of the non-tensor items."""
    from bnb_b200.functional import QuantState
    g = np.load(os.path.join(GOLDEN, "ref_quantstate_packed.npz"))
    ref = {k: torch.from_numpy(g[k].copy()) for k in g.files if k != "weight"}
    st = QuantState.from_dict({k: v.clone() for k, v in ref.items()}, device=torch.device("cpu"))
    assert st.nested and st.quant_type == "nf4" and st.blocksize == 64 and st.dtype == torch.bfloat16
    assert tuple(st.shape) == (64, 256) and st.state2.blocksize == 256 and st.state2.dtype == torch.float32
    assert st.absmax.dtype == torch.uint8 and torch.equal(st.absmax, ref["absmax"])
    assert torch.equal(st.state2.absmax, ref["nested_absmax"]) and torch.equal(st.state2.code, ref["nested_quant_map"])
    assert torch.equal(st.code, ref["quant_map"]) and float(st.offset) == 0.017303466796875
    mine = st.as_dict(packed=True)
    assert sorted(mine) == sorted(ref)
    for k in ref:
        assert torch.equal(mine[k], ref[k]), k          # byte-identical, including the packed JSON


@pytest.mark.parametrize("fmt", ["col32", "col_turing", "col_ampere"])
@pytest.mark.parametrize("shape", [(64, 64), (40, 96), (129, 160)])
def test_undo_layout_to_row_inverts_the_reference_layouts(fmt, shape):
    """Checkpoint load path (reference nn/modules.py:635-654 maybe_rearrange_weight -> undo_layout): the host-side inverse
    of col32 / col_turing / col_ampere against the oracle's forward maps (kernel_quant.cpp:3673-3832)."""
    from bnb_b200 import functional as F
    from oracle import oracle as orc
    rows, cols = shape
    rng = np.random.RandomState(rows + cols)
    A = rng.randint(-128, 128, (rows, cols)).astype(np.int8)
    buf = orc.transform(A, fmt)
    pr = {"col32": rows, "col_turing": (rows + 7) // 8 * 8, "col_ampere": (rows + 31) // 32 * 32}[fmt]
    pc = (cols + 31) // 32 * 32
    assert buf.size == pr * pc
    back = F.undo_layout_to_row(torch.from_numpy(buf.reshape(pr, pc)), fmt, rows, cols)
    assert np.array_equal(back.numpy(), A)
