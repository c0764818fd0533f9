"""The C-ABI boundary: header <-> shared object <-> ctypes binding agree. No GPU needed."""
import ctypes
import re
import subprocess

import fedvit_b200  # noqa: F401
from fedvit_b200 import _lib


def test_header_parses_and_lists_the_path():
    protos = _lib.parse_header()
    for name in ["fv_gemm_bf16", "fv_gemm_f32", "fv_layernorm_fwd", "fv_layernorm_bwd", "fv_attention_fwd",
                 "fv_attention_bwd", "fv_asl_loss", "fv_ce_loss", "fv_adamw_flat", "fv_sumsq",
                 "fv_fedavg_accum", "fv_fedavg_fold_into", "fv_adamw_tick", "fv_cls_grad_rows",
                 "fv_kernel_launches", "fv_patchify", "fv_colsum", "fv_assemble_batch", "fv_mix_batch", "fv_last_error",
                 "fv_launch_count"]:
        assert name in protos, name
    # plain C only: no C++ / torch types may appear in any signature
    allowed = {"int", "int64_t", "float", "size_t", "void*", "const void*", "float*", "const float*",
               "const int64_t*", "int64_t*", "const char*", "const uint8_t*"}
    for name, (ret, args) in protos.items():
        assert ret in allowed, (name, ret)
        for t, _ in args:
            assert t in allowed, (name, t)


def test_library_exports_every_declared_symbol():
    assert _lib.LIB_PATH.exists(), "libfedvit.so missing: run __graft_entry__.build()"
    dll = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in _lib.parse_header():
        assert hasattr(dll, name), f"{name} declared in include/fedvit.h but not exported"


def test_no_undeclared_fv_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fv_\w+)", out))
    assert exported == set(_lib.parse_header()), exported ^ set(_lib.parse_header())


def test_binding_loads_and_reports_version_without_gpu():
    _lib.LIB.load()
    assert _lib.LIB.raw("fv_version")() >= 100
    assert _lib.launch_count() >= 0


def test_library_is_sm100a_tensor_core_code():
    """The shipped GEMM is tcgen05/TMA code, not a recompiled legacy path (SASS mnemonics from
    the profiling guide: UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM = tcgen05.ld)."""
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    if not sass:
        import pytest
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass or "SM100" in sass.upper()
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
