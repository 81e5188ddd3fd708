"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU and exports
every symbol include/stablets_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "stablets_b200.h")).read()
    return sorted(set(re.findall(r"STB_API[^;(]*?\b(stb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_declared_symbols():
    import importlib.util
    spec = importlib.util.spec_from_file_location("stb_build", os.path.join(ROOT, "stable-ts_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    so = mod.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/stablets_b200.h but not exported"
    lib.stb_abi_version.restype = ctypes.c_int
    assert lib.stb_abi_version() >= 1


def test_python_binding_covers_header():
    from stable_ts_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared()


def test_sass_has_blackwell_tensor_and_tma_instructions():
    so = os.path.join(ROOT, "stable-ts_b200", "libstablets_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out and "UTMALDG" in out and "LDTM" in out


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stable_ts_b200.model import B200Whisper
    with pytest.raises(RuntimeError):
        B200Whisper(None, {}, device="cuda")


def test_option_switches_and_error_reporting_without_a_gpu():
    """stb_set_option / stb_get_option / stb_last_error work without a device: defaults of the kernel-variant switches, unknown
    names are an error with a message (include/stablets_b200.h)."""
    from stable_ts_b200 import _lib as L
    lib = L.lib()
    assert lib.stb_abi_version() >= 2
    assert L.get_option("xattn_tc") in (0, 1) and L.get_option("decode_splitk_legacy") in (0, 1)
    assert L.get_option("decode_fused_ln") in (0, 1) and L.get_option("no_such_option") == -1
    old = L.get_option("decode_splitk_legacy")
    L.set_option("decode_splitk_legacy", 1 - old)
    assert L.get_option("decode_splitk_legacy") == 1 - old
    L.set_option("decode_splitk_legacy", old)
    try:
        L.set_option("no_such_option", 1)
    except L.StbError as e:
        assert "no_such_option" in str(e)
    else:
        raise AssertionError("unknown option accepted")
    # argument validation happens before any CUDA call
    assert lib.stb_decode_state_bytes(None, 4) == 0 and lib.stb_decode_ws_bytes(None, 4) == 0
    assert lib.stb_axpby(None, None, 1.0, 0.0, 8, None) != 0 and b"stb_axpby" in lib.stb_last_error()
