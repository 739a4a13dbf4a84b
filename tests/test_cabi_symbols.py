"""CPU check: libar_b200.so loads (no libcuda needed at load time) and exports every entry point that
include/ar_b200.h declares; the ctypes table in auto_round_b200/_lib.py covers the same set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ar_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from auto_round_b200 import _lib
    from auto_round_b200.build import build_library

    build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ar_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES) | {"ar_version", "ar_last_error"}
    lib.ar_version.restype = ctypes.c_int
    assert lib.ar_version() == 100


def test_argument_errors_do_not_need_a_gpu():
    """Argument validation happens before any launch: bad args return AR_E_* and set ar_last_error()."""
    from auto_round_b200 import _lib

    lib = _lib.load()
    spec = _lib.QSpec(0, 4, 100, 8, 256, 1e-5, 1.0)   # group_size 100 unsupported
    rc = lib.ar_qdq_fwd(ctypes.byref(spec), None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -2 and b"group_size" in lib.ar_last_error()
    rc = lib.ar_pack_int(None, None, None, 0, 32, 32, 4, 32, 1, None, None, None, None, None)
    assert rc == -1


def test_ops_refuse_cpu_tensors():
    import torch

    from auto_round_b200 import ops

    spec = ops.make_spec("int_sym", 4, 128, 8, 256)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.qdq_fwd(spec, torch.zeros(8, 256, dtype=torch.bfloat16))
