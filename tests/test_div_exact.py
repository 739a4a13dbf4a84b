"""CPU tier: the division-free quotient of the hot kernels (ar::div_exact, auto_round_b200/csrc/ar_qdq_math.cuh) is
bit-identical to the IEEE `w / s` the reference computes (`weight / scale`, auto_round/data_type/int.py:233,292) for EVERY
operand pair that can occur: w a finite bf16 value, s an fp16-valued scale with |s| >= q_scale_thresh (1e-5)."""
import ctypes as C

from test_host_math import host_math  # noqa: F401  (session fixture: compiles tests/host_math/math_host.cpp with g++)


def test_div_exact_all_bf16_by_fp16_pairs(host_math):  # noqa: F811
    host_math.host_div_exact_check.argtypes = [C.POINTER(C.c_long)]
    host_math.host_div_exact_check.restype = C.c_long
    pairs = C.c_long(0)
    bad = host_math.host_div_exact_check(C.byref(pairs))
    assert pairs.value > 4_000_000_000          # 65280 finite bf16 values x ~63k admissible scales
    assert bad == 0
