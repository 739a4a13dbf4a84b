"""CPU tier: the division-free quotient of the hot kernels (ar::div_exact, auto_round_b200/csrc/ar_qdq_math.cuh) against the
IEEE `w / s` the reference computes (`weight / scale`, auto_round/data_type/int.py:233,292) for EVERY operand pair that can
occur: w a finite bf16 value, s an fp16-valued scale with |s| >= q_scale_thresh (1e-5).  Bit-identical for every
|w| >= 2^-100; below that (and for +-0) both quotients are < 2^-80 and cannot change round(w/s + V)."""
import ctypes as C

from test_host_math import host_math  # noqa: F401  (session fixture: compiles tests/host_math/math_host.cpp with g++)


def test_div_exact_all_bf16_by_fp16_pairs(host_math):  # noqa: F811
    host_math.host_div_exact_check.argtypes = [C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_float)]
    host_math.host_div_exact_check.restype = C.c_long
    pairs, exact, worst = C.c_long(0), C.c_long(0), C.c_float(0)
    bad = host_math.host_div_exact_check(C.byref(pairs), C.byref(exact), C.byref(worst))
    assert pairs.value > 2_000_000_000
    assert bad == 0
    assert worst.value < 2.0 ** -100            # every pair that is not bit-identical has a numerator far below any real weight
    assert exact.value / pairs.value > 0.75     # (the rest: tiny numerators and signed zeros)
