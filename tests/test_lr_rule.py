"""Host logic (CPU): the learning-rate rule and schedule the product uploads to the device equal the recursion the oracle
runs -- which tests/test_oracle_golden.py pins bit-exactly against 8- and 1000-iteration runs of the reference."""
import pytest
import torch

from auto_round_b200.quantizer import SignRoundQuantizer, lr_schedule_table
from auto_round_b200.schemes import parse_scheme


@pytest.mark.parametrize("scheme,kw,iters,expect", [
    ("W4A16", {}, 200, 1 / 200), ("W2A16", {"sym": False, "group_size": 32}, 1000, 2 / 1000),     # sign_round/config.py:107-136
    ("W2A16", {}, 999, 1 / 999), ("W4A16", {}, 1000, 1 / 1000), ("W3A16", {}, 1000, 2 / 1000)])
def test_lr_rule_and_linear_schedule(scheme, kw, iters, expect):
    sc = parse_scheme(scheme, kw)
    q = SignRoundQuantizer(sc, iters=iters)
    lr0 = q.compute_lr(sc.bits)
    assert lr0 == expect
    tab = lr_schedule_table(iters, lr0, lr0 * 2)
    lr, mm = torch.tensor(float(lr0)), torch.tensor(float(lr0 * 2))
    for it in range(iters):
        assert float(lr) == float(tab[it, 0]) and float(mm) == float(tab[it, 1]), it
        f = 1.0 + (0.0 - 1.0) / (iters * 1.0 + it * (0.0 - 1.0))         # LinearLR(1 -> 0), chainable form, fp32 tensor
        lr.mul_(f)
        mm.mul_(f)
    assert SignRoundQuantizer(sc, iters=iters, lr=0.01).compute_lr(sc.bits) == 0.01
