"""`quant_lm_head=True` on the GPU: quantizer.quantize_layer against the oracle's tune_layer on the inputs
the reference fed its quantize_layer_outside_block (tests/golden/lm_head_w4a16_sym_g32.pt).

The oracle side is pinned bit-exact on the CPU
(tests/test_oracle_golden.py::test_tune_layer_lm_head_matches_reference_bit_exact).  Bars as in tests/test_gpu_engine.py: first-iteration loss within 2e-2
of the oracle's (identical parameters, only GEMM rounding differs), final layer-output MSE within +-25 %."""
import os
import random

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from auto_round_b200.quantizer import SignRoundQuantizer  # noqa: E402
from auto_round_b200.schemes import parse_scheme  # noqa: E402
from oracle import signround as S  # noqa: E402

DEV = torch.device("cuda", 0)


def _out_mse(lin, fp_lin_w, q_in, fp_in, masks, device):
    tot, cnt = 0.0, 0
    with torch.no_grad():
        for x, xf, m in zip(q_in, fp_in, masks):
            y = torch.nn.functional.linear(x.to(device).to(torch.bfloat16), lin.weight.to(device))
            t = torch.nn.functional.linear(xf.to(device).to(torch.bfloat16), fp_lin_w.to(device))
            d = (y.float() - t.float()) * m.to(device).reshape(1, -1, 1)
            tot += float((d ** 2).sum())
            cnt += int(m.sum()) * y.shape[-1]
    return tot / cnt


def test_quantize_layer_vs_oracle(golden_dir):
    rec = torch.load(os.path.join(golden_dir, "lm_head_w4a16_sym_g32.pt"), weights_only=False)
    lay = rec["layers"][0]
    n, k = lay["weight"].shape
    masks = [(ids != -100).to(torch.long) for ids in lay["input_ids"]]
    iters = 40
    random.seed(99)
    olin = torch.nn.Linear(k, n, bias=False).to(torch.bfloat16)
    olin.weight.data.copy_(lay["weight"])
    ores = S.tune_layer(olin, lay["fp_inputs"], lay["q_inputs"], S.LayerScheme(4, 32, True, "int"), iters=iters,
                        batch_size=rec["batch_size"], token_masks=masks)
    o_mse = _out_mse(olin, lay["weight"], lay["q_inputs"], lay["fp_inputs"], masks, "cpu")

    lin = torch.nn.Linear(k, n, bias=False).to(torch.bfloat16).to(DEV)
    lin.weight.data.copy_(lay["weight"])
    lin.weight.requires_grad_(False)
    q = SignRoundQuantizer(parse_scheme("W4A16", {"group_size": 32}), iters=iters, batch_size=rec["batch_size"])
    q.quantize_layer(lin, [t.to(DEV) for t in lay["fp_inputs"]], [t.to(DEV) for t in lay["q_inputs"]], input_ids=lay["input_ids"],
                     sampler=S.ReplaySampler(ores.batches))
    res = q.last_result
    assert res.batches == ores.batches and len(res.losses) == iters
    assert res.losses[0] == pytest.approx(ores.losses[0], rel=2e-2)
    assert res.best_loss <= res.losses[0] * (1 + 1e-6)
    g_mse = _out_mse(lin, lay["weight"], lay["q_inputs"], lay["fp_inputs"], masks, DEV)
    assert g_mse == pytest.approx(o_mse, rel=0.25), (g_mse, o_mse)
    assert tuple(lin.scale.shape) == (n, k // 32) and lin.zp == 8
