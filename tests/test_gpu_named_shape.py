"""Loop-level parity at the NAMED shapes of BASELINE.json (VERDICT r1 "next" 1b): one full-size decoder block of
the named architecture (random-init, HF default init), 16 samples x 2048 synthetic tokens, batches of 8, 200 sign-SGD iterations,
tuned (a) by this engine and (b) by the reference's algorithm -- oracle/signround.BlockTuner as torch eager with autograd on the
same GPU, the SAME batch sequence.  Bars (the tiny-block tests of test_gpu_engine.py use 2e-2 / +-25 %; at 16384 tokens per
iteration the sign-SGD trajectories stay together):
  * iteration-0 loss (V = 0: pure RTN forward)              rel <= 1e-3        (measured by the bench at this shape: 1.3e-5)
  * final block-output MSE vs the FP block                  within +-5 %       (measured: 0.04 % / 0.007 %), both below RTN;
                                                             NVFP4: +-8 % (measured -2.3 %: every 16 weights share a tunable scale
                                                             whose gradient the reference rounds through the scale tensor's dtype
                                                             while this engine keeps fp32)
  * sign(dV) agreement at iteration 0 (Llama W4A16 case)    >= 97 % of the elements above 5 % of the largest |dV|
Configs: Llama-3-8B W4A16 g128 (configs[1]), Llama-3-8B W2A16 asym g32 + enable_alg_ext (configs[2], 200 of its 1000
iterations), Qwen2-7B NVFP4 weight-only (configs[3])."""
import copy
import os
import random
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from auto_round_b200 import AutoRound  # noqa: E402
from auto_round_b200.quantizer import SignRoundQuantizer  # noqa: E402
from auto_round_b200.schemes import parse_scheme  # noqa: E402
from oracle import signround as S  # noqa: E402

DEV = torch.device("cuda", 0)
NS, SEQ, BS, ITERS = 16, 2048, 8, 200

CASES = {
    "llama3_8b_w4a16": ("llama3_8b", dict(scheme="W4A16"), S.LayerScheme(4, 128, True, "int"), False),
    "llama3_8b_w2asym_g32_algext": ("llama3_8b", dict(scheme="W2A16", group_size=32, sym=False), S.LayerScheme(2, 32, False, "int"), True),
    "qwen2_7b_nvfp4": ("qwen2_7b", dict(scheme="NVFP4", act_bits=16, act_data_type="float"), S.LayerScheme(4, 16, True, "nv_fp"), False),
}


def _setup(model_name):
    model = bench.build_model(model_name, 1, DEV)
    blk = model.model.layers[0].to(DEV)
    for p in blk.parameters():
        p.requires_grad_(False)
    vocab = bench.MODELS[model_name][2]["vocab_size"]
    tokens = torch.randint(0, vocab, (NS, SEQ), generator=torch.Generator().manual_seed(1))
    emb = model.model.embed_tokens.to(DEV)
    with torch.no_grad():
        xs = [emb(tokens[i:i + 1].to(DEV)).to(torch.bfloat16) for i in range(NS)]
        pos = torch.arange(SEQ, device=DEV).unsqueeze(0)
        cos, sin = model.model.rotary_emb.to(DEV)(xs[0], pos)
    others = {"position_embeddings": [(cos.to(torch.bfloat16), sin.to(torch.bfloat16))], "position_ids": [pos], "attention_mask": None}
    ids = []
    for i in range(NS):                                  # like the calibrator: the last position is excluded from the loss
        t = tokens[i:i + 1].clone()
        t[:, -1] = -100
        ids.append(t)
    masks = [(t != -100).to(torch.long).to(DEV) for t in ids]
    with torch.no_grad():
        refs = [S.block_forward(blk, x, {"position_embeddings": others["position_embeddings"][0], "position_ids": pos}) for x in xs]
    rng = random.Random(0)
    order, batches = [], []
    for _ in range(ITERS):
        if len(order) < BS:
            order = list(range(NS))
            rng.shuffle(order)
        batches.append([order.pop() for _ in range(BS)])
    return blk, xs, others, refs, ids, masks, batches


def _mse(block, xs, others, refs, masks):
    tot, cnt = torch.zeros((), dtype=torch.float64, device=DEV), 0
    kw = {"position_embeddings": others["position_embeddings"][0], "position_ids": others["position_ids"][0]}
    with torch.no_grad():
        for x, r, m in zip(xs, refs, masks):
            y = S.block_forward(block, x, kw)
            d = (y.float() - r.float()) * m.reshape(1, -1, 1)
            tot += (d.double() ** 2).sum()
            cnt += int(m.sum()) * y.shape[-1]
    return float(tot) / cnt


@pytest.mark.parametrize("case", list(CASES))
def test_named_shape_block_matches_the_reference_algorithm(case):
    model_name, skw, osc, alg_ext = CASES[case]
    blk, xs, others, refs, ids, masks, batches = _setup(model_name)
    scheme = parse_scheme(skw["scheme"], {k: v for k, v in skw.items() if k != "scheme"})
    names = [n for n, m in blk.named_modules() if type(m) is torch.nn.Linear]
    nv = AutoRound._fuse_nv_global_scales(blk, names) if scheme.qdq_name == "nv_fp4" else None
    # the RTN start both tuners share
    rtn = copy.deepcopy(blk)
    S.unwrap_block(rtn, S.wrap_block(rtn, lambda n, m: osc, nv), {})
    mse_rtn = _mse(rtn, xs, others, refs, masks)
    del rtn
    # (a) this engine
    ours = copy.deepcopy(blk)
    q = SignRoundQuantizer(scheme, iters=ITERS, batch_size=BS, enable_alg_ext=alg_ext)
    q.quantize_block(ours, xs, others, refs, None, None, input_ids=ids, nv_global_scales=nv, sampler=S.ReplaySampler(batches))
    res = q.last_result
    assert res.used_cuda_graph
    mse_ours = _mse(ours, xs, others, refs, masks)
    # (b) the reference's algorithm (torch eager + autograd) on the same GPU
    oblk = copy.deepcopy(blk)
    tuner = S.BlockTuner(oblk, xs, others, refs, lambda n, m: osc, iters=ITERS, batch_size=BS, token_masks=masks,
                         sampler=S.ReplaySampler(batches), nv_global_scales=nv, alg_ext=alg_ext)
    for it in range(ITERS):
        tuner.step(it)
    ores = tuner.finish()
    mse_ref = _mse(oblk, xs, others, refs, masks)
    print(f"\\n[{case}] iter0 loss ours {res.losses[0]:.6e} ref {ores.losses[0]:.6e} | final MSE ours {mse_ours:.6e} ref {mse_ref:.6e} "
          f"rtn {mse_rtn:.6e} | best iter {res.best_iter} / {ores.best_iter}")
    assert res.losses[0] == pytest.approx(ores.losses[0], rel=1e-3)
    assert mse_ours == pytest.approx(mse_ref, rel=0.08 if scheme.qdq_name == "nv_fp4" else 0.05), (mse_ours, mse_ref)
    assert mse_ours < mse_rtn and mse_ref < mse_rtn


def test_named_shape_sign_agreement_iteration0():
    """sign(dV) of the fused update kernel vs the oracle's autograd on a full-size Llama-3-8B block, one batch of 8 x 2048."""
    blk, xs, others, refs, ids, masks, batches = _setup("llama3_8b")
    osc = S.LayerScheme(4, 128, True, "int")
    idx = batches[0]
    oblk = copy.deepcopy(blk)
    wrapped = S.wrap_block(oblk, lambda n, m: osc)
    x, sel = S.select_batch(xs, others, idx)
    pred = S.block_forward(oblk, x, sel)
    mask = torch.cat([masks[i] for i in idx], dim=0).unsqueeze(-1)
    loss = S.masked_mse(pred, torch.cat([refs[i] for i in idx], dim=0), mask)
    (loss * 1000).backward()
    ours = copy.deepcopy(blk)
    q = SignRoundQuantizer(parse_scheme("W4A16"), iters=1, batch_size=BS, lr=1.0)          # V' = -sign(dV)
    q.quantize_block(ours, xs, others, refs, None, None, input_ids=ids, sampler=S.ReplaySampler([idx]), keep_arena=True)
    arena = q.last_arena
    agree, total = 0.0, 0
    for name in q.last_result.quantized_layers:
        o, n, shape = arena.views[name]["value"]
        g = -arena.params[o:o + n].view(shape)
        ref = wrapped[name].value.grad.reshape(shape)
        big = ref.abs() > 0.05 * ref.abs().max()
        agree += float((torch.sign(g)[big] == torch.sign(ref)[big]).sum())
        total += int(big.sum())
    print(f"\\nsign(dV) agreement on {total} elements above the noise floor: {agree / total:.5f}")
    assert agree / total >= 0.97, agree / total
