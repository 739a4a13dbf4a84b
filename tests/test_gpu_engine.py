"""GPU parity of the tuning engine (quantize_block / AutoRound) against the reference fixtures and the CPU oracle.

The sign-SGD trajectory is chaotic (sign flips on tiny gradients; bf16 GEMM summation order differs between
cuBLAS-on-CPU/oneDNN and tcgen05), so loop-level parity is stated as:
  (ii)  iteration-0 loss per block (V=0: pure RTN forward):           |rel err| <= 2e-2 on tiny blocks
        (bf16 activations of a 64-wide block: one bf16 ulp of the output is already 4e-3 of the error signal)
  (iii) sign(dV) agreement with the oracle's autograd at iteration 0:  >= 97 % on elements above the noise floor
  (iv)  final per-block output MSE vs the FP block:                    within +-25 % of the oracle's on tiny blocks,
        and never worse than 1.05x the RTN (iteration-0) MSE
  (v)   packed tensors bit-exact given the same (W_qdq, scale, zp): test_gpu_kernels.py
"""
import os
import random

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from auto_round_b200 import AutoRound, ops  # noqa: E402
from auto_round_b200.quantizer import SignRoundQuantizer  # noqa: E402
from auto_round_b200.schemes import parse_scheme  # noqa: E402
from oracle import signround as S  # noqa: E402

DEV = torch.device("cuda", 0)

SCHEMES = {
    "w4a16_sym_g32": (dict(scheme="W4A16", group_size=32), S.LayerScheme(4, 32, True, "int")),
    "w2a16_asym_g32": (dict(scheme="W2A16", group_size=32, sym=False), S.LayerScheme(2, 32, False, "int")),
    "nvfp4": (dict(scheme="NVFP4", act_bits=16, act_data_type="float"), S.LayerScheme(4, 16, True, "nv_fp")),
    "mxfp4": (dict(scheme="MXFP4", act_bits=16), S.LayerScheme(4, 32, True, "mx_fp")),
}


def _load(golden_dir, tag):
    return torch.load(os.path.join(golden_dir, f"block_{tag}.pt"), weights_only=False)


def _tiny_cfg():
    from transformers import LlamaConfig
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, rms_norm_eps=1e-5,
                      rope_theta=10000.0, tie_word_embeddings=False)
    cfg._attn_implementation = "sdpa"
    return cfg


def _tiny_block(state, device="cpu"):
    from transformers.models.llama.modeling_llama import LlamaDecoderLayer
    blk = LlamaDecoderLayer(_tiny_cfg(), 0).to(torch.bfloat16).eval()
    blk.load_state_dict(state)
    return blk.to(device)


def _block_mse(block, inputs, others, fp_outputs, masks, device):
    """fp32 MSE of block(inputs) vs fp_outputs over valid tokens, computed with plain torch on `device`."""
    tot, cnt = 0.0, 0
    with torch.no_grad():
        for i in range(len(inputs)):
            x, sel = S.select_batch(inputs, others, [i])
            sel = {k: (v.to(device) if isinstance(v, torch.Tensor) else
                       tuple(t.to(device) for t in v) if isinstance(v, tuple) else v) for k, v in sel.items()}
            with torch.autocast(device_type=torch.device(device).type, dtype=torch.bfloat16):
                y = block(x.to(device), **sel)
            y = y[0] if isinstance(y, (tuple, list)) else y
            m = masks[i].reshape(1, -1, 1).to(device)
            d = ((y.float() - fp_outputs[i].to(device).float()) * m)
            tot += float((d ** 2).sum())
            cnt += int(m.sum()) * y.shape[-1]
    return tot / cnt


@pytest.mark.parametrize("tag", list(SCHEMES))
def test_quantize_block_vs_oracle(golden_dir, tag):
    rec = _load(golden_dir, tag)
    kw, osc = SCHEMES[tag]
    scheme = parse_scheme(kw["scheme"], {k: v for k, v in kw.items() if k != "scheme"})
    b = rec["blocks"][0]
    masks = [(ids != -100).to(torch.long) for ids in b["input_ids"]]
    iters = 40

    # oracle (CPU) on the same inputs with the same batch sequence
    random.seed(1234)
    oblk = _tiny_block(b["block_state"])
    ores = S.tune_block(oblk, b["inputs"], b["others"], b["fp_outputs"], lambda n, m: osc, iters=iters,
                        batch_size=rec["batch_size"], token_masks=masks, nv_global_scales=b["nv_gs"] or None)
    o_mse = _block_mse(oblk, b["inputs"], b["others"], b["fp_outputs"], masks, "cpu")

    blk = _tiny_block(b["block_state"], DEV)
    for p in blk.parameters():
        p.requires_grad_(False)
    q = SignRoundQuantizer(scheme, iters=iters, batch_size=rec["batch_size"])
    nv = {n: g.to(DEV).reshape(1) for n, g in b["nv_gs"].items()} if b["nv_gs"] else None
    q.quantize_block(blk, [t.to(DEV) for t in b["inputs"]], b["others"], [t.to(DEV) for t in b["fp_outputs"]], None, None,
                     input_ids=b["input_ids"], nv_global_scales=nv, sampler=S.ReplaySampler(ores.batches))
    res = q.last_result
    assert len(res.losses) == iters and res.batches == ores.batches
    # (ii) iteration-0 loss: identical parameters (V=0, scales=1) -> only GEMM/attention rounding differs
    assert res.losses[0] == pytest.approx(ores.losses[0], rel=2e-2)
    # reference fixture itself (first recorded loss of the unmodified reference, same first batch by construction?)
    # -> compared in test_autoround_model_level
    g_mse = _block_mse(blk, b["inputs"], b["others"], b["fp_outputs"], masks, DEV)
    rtn_like = res.losses[0]
    assert res.best_loss <= res.losses[0] + 1e-12
    assert g_mse == pytest.approx(o_mse, rel=0.25), (g_mse, o_mse)
    # quantised layers carry the reference's attributes
    for name, lay in b["layers"].items():
        mod = blk.get_submodule(name)
        assert type(mod) is torch.nn.Linear
        assert tuple(mod.scale.shape) == tuple(lay["scale"].shape)
        if isinstance(lay["zp"], torch.Tensor):
            assert tuple(mod.zp.shape) == tuple(lay["zp"].shape)
        else:
            assert mod.zp == lay["zp"]
    assert rtn_like > 0


def test_sign_agreement_iteration0(golden_dir):
    """(iii): pre-sign dV of the fused GEMM epilogue vs the oracle's autograd on the reference's block inputs."""
    rec = _load(golden_dir, "w4a16_sym_g32")
    b = rec["blocks"][0]
    osc = S.LayerScheme(4, 32, True, "int")
    masks = [(ids != -100).to(torch.long) for ids in b["input_ids"]]
    idx = [0, 1, 2, 3]
    oblk = _tiny_block(b["block_state"])
    wrapped = S.wrap_block(oblk, lambda n, m: osc)
    x, sel = S.select_batch(b["inputs"], b["others"], idx)
    pred = S.block_forward(oblk, x, sel)
    mask = torch.cat([masks[i] for i in idx], dim=0).unsqueeze(-1)
    loss = S.masked_mse(pred, torch.cat([b["fp_outputs"][i] for i in idx], dim=0), mask)
    (loss * 1000).backward()

    blk = _tiny_block(b["block_state"], DEV)
    for p in blk.parameters():
        p.requires_grad_(False)
    # one iteration with lr = 1: V' = 0 - 1 * sign(dV), so the arena holds -sign(dV) of the fused update kernel
    q = SignRoundQuantizer(parse_scheme("W4A16", {"group_size": 32}), iters=1, batch_size=4, lr=1.0)
    q.quantize_block(blk, [t.to(DEV) for t in b["inputs"]], b["others"], [t.to(DEV) for t in b["fp_outputs"]], None, None,
                     input_ids=b["input_ids"], sampler=S.ReplaySampler([idx]), keep_arena=True)
    arena, names = q.last_arena, q.last_result.quantized_layers
    agree_all, n_all = 0.0, 0
    for name in names:
        o, n, shape = arena.views[name]["value"]
        g = -arena.params[o:o + n].view(shape).float().cpu()
        ref = wrapped[name].value.grad.reshape(shape)
        big = ref.abs() > 0.05 * ref.abs().max()
        agree_all += float((torch.sign(g)[big] == torch.sign(ref)[big]).sum())
        n_all += int(big.sum())
    assert agree_all / n_all >= 0.97, agree_all / n_all


class _Tok:
    pad_token_id = None
    pad_token = None

    def save_pretrained(self, *a, **k):
        return None


@pytest.mark.parametrize("tag", ["w4a16_sym_g32", "nvfp4"])
def test_autoround_model_level(golden_dir, tag, tmp_path):
    """Public API end to end on the tiny Llama of the fixtures: cached block inputs equal the reference's, the
    first logged loss matches the reference's first loss, every block improves on RTN, checkpoint tensors have the
    reference's names/dtypes/shapes."""
    from transformers import LlamaForCausalLM
    rec = _load(golden_dir, tag)
    kw, _ = SCHEMES[tag]
    model = LlamaForCausalLM(_tiny_cfg()).to(torch.bfloat16).eval()
    model.load_state_dict(rec["init_state"])
    tokens = rec["tokens"]
    bs = rec["batch_size"]
    dataset = [tokens[i:i + bs] for i in range(0, tokens.shape[0], bs)]
    ar = AutoRound(model, tokenizer=_Tok(), iters=rec["iters"], nsamples=tokens.shape[0], seqlen=tokens.shape[1],
                   batch_size=bs, dataset=dataset, device_map=0, seed=42, reference_mask_cast=True, **kw)
    hidden, others, ids = ar.cache_block_inputs(model.model.layers[0])
    b0 = rec["blocks"][0]
    for h, r in zip(hidden, b0["inputs"]):
        assert torch.equal(h.cpu(), r)
    for a, r in zip(ids, b0["input_ids"]):
        assert torch.equal(a, r)
    out_dir = str(tmp_path / "q")
    model, folders = ar.quantize_and_save(out_dir, format="auto_round")
    # same seed -> same python-random sampler stream as the reference for block 0
    assert ar.quantizer.last_result is not None
    first = ar.block_results[0]
    n0 = sum(int((b0["input_ids"][i] != -100).sum()) for i in b0["batches"][0])
    assert first["losses"][0] * n0 == pytest.approx(b0["losses"][0], rel=2e-2)
    for r in ar.block_results:
        assert r["best_loss"] <= r["init_loss"] * (1 + 1e-6)      # history is fp32, the device state is double
    from safetensors import safe_open
    names = {}
    with safe_open(os.path.join(out_dir, "model.safetensors"), "pt") as f:
        for k in f.keys():
            names[k] = f.get_tensor(k)
    pre = "model.layers.0.self_attn.q_proj."
    if tag.startswith("w4"):
        assert names[pre + "qweight"].dtype == torch.int32 and tuple(names[pre + "qweight"].shape) == (64 * 4 // 32, 64)
        assert names[pre + "qzeros"].dtype == torch.int32 and int(names[pre + "qzeros"][0, 0]) == 0x77777777
        assert names[pre + "scales"].dtype == torch.float16 and tuple(names[pre + "scales"].shape) == (2, 64)
        assert pre + "g_idx" not in names      # rebuilt by the loader; the reference's checkpoints omit it too
    else:
        assert names[pre + "weight_packed"].dtype == torch.uint8 and tuple(names[pre + "weight_packed"].shape) == (64, 32)
        assert names[pre + "weight_scale"].dtype == torch.float8_e4m3fn
        assert names[pre + "weight_global_scale"].dtype == torch.float32
        # q/k/v share one global scale (min over the three)
        gq = names["model.layers.0.self_attn.q_proj.weight_global_scale"]
        gk = names["model.layers.0.self_attn.k_proj.weight_global_scale"]
        assert torch.equal(gq, gk)
    import json
    cfg = json.load(open(os.path.join(out_dir, "config.json")))["quantization_config"]
    assert cfg["quant_method"] == "auto-round" and cfg["bits"] == 4
    assert cfg["packing_format"] == ("auto_round:auto_gptq" if tag.startswith("w4") else "auto_round:llm_compressor")


def test_eager_attention_block_keeps_cached_mask(golden_dir):
    """with attn_implementation='eager' HF applies NO mask when attention_mask is None, so the engine must keep
    the cached mask there.  The additive causal mask HF prepares for eager attention (0 / dtype-min) on an eager block must
    give the loss the oracle gets with the same mask through sdpa -- bidirectional attention is off by orders of magnitude."""
    from transformers.models.llama.modeling_llama import LlamaDecoderLayer

    rec = _load(golden_dir, "w4a16_sym_g32")
    b = rec["blocks"][0]
    osc = S.LayerScheme(4, 32, True, "int")
    masks = [(ids != -100).to(torch.long) for ids in b["input_ids"]]
    seq = b["inputs"][0].shape[1]
    causal = torch.ones(seq, seq, dtype=torch.bool).tril().reshape(1, 1, seq, seq)
    additive = torch.zeros(1, 1, seq, seq, dtype=torch.bfloat16).masked_fill(~causal, torch.finfo(torch.bfloat16).min)
    others = dict(b["others"])
    others["attention_mask"] = [additive.clone() for _ in b["inputs"]]
    iters = 6
    oblk = _tiny_block(b["block_state"])
    with torch.no_grad():
        refs = [S.block_forward(oblk, *S.select_batch(b["inputs"], others, [i])) for i in range(len(b["inputs"]))]
    batches = [[0, 1, 2, 3]] * iters
    ores = S.tune_block(oblk, b["inputs"], others, refs, lambda n, m: osc, iters=iters, batch_size=4, token_masks=masks,
                        sampler=S.ReplaySampler(batches))
    cfg = _tiny_cfg()
    cfg._attn_implementation = "eager"
    blk = LlamaDecoderLayer(cfg, 0).to(torch.bfloat16).eval()
    blk.load_state_dict(b["block_state"])
    blk = blk.to(DEV)
    for p in blk.parameters():
        p.requires_grad_(False)
    q = SignRoundQuantizer(parse_scheme("W4A16", {"group_size": 32}), iters=iters, batch_size=4)
    q.quantize_block(blk, [t.to(DEV) for t in b["inputs"]], others, [t.to(DEV) for t in refs], None, None,
                     input_ids=b["input_ids"], sampler=S.ReplaySampler(batches))
    assert q.last_result.losses[0] == pytest.approx(ores.losses[0], rel=3e-2)


def test_gradient_accumulation_vs_oracle(golden_dir):
    """gradient_accumulate_steps = 2 (quantizer.py:436-452): iterations of 2 micro-batches x 2 samples, sum-reduced loss, one
    sign-SGD step per iteration.  The oracle's accumulation is pinned bit-exact against the live reference
    (oracle/diff_fuzz.py --loops, case accum2_bs2)."""
    rec = _load(golden_dir, "w4a16_sym_g32")
    b = rec["blocks"][0]
    osc = S.LayerScheme(4, 32, True, "int")
    masks = [(ids != -100).to(torch.long) for ids in b["input_ids"]]
    iters = 30
    random.seed(77)
    oblk = _tiny_block(b["block_state"])
    ores = S.tune_block(oblk, b["inputs"], b["others"], b["fp_outputs"], lambda n, m: osc, iters=iters, batch_size=2,
                        token_masks=masks, gradient_accumulate_steps=2)
    o_mse = _block_mse(oblk, b["inputs"], b["others"], b["fp_outputs"], masks, "cpu")
    blk = _tiny_block(b["block_state"], DEV)
    for p in blk.parameters():
        p.requires_grad_(False)
    q = SignRoundQuantizer(parse_scheme("W4A16", {"group_size": 32}), iters=iters, batch_size=2, gradient_accumulate_steps=2)
    q.quantize_block(blk, [t.to(DEV) for t in b["inputs"]], b["others"], [t.to(DEV) for t in b["fp_outputs"]], None, None,
                     input_ids=b["input_ids"], sampler=S.ReplaySampler(ores.batches))
    res = q.last_result
    assert res.batches == ores.batches and len(res.batches[0]) == 4
    assert res.losses[0] == pytest.approx(ores.losses[0], rel=2e-2)
    assert res.best_loss <= res.losses[0] + 1e-12
    g_mse = _block_mse(blk, b["inputs"], b["others"], b["fp_outputs"], masks, DEV)
    assert g_mse == pytest.approx(o_mse, rel=0.25), (g_mse, o_mse)


@pytest.mark.parametrize("opts", [dict(enable_minmax_tuning=False), dict(not_use_best_mse=True), dict(lr=0.01, minmax_lr=0.02)])
def test_loop_options_vs_oracle(golden_dir, opts):
    """The loop options the oracle is pinned on against the live reference (oracle/diff_fuzz.py --loops) on the GPU engine."""
    rec = _load(golden_dir, "w4a16_sym_g32")
    b = rec["blocks"][0]
    osc = S.LayerScheme(4, 32, True, "int")
    masks = [(ids != -100).to(torch.long) for ids in b["input_ids"]]
    iters = 30
    random.seed(99)
    oblk = _tiny_block(b["block_state"])
    ores = S.tune_block(oblk, b["inputs"], b["others"], b["fp_outputs"], lambda n, m: osc, iters=iters, batch_size=4,
                        token_masks=masks, **opts)
    o_mse = _block_mse(oblk, b["inputs"], b["others"], b["fp_outputs"], masks, "cpu")
    blk = _tiny_block(b["block_state"], DEV)
    for p in blk.parameters():
        p.requires_grad_(False)
    q = SignRoundQuantizer(parse_scheme("W4A16", {"group_size": 32}), iters=iters, batch_size=4, **opts)
    q.quantize_block(blk, [t.to(DEV) for t in b["inputs"]], b["others"], [t.to(DEV) for t in b["fp_outputs"]], None, None,
                     input_ids=b["input_ids"], sampler=S.ReplaySampler(ores.batches))
    res = q.last_result
    assert res.losses[0] == pytest.approx(ores.losses[0], rel=2e-2)
    if opts.get("not_use_best_mse"):
        assert res.best_iter == iters - 1
    if opts.get("enable_minmax_tuning") is False:                       # scales must still be the RTN ones
        for name, lay in b["layers"].items():
            mod = blk.get_submodule(name)
            sc0 = ops.qdq_fwd(ops.make_spec("int_sym", 4, 32, *mod.weight.shape), torch.load(
                os.path.join(golden_dir, "block_w4a16_sym_g32.pt"), weights_only=False)["blocks"][0]["block_state"][name + ".weight"].to(DEV),
                want_wq=False, want_scale=True)[1]
            assert torch.equal(mod.scale.reshape(-1).cpu(), sc0.cpu().reshape(-1)), name
    g_mse = _block_mse(blk, b["inputs"], b["others"], b["fp_outputs"], masks, DEV)
    assert g_mse == pytest.approx(o_mse, rel=0.25), (g_mse, o_mse)


def test_per_channel_block_vs_oracle(golden_dir):
    """group_size = -1 (one group per output channel) through the whole loop: per-row update kernels, pack with g = K."""
    rec = _load(golden_dir, "w4a16_sym_g32")
    b = rec["blocks"][0]
    osc = S.LayerScheme(4, -1, True, "int")
    masks = [(ids != -100).to(torch.long) for ids in b["input_ids"]]
    iters = 30
    random.seed(5)
    oblk = _tiny_block(b["block_state"])
    ores = S.tune_block(oblk, b["inputs"], b["others"], b["fp_outputs"], lambda n, m: osc, iters=iters, batch_size=4, token_masks=masks)
    o_mse = _block_mse(oblk, b["inputs"], b["others"], b["fp_outputs"], masks, "cpu")
    blk = _tiny_block(b["block_state"], DEV)
    for p in blk.parameters():
        p.requires_grad_(False)
    scheme = parse_scheme("W4A16", {"group_size": -1})
    q = SignRoundQuantizer(scheme, iters=iters, batch_size=4)
    q.quantize_block(blk, [t.to(DEV) for t in b["inputs"]], b["others"], [t.to(DEV) for t in b["fp_outputs"]], None, None,
                     input_ids=b["input_ids"], sampler=S.ReplaySampler(ores.batches))
    res = q.last_result
    assert res.losses[0] == pytest.approx(ores.losses[0], rel=2e-2)
    g_mse = _block_mse(blk, b["inputs"], b["others"], b["fp_outputs"], masks, DEV)
    assert g_mse == pytest.approx(o_mse, rel=0.25), (g_mse, o_mse)
    from auto_round_b200 import export
    from oracle import pack as P
    lin = blk.self_attn.q_proj
    assert tuple(lin.scale.shape) == (lin.weight.shape[0], 1)
    wq, sc = lin.weight.data.clone(), lin.scale.clone()
    ql = export.pack_linear(lin, scheme, DEV)
    want = P.pack_int(wq.cpu(), sc.cpu().reshape(wq.shape[0], -1), 8, 4, wq.shape[1], zp_minus_one=True)
    import numpy as np
    assert np.array_equal(ql.qweight.numpy(), want["qweight"]) and np.array_equal(ql.scales.numpy(), want["scales"])
