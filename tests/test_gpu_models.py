"""Other architectures of BASELINE.json's configs through the public API on the GPU (tiny random-init versions):
  * OPT  (config 1: LayerNorm, ReLU, biases, learned positions)      W4A16 sym
  * Qwen2 (config 4: qkv bias)                                        NVFP4 weight-only
  * Mixtral (config 5: fused experts -> un-fused per-expert linears)  MXFP4 weight-only
For OPT and Qwen2 the tuned blocks are compared with the CPU oracle run on the SAME cached inputs and batch sequence
(iteration-0 loss within 2e-2, final block MSE within 25 %); Mixtral is checked functionally (every expert projection
quantised and packed under the reference's names, loss never worse than RTN)."""
import os
import random

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from auto_round_b200 import AutoRound  # noqa: E402
from auto_round_b200.quantizer import SignRoundQuantizer  # noqa: E402
from oracle import signround as S  # noqa: E402

DEV = torch.device("cuda", 0)


class _Tok:
    pad_token_id = None
    pad_token = None

    def save_pretrained(self, *a, **k):
        return None


def _tokens(vocab, n=8, s=32):
    return torch.randint(0, vocab, (n, s), generator=torch.Generator().manual_seed(1))


def _opt():
    from transformers import OPTConfig, OPTForCausalLM
    torch.manual_seed(0)
    cfg = OPTConfig(hidden_size=64, ffn_dim=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=128,
                    max_position_embeddings=64, word_embed_proj_dim=64)
    cfg._attn_implementation = "sdpa"
    return OPTForCausalLM(cfg).to(torch.bfloat16).eval()


def _qwen2():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(0)
    cfg = Qwen2Config(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, tie_word_embeddings=False)
    cfg._attn_implementation = "sdpa"
    return Qwen2ForCausalLM(cfg).to(torch.bfloat16).eval()


def _mixtral():
    from transformers import MixtralConfig, MixtralForCausalLM
    torch.manual_seed(0)
    cfg = MixtralConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                        num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, num_local_experts=4,
                        num_experts_per_tok=2, tie_word_embeddings=False)
    cfg._attn_implementation = "sdpa"
    m = MixtralForCausalLM(cfg).to(torch.bfloat16).eval()
    for p in m.parameters():
        if p.dim() >= 2:
            p.data.normal_(0, 0.05)
    return m


def _block_mse(block, xs, others, refs, masks, dev):
    tot, cnt = 0.0, 0
    with torch.no_grad():
        for i in range(len(xs)):
            x, sel = S.select_batch(xs, others, [i])
            sel = {k: (v.to(dev) if isinstance(v, torch.Tensor) else tuple(t.to(dev) for t in v) if isinstance(v, tuple) else v)
                   for k, v in sel.items()}
            with torch.autocast(device_type=torch.device(dev).type, dtype=torch.bfloat16):
                y = block(x.to(dev), **sel)
            y = y[0] if isinstance(y, (tuple, list)) else y
            m = masks[i].reshape(1, -1, 1).to(dev)
            tot += float((((y.float() - refs[i].to(dev).float()) * m) ** 2).sum())
            cnt += int(m.sum()) * y.shape[-1]
    return tot / cnt


@pytest.mark.parametrize("arch,kw,osc", [
    ("opt", dict(scheme="W4A16", group_size=32), S.LayerScheme(4, 32, True, "int")),
    ("qwen2", dict(scheme="NVFP4", act_bits=16, act_data_type="float"), S.LayerScheme(4, 16, True, "nv_fp")),
])
def test_first_block_vs_oracle(arch, kw, osc):
    import copy
    model = _opt() if arch == "opt" else _qwen2()
    tokens = _tokens(128)
    ar = AutoRound(model, tokenizer=_Tok(), iters=24, nsamples=8, seqlen=32, batch_size=4,
                   dataset=[tokens[:4], tokens[4:]], device_map=0, seed=42, reference_mask_cast=True, **kw)
    from auto_round_b200.autoround import find_blocks
    _, blocks = find_blocks(model)
    hidden, others, ids = ar.cache_block_inputs(blocks[0])
    masks = [(i != -100).to(torch.long) for i in ids]
    blk_gpu = blocks[0].to(DEV)
    for p in blk_gpu.parameters():
        p.requires_grad_(False)
    blk_cpu = copy.deepcopy(blk_gpu).to("cpu")
    others_cpu = {k: ([t.cpu() if isinstance(t, torch.Tensor) else tuple(u.cpu() for u in t) if isinstance(t, tuple) else t for t in v]
                      if isinstance(v, list) else (v.cpu() if isinstance(v, torch.Tensor) else
                                                   tuple(u.cpu() for u in v) if isinstance(v, tuple) else v))
                  for k, v in others.items()}
    xs = [h.cpu() for h in hidden]
    with torch.no_grad():
        refs = []
        for i in range(len(xs)):
            x, sel = S.select_batch(xs, others_cpu, [i])
            refs.append(S.block_forward(blk_cpu, x, sel))
    random.seed(7)
    ores = S.tune_block(blk_cpu, xs, others_cpu, refs, lambda n, m: osc, iters=24, batch_size=4, token_masks=masks)
    o_mse = _block_mse(blk_cpu, xs, others_cpu, refs, masks, "cpu")

    q = SignRoundQuantizer(ar.scheme, iters=24, batch_size=4)
    # per-layer NVFP4 global scales on both sides (the q/k/v fusion is covered by test_gpu_rtn_export / test_gpu_engine)
    q.quantize_block(blk_gpu, [h.to(DEV) for h in hidden], others, [r.to(DEV) for r in refs], None, None, input_ids=ids,
                     sampler=S.ReplaySampler(ores.batches))
    res = q.last_result
    assert res.losses[0] == pytest.approx(ores.losses[0], rel=2e-2)
    g_mse = _block_mse(blk_gpu, xs, others_cpu, refs, masks, DEV)
    assert g_mse == pytest.approx(o_mse, rel=0.25), (g_mse, o_mse)
    assert res.best_loss <= res.losses[0] * (1 + 1e-6)
    n_lin = sum(1 for m in blk_gpu.modules() if type(m) is torch.nn.Linear)
    assert len(res.quantized_layers) == n_lin and all(hasattr(m, "scale") for m in blk_gpu.modules() if type(m) is torch.nn.Linear)


def test_mixtral_experts_unfused_and_packed(tmp_path):
    from safetensors import safe_open
    model = _mixtral()
    tokens = _tokens(128)
    ar = AutoRound(model, tokenizer=_Tok(), iters=6, nsamples=8, seqlen=32, batch_size=4, dataset=[tokens[:4], tokens[4:]],
                   device_map=0, seed=42, scheme="MXFP4", act_bits=16)
    out = str(tmp_path / "mx")
    ar.quantize_and_save(out, format="auto_round")
    for r in ar.block_results:
        assert r["best_loss"] <= r["init_loss"] * (1 + 1e-6)
    names = set()
    with safe_open(os.path.join(out, "model.safetensors"), "pt") as f:
        names = set(f.keys())
    for e in range(4):
        for proj in ("gate_proj", "up_proj", "down_proj"):
            base = f"model.layers.0.block_sparse_moe.experts.{e}.{proj}"      # HF >= 5 names the MoE block so
            assert base + ".weight_packed" in names and base + ".weight_scale" in names, base
    assert "model.layers.0.self_attn.q_proj.weight_packed" in names
    assert not any(k.endswith("gate_up_proj") for k in names)
