"""north_star: "packed integer outputs are bit-exact to the reference's RTN/pack path on the same inputs".
`AutoRound(iters=0, disable_opt_rtn=True).quantize_and_save()` on the GPU must write the same tensors (names, dtypes,
shapes, every bit) and the same quantization_config as the unmodified reference did for the fixtures."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from auto_round_b200 import AutoRound  # noqa: E402

KW = {"w4a16_sym_g32": dict(scheme="W4A16", group_size=32), "w2a16_asym_g32": dict(scheme="W2A16", group_size=32, sym=False),
      "nvfp4": dict(scheme="NVFP4", act_bits=16, act_data_type="float"), "mxfp4": dict(scheme="MXFP4", act_bits=16)}


class _Tok:
    pad_token_id = None
    pad_token = None

    def save_pretrained(self, *a, **k):
        return None


@pytest.mark.parametrize("tag", list(KW))
def test_rtn_checkpoint_bit_exact(golden_dir, tag, tmp_path):
    from safetensors import safe_open
    from transformers import LlamaConfig, LlamaForCausalLM

    rec = torch.load(os.path.join(golden_dir, f"rtn_export_{tag}.pt"), weights_only=False)
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, rms_norm_eps=1e-5,
                      rope_theta=10000.0, tie_word_embeddings=False)
    model = LlamaForCausalLM(cfg).to(torch.bfloat16).eval()
    model.load_state_dict(rec["init_state"])
    ar = AutoRound(model, tokenizer=_Tok(), iters=0, disable_opt_rtn=True, nsamples=8, seqlen=16, batch_size=4,
                   dataset=None, device_map=0, seed=42, **KW[tag])
    out = str(tmp_path / "ckpt")
    ar.quantize_and_save(out, format="auto_round")
    got = {}
    with safe_open(os.path.join(out, "model.safetensors"), "pt") as f:
        for k in f.keys():
            t = f.get_tensor(k)
            got[k] = t.view(torch.uint8) if t.dtype == torch.float8_e4m3fn else t
    for k, ref in rec["tensors"].items():
        assert k in got, k
        assert got[k].dtype == ref.dtype and tuple(got[k].shape) == tuple(ref.shape), k
        assert torch.equal(got[k], ref), k
    extra = {k for k in got if ".layers." in k and "layernorm" not in k} - set(rec["tensors"])
    assert not extra, extra
    qc = json.load(open(os.path.join(out, "config.json")))["quantization_config"]
    assert qc == rec["quantization_config"]
