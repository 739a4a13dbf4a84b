"""AR_RESUME_DIR end to end on the GPU: a run killed after block 0 and restarted must finish with the same
checkpoint as an uninterrupted run (block results, chain values and the python/torch RNG state are restored; the kernels
are deterministic apart from the order of the double-precision loss atomics).
The host logic is covered by tests/test_resume.py on the CPU."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from auto_round_b200 import AutoRound  # noqa: E402


class _Tok:
    pad_token_id = None
    pad_token = None

    def save_pretrained(self, *a, **k):
        return None


class _Crash(Exception):
    pass


def _model(state):
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, rms_norm_eps=1e-5,
                      rope_theta=10000.0, tie_word_embeddings=False)
    m = LlamaForCausalLM(cfg).to(torch.bfloat16).eval()
    m.load_state_dict(state)
    return m


def _run(state, tokens, out_dir, crash_at=None):
    model = _model(state)
    ar = AutoRound(model, tokenizer=_Tok(), scheme="W4A16", group_size=32, iters=12, nsamples=8, seqlen=16, batch_size=4,
                   dataset=[tokens[:4], tokens[4:]], device_map=0, seed=42)
    if crash_at is not None:
        def hook(bi, phase):
            if bi == crash_at and phase == "h2d0":
                raise _Crash()
        ar.block_hook = hook
    ar.quantize_and_save(out_dir, format="auto_round")
    return ar


def _tensors(out_dir):
    from safetensors import safe_open

    got = {}
    with safe_open(os.path.join(out_dir, "model.safetensors"), "pt") as f:
        for k in f.keys():
            got[k] = f.get_tensor(k)
    return got


def test_resume_after_crash_equals_uninterrupted_run(golden_dir, tmp_path, monkeypatch):
    rec = torch.load(os.path.join(golden_dir, "rtn_export_w4a16_sym_g32.pt"), weights_only=False)
    state, tokens = rec["init_state"], rec["tokens"]
    monkeypatch.delenv("AR_RESUME_DIR", raising=False)
    _run(state, tokens, str(tmp_path / "plain"))
    want = _tensors(str(tmp_path / "plain"))

    rdir = tmp_path / "resume"
    monkeypatch.setenv("AR_RESUME_DIR", str(rdir))
    with pytest.raises(_Crash):
        _run(state, tokens, str(tmp_path / "crashed"), crash_at=1)
    assert (rdir / "group_0" / "resume_manifest.json").exists()
    ar = _run(state, tokens, str(tmp_path / "resumed"))
    assert ar.block_results[0].get("resumed") is True and "losses" in ar.block_results[1]
    got = _tensors(str(tmp_path / "resumed"))
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    assert not (rdir / "group_0" / "resume_manifest.json").exists()          # a finished run clears its state
