"""Host logic of the engine that needs no GPU: the attention-mask shortcut and the data-parallel shard arithmetic."""
import torch

from auto_round_b200.quantizer import DataParallel, _causal_equivalent


def _bool_causal(s, last_key_masked):
    m = torch.ones(s, s, dtype=torch.bool).tril()
    if last_key_masked:
        m[:, -1] = False                                   # the reference masks the last key of tensor datasets (calibration/llm.py:375-398)
    return m.reshape(1, 1, s, s)


def test_causal_equivalent_only_when_loss_rows_agree():
    s = 8
    valid = torch.ones(s, dtype=torch.uint8)
    last_invalid = valid.clone()
    last_invalid[-1] = 0
    # plain causal mask: always replaceable by is_causal
    assert _causal_equivalent([_bool_causal(s, False)], [valid])
    # last key masked: differs from causal only in the LAST row -> replaceable iff that row is excluded from the loss
    assert _causal_equivalent([_bool_causal(s, True)], [last_invalid])
    assert not _causal_equivalent([_bool_causal(s, True)], [valid])
    # a mask that differs on an earlier row is never replaceable
    m = _bool_causal(s, False).clone()
    m[0, 0, 3, 1] = False
    assert not _causal_equivalent([m], [last_invalid])
    # additive float masks (0 = visible) follow the same rule; no token mask -> keep the mask
    add = torch.zeros(1, 1, s, s).masked_fill(~_bool_causal(s, False), float("-inf"))
    assert _causal_equivalent([add], [valid])
    assert not _causal_equivalent([_bool_causal(s, False)], None)
    # the reference's bf16 cast of a boolean mask (+1 / 0 "bias") is NOT causal-equivalent: it must be applied as is
    assert not _causal_equivalent([_bool_causal(s, True).to(torch.bfloat16)], [last_invalid])


def test_dp_shard_partitions_every_batch():
    batch = [5, 2, 7, 1, 0, 6, 3, 4]
    for world in (1, 2, 4, 8):
        parts = [DataParallel(r, world).shard(batch) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(batch) and all(len(p) == 8 // world for p in parts)


def test_importance_weights_zero_repair_matches_oracle(golden_dir):
    """wrapper.importance_weights (device-agnostic torch) == the oracle's restatement of _imatrix_handle_zero
    (data_type/gguf.py:437-484), which reproduces the reference's `int_sym_w4g32_imzero` fixture bit-for-bit."""
    import os

    from auto_round_b200.wrapper import importance_weights
    from oracle import qdq as Q

    r = torch.load(os.path.join(golden_dir, "opt_rtn.pt"), weights_only=False)["int_sym_w4g32_imzero"]
    w, im = r["w"], r["imatrix"]
    g, _, _ = Q.to_groups(w, 32)
    want = Q.imatrix_weights(im, g, 4, 32)
    got = importance_weights(im, w, 4, 32)
    assert got.shape == (w.shape[0], w.shape[1]) and torch.equal(got.reshape(want.shape), want)
    # no zero channel: the [K] vector passes through untouched (the kernel broadcasts it and pads K with 1e-5)
    im2 = im.clone()
    im2[im2 == 0] = 0.5
    assert torch.equal(importance_weights(im2, w, 4, 32), im2)
    assert importance_weights(None, w, 4, 32) is None
    # K not a multiple of the group size: repaired matrix is [N, Kpad]
    r2 = torch.load(os.path.join(golden_dir, "opt_rtn.pt"), weights_only=False)["int_sym_w4g128_pad"]
    im3 = r2["imatrix"].clone()
    im3[:3] = 0
    got3 = importance_weights(im3, r2["w"], 4, 128)
    g3, _, _ = Q.to_groups(r2["w"], 128)
    assert torch.equal(got3.reshape(-1, 128), Q.imatrix_weights(im3, g3, 4, 128))


def test_layer_config_keys_are_full_module_names():
    """reference-style keys ("model.layers.3.mlp.down_proj": {...}) must reach the layer of THAT block only."""
    import torch.nn as nn

    from auto_round_b200.quantizer import SignRoundQuantizer
    from auto_round_b200.schemes import parse_scheme

    q = SignRoundQuantizer(parse_scheme("W4A16"), iters=1,
                           layer_config={"model.layers.3.mlp.down_proj": {"bits": 8}, "self_attn.o_proj": {"group_size": 32},
                                         "model.layers.1.mlp.up_proj": {"bits": 16}})
    lin = nn.Linear(8, 8)
    q.block_prefix = "model.layers.3"
    assert q.scheme_for("mlp.down_proj", lin).bits == 8
    assert q.scheme_for("mlp.up_proj", lin).bits == 4
    assert q.scheme_for("self_attn.o_proj", lin).group_size == 32       # bare block-relative key: every block
    q.block_prefix = "model.layers.1"
    assert q.scheme_for("mlp.down_proj", lin).bits == 4
    assert q.scheme_for("mlp.up_proj", lin).bits == 16                  # > 8 bits: wrapper_block leaves the layer alone
    assert q.scheme_for("self_attn.o_proj", lin).group_size == 32


def test_causal_mask_is_only_dropped_for_sdpa_or_flash_attention():
    """HF's eager attention applies NO mask when attention_mask is None, so the cached causal mask may only be
    dropped (-> is_causal fast path) for the sdpa / flash integrations."""
    import torch
    import torch.nn as nn

    from auto_round_b200.quantizer import SignRoundQuantizer, _causal_when_unmasked
    from auto_round_b200.schemes import parse_scheme

    class Cfg:
        def __init__(self, impl):
            self._attn_implementation = impl

    class Attn(nn.Module):
        def __init__(self, impl):
            super().__init__()
            self.config = Cfg(impl)

    class Block(nn.Module):
        def __init__(self, impl):
            super().__init__()
            self.self_attn = Attn(impl)

    assert _causal_when_unmasked(Block("sdpa")) and _causal_when_unmasked(Block("flash_attention_2"))
    assert not _causal_when_unmasked(Block("eager")) and not _causal_when_unmasked(nn.Linear(2, 2)) and not _causal_when_unmasked(None)
    s = 6
    causal = torch.ones(s, s, dtype=torch.bool).tril().reshape(1, 1, s, s)
    masks = [causal.clone(), causal.clone()]
    tok = [torch.ones(s, dtype=torch.uint8), torch.ones(s, dtype=torch.uint8)]
    q = SignRoundQuantizer(parse_scheme("W4A16"), iters=1)
    st, per = q._prepare_others({"attention_mask": masks}, tok, "cpu", Block("sdpa"))
    assert st.get("attention_mask", "absent") is None and "attention_mask" not in per
    st, per = q._prepare_others({"attention_mask": masks}, tok, "cpu", Block("eager"))
    assert "attention_mask" in per and per["attention_mask"].shape == (2, 1, s, s)
