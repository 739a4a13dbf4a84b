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
