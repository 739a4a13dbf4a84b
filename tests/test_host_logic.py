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
