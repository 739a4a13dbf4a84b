"""Crash-resume state (host logic only, no GPU): manifest rules of auto_round/utils/resume.py:74-164 and
the per-block result snapshot this engine adds (the reference keeps finished blocks in its ShardWriter / offloader)."""
import json
import os

import pytest
import torch
import torch.nn as nn

from auto_round_b200 import resume as R
from auto_round_b200.export import QuantLinear
from auto_round_b200.schemes import parse_scheme

BLOCKS = ["model.layers.0", "model.layers.1", "model.layers.2"]


def _sig(**kw):
    base = dict(model_id="m", scheme_desc="W4A16|<no-layer-config>", dataset_desc="tokens:abc", nsamples=8, seqlen=16,
                block_names=BLOCKS)
    base.update(kw)
    return R.compute_run_signature(**base)


def _block():
    torch.manual_seed(0)
    b = nn.Module()
    b.attn = nn.Module()
    b.attn.q_proj = nn.Linear(32, 16, bias=False)
    b.attn.o_proj = nn.Linear(16, 32, bias=True)
    b.norm = nn.LayerNorm(32)
    return b


def _quantise_in_place(b):
    """Stand-in for what the tuner leaves behind: one packed holder, one qdq linear with attributes."""
    sc = parse_scheme("W4A16", {"group_size": 16})
    bufs = {"qweight": torch.randint(-2**31, 2**31 - 1, (4, 16), dtype=torch.int32), "qzeros": torch.zeros(2, 2, dtype=torch.int32),
            "scales": torch.rand(2, 16).half(), "g_idx": torch.arange(32, dtype=torch.int32) // 16}
    b.attn.q_proj = QuantLinear(32, 16, sc, bufs, None)
    lin = b.attn.o_proj
    lin.weight.data = torch.randn_like(lin.weight).bfloat16().float()
    lin.scale = torch.rand(32, 1).half()
    lin.zp = 8
    lin.weight_global_scale = torch.tensor([3.5])
    return sc


def test_signature_depends_on_every_part():
    s0 = _sig()
    assert s0 == _sig()
    for k, v in dict(model_id="n", scheme_desc="W2A16|x", dataset_desc="tokens:abd", nsamples=9, seqlen=17, block_names=BLOCKS[:2]).items():
        assert _sig(**{k: v}) != s0, k
    a = R.dataset_fingerprint([torch.arange(8).reshape(2, 4)])
    assert a == R.dataset_fingerprint([torch.arange(8).reshape(2, 4)]) != R.dataset_fingerprint([torch.arange(8).reshape(4, 2)])
    assert R.layer_config_fingerprint(None) == "<no-layer-config>"
    assert R.layer_config_fingerprint({"b": {"bits": 2, "x": [1]}, "a": {"bits": 4}}) == "a:bits=4;b:bits=2"


def test_mark_resume_and_prefix_rules(tmp_path):
    st = R.ResumeState(str(tmp_path), _sig(), BLOCKS)
    assert st.resume_index == 0 and st.load_q_input() is None and st.load_input_ids() is None
    with pytest.raises(AssertionError):
        st.mark_block_done(BLOCKS[1], {}, None, [torch.zeros(1)])            # out of order
    st.mark_block_done(BLOCKS[0], {"x": {"kind": "qdq", "weight": torch.ones(2, 2)}}, [torch.ones(1, 3)], [torch.zeros(1, 3)])
    st.mark_block_done(BLOCKS[1], {}, None, [torch.full((1, 3), 2.0)])       # enable_quanted_input off: q_input file removed
    again = R.ResumeState(str(tmp_path), _sig(), BLOCKS)
    assert again.completed_blocks == BLOCKS[:2] and again.resume_index == 2
    assert again.load_q_input() is None
    assert torch.equal(again.load_input_ids()[0], torch.full((1, 3), 2.0))
    assert torch.equal(again.load_block(BLOCKS[0])["x"]["weight"], torch.ones(2, 2))
    # a different run must not pick the state up
    assert R.ResumeState(str(tmp_path), _sig(nsamples=9), BLOCKS).resume_index == 0
    # a manifest that is not a prefix of the block order is ignored
    m = json.load(open(tmp_path / R.MANIFEST))
    m["completed_blocks"] = [BLOCKS[1]]
    json.dump(m, open(tmp_path / R.MANIFEST, "w"))
    assert R.ResumeState(str(tmp_path), _sig(), BLOCKS).resume_index == 0
    # a manifest that names a block whose result file is gone is ignored as well
    m["completed_blocks"] = BLOCKS[:2]
    json.dump(m, open(tmp_path / R.MANIFEST, "w"))
    os.remove(again._block_path(BLOCKS[1]))
    assert R.ResumeState(str(tmp_path), _sig(), BLOCKS).resume_index == 0
    # corrupt json: start fresh, do not raise
    open(tmp_path / R.MANIFEST, "w").write("{not json")
    assert R.ResumeState(str(tmp_path), _sig(), BLOCKS).resume_index == 0
    again.clear()
    assert not any(p.name.startswith(("resume_", "block_")) for p in tmp_path.iterdir())


def test_block_snapshot_roundtrip():
    b = _block()
    sc = _quantise_in_place(b)
    snap = R.snapshot_block(b)
    assert set(snap) == {"attn.q_proj", "attn.o_proj"} and snap["attn.q_proj"]["kind"] == "packed"
    fresh = _block()
    done = R.restore_block(fresh, snap, lambda n, m: sc)
    assert sorted(done) == ["attn.o_proj", "attn.q_proj"]
    q = fresh.attn.q_proj
    assert isinstance(q, QuantLinear) and q.in_features == 32 and q.bits == 4
    for k in ("qweight", "qzeros", "scales", "g_idx"):
        assert torch.equal(getattr(q, k), getattr(b.attn.q_proj, k)), k
    assert "g_idx" not in q.state_dict() and "qweight" in q.state_dict()      # g_idx stays non-persistent
    o = fresh.attn.o_proj
    assert torch.equal(o.weight.data, b.attn.o_proj.weight.data) and o.zp == 8
    assert torch.equal(o.scale, b.attn.o_proj.scale) and torch.equal(o.weight_global_scale, torch.tensor([3.5]))
    assert torch.equal(fresh.norm.weight, b.norm.weight)                     # untouched modules stay as loaded


def test_restore_block_with_fused_experts_after_unfusing():
    """snapshots are taken AFTER the experts were un-fused (per-expert layer names), so a freshly loaded block --
    whose experts are still HF's fused 3-D parameters -- must be un-fused before restore_block can find the layers (what
    AutoRound.quantize does on resume)."""
    from auto_round_b200.moe import GroupedExperts, unfuse_experts

    class Fused(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(1)
            self.gate_up_proj = nn.Parameter(torch.randn(2, 2 * 16, 8, generator=g).bfloat16())
            self.down_proj = nn.Parameter(torch.randn(2, 8, 16, generator=g).bfloat16())
            self.act_fn = nn.SiLU()

    def fresh():
        b = nn.Module()
        b.mlp = nn.Module()
        b.mlp.experts = Fused()
        return b

    done = fresh()
    assert unfuse_experts(done) == 1 and isinstance(done.mlp.experts, GroupedExperts)
    names = [n for n, m in done.named_modules() if type(m) is nn.Linear]
    assert "mlp.experts.1.down_proj" in names and len(names) == 6
    for n in names:                                         # stand-in for what the tuner leaves behind
        lin = done.get_submodule(n)
        lin.weight.data.mul_(0.5)
        lin.scale = torch.full((lin.weight.shape[0], 1), 0.25, dtype=torch.bfloat16)
        lin.zp = None
    snap = R.snapshot_block(done)
    assert set(snap) == set(names)
    again = fresh()
    with pytest.raises(AttributeError):
        R.restore_block(again, snap, lambda n, m: parse_scheme("MXFP4", {"act_bits": 16}))     # still fused: no such layers
    again = fresh()
    unfuse_experts(again)
    R.restore_block(again, snap, lambda n, m: parse_scheme("MXFP4", {"act_bits": 16}))
    for n in names:
        a, b = again.get_submodule(n), done.get_submodule(n)
        assert torch.equal(a.weight.data, b.weight.data) and torch.equal(a.scale, b.scale)
    # the restored weights still alias the stacked tensors the grouped GEMMs read
    st = again.mlp.experts.stack("down_proj")
    assert torch.equal(st[1], again.get_submodule("mlp.experts.1.down_proj").weight.data)
