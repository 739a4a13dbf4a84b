"""Pins the oracle's RTN + pack path against CHECKPOINTS written by the unmodified reference
(`AutoRound(..., iters=0, disable_opt_rtn=True).quantize_and_save(format="auto_round")` on a tiny Llama,
oracle/gen_golden.py rtn): every packed tensor of every quantised layer must be reproduced bit-for-bit from the
initial weights alone."""
import os

import numpy as np
import pytest
import torch

from oracle import pack as P
from oracle import qdq as Q

ATTN = ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"]
MLP = ["mlp.gate_proj", "mlp.up_proj"]


def fused_nv_scales(state, prefix):
    """q/k/v and gate/up share min(global_scale) -- auto_round/data_type/utils.py:433-530."""
    names = ATTN + ["self_attn.o_proj"] + MLP + ["mlp.down_proj"]
    g = {n: Q.nv_global_scale(state[prefix + n + ".weight"]) for n in names}
    for grp in (ATTN, MLP):
        m = min(g[n] for n in grp)
        for n in grp:
            g[n] = m
    return {prefix + n: v for n, v in g.items()}


def oracle_pack(tag, name, w, gs=None):
    if tag == "w4a16_sym_g32":
        wq, sc, zp = Q.rtn_int_sym(w.clone(), 4, 32)
        return P.pack_int(wq, sc.reshape(w.shape[0], -1), zp, 4, 32, True)
    if tag == "w2a16_asym_g32":
        wq, sc, zp = Q.int_asym(w, 2, 32)                      # python-float scales: range math stays bf16
        return P.pack_int(wq, sc.reshape(w.shape[0], -1), zp.reshape(w.shape[0], -1), 2, 32, False)
    if tag == "nvfp4":
        wq, sc, _ = Q.nv_fp4(w, 16, 0, gs)
        return P.pack_nvfp4(wq, sc.reshape(w.shape[0], -1), gs)
    wq, e, _ = Q.mx_fp4(w, 32, 0)
    return P.pack_mxfp4(wq, e.reshape(w.shape[0], -1))


@pytest.mark.parametrize("tag", ["w4a16_sym_g32", "w2a16_asym_g32", "nvfp4", "mxfp4"])
def test_oracle_reproduces_reference_checkpoint(golden_dir, tag):
    rec = torch.load(os.path.join(golden_dir, f"rtn_export_{tag}.pt"), weights_only=False)
    tensors, state = rec["tensors"], rec["init_state"]
    layers = sorted({k.rsplit(".", 1)[0] for k in tensors})
    assert len(layers) == 14
    gs_all = {}
    if tag == "nvfp4":
        for li in range(2):
            gs_all.update(fused_nv_scales(state, f"model.layers.{li}."))
    for name in layers:
        out = oracle_pack(tag, name, state[name + ".weight"], gs_all.get(name))
        keys = [k.rsplit(".", 1)[1] for k in tensors if k.rsplit(".", 1)[0] == name]
        for key in keys:
            assert np.array_equal(np.asarray(out[key]), tensors[f"{name}.{key}"].numpy()), (tag, name, key)
    cfg = rec["quantization_config"]
    assert cfg["quant_method"] == "auto-round"
    assert cfg["packing_format"] == {"w4a16_sym_g32": "auto_round:auto_gptq", "w2a16_asym_g32": "auto_round"}.get(tag, "auto_round:llm_compressor")


# ------------------------------------------------------------------------------------------------------------------
# optimized RTN (the reference's DEFAULT for iters=0): function-level fixtures + checkpoints with the importance matrix
# each layer saw (recorded, not altered, by oracle/gen_golden.py)
# ------------------------------------------------------------------------------------------------------------------
def run_opt(fn, w, kw, imatrix, gs):
    if fn == "opt_rtn_int_sym":
        return Q.opt_rtn_int_sym(w.clone(), bits=kw["bits"], group_size=kw["group_size"], imatrix=imatrix)
    if fn == "opt_rtn_nv_fp4":
        return Q.opt_rtn_nv_fp4(w.clone(), group_size=16, global_scale=gs, imatrix=imatrix)[:3]
    return Q.opt_rtn_mx_fp4(w.clone(), group_size=32, imatrix=imatrix)[:3]


def test_oracle_opt_rtn_functions(golden_dir):
    g = torch.load(os.path.join(golden_dir, "opt_rtn.pt"), weights_only=False)
    assert len(g) == 9
    for name, r in g.items():
        q, s, _ = run_opt(r["fn"], r["w"], r["kw"], r["imatrix"], r["global_scale"])
        assert torch.equal(q, r["qdq"]), name
        assert torch.equal(s.reshape(-1).float(), r["scale"].reshape(-1).float()), name
        assert s.dtype == r["scale"].dtype, name


def oracle_pack_opt(tag, w, imatrix, gs=None):
    if tag == "opt_w4a16_sym_g32":
        wq, sc, zp = Q.opt_rtn_int_sym(w.clone(), 4, 32, imatrix)
        return P.pack_int(wq, sc.reshape(w.shape[0], -1), zp, 4, 32, True)
    if tag == "opt_nvfp4":
        wq, sc, _, _ = Q.opt_rtn_nv_fp4(w, 16, gs, 1.0, imatrix)
        return P.pack_nvfp4(wq, sc.reshape(w.shape[0], -1), gs)
    wq, e, _, _ = Q.opt_rtn_mx_fp4(w, 32, imatrix)
    return P.pack_mxfp4(wq, e.reshape(w.shape[0], -1))


@pytest.mark.parametrize("tag", ["opt_w4a16_sym_g32", "opt_nvfp4", "opt_mxfp4"])
def test_oracle_reproduces_reference_opt_rtn_checkpoint(golden_dir, tag):
    rec = torch.load(os.path.join(golden_dir, f"rtn_export_{tag}.pt"), weights_only=False)
    tensors, state, imx = rec["tensors"], rec["init_state"], rec["imatrix"]
    layers = sorted({k.rsplit(".", 1)[0] for k in tensors})
    # MXFP4 runs zero-shot in the reference (no calibration, imatrix=None); int sym and NVFP4 collect an importance matrix
    assert len(layers) == 14 and set(imx) == (set() if tag == "opt_mxfp4" else set(layers))
    gs_all = {}
    if tag == "opt_nvfp4":
        for li in range(2):
            gs_all.update(fused_nv_scales(state, f"model.layers.{li}."))
    for name in layers:
        out = oracle_pack_opt(tag, state[name + ".weight"], imx.get(name), gs_all.get(name))
        for key in [k.rsplit(".", 1)[1] for k in tensors if k.rsplit(".", 1)[0] == name]:
            assert np.array_equal(np.asarray(out[key]), tensors[f"{name}.{key}"].numpy()), (tag, name, key)
    assert rec["quantization_config"]["enable_quanted_input"] is False
