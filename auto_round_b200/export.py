"""Low-bit pack + `auto_round` checkpoint writer -- host-side mirror of
  auto_round/export/formats/backends/autoround.py:134-170      (pack_layer dispatch by scheme)
  auto_round/export/export_to_autogptq/export.py:133-185        (int sym  -> qlinear_torch_zp, zp-1 layout)
  auto_round/export/export_to_autoround/export.py:143-239       (int asym -> qlinear_torch, plain layout)
  auto_round/export/export_to_autoround/export_to_nvfp_mx.py:60-134  (fp4 -> qlinear_fp)
The packing arithmetic runs in the CUDA kernels of csrc/ar_pack.cu; buffer names, dtypes and shapes are the
checkpoint wire format of the reference so that its loaders read our output unchanged.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .schemes import QuantizationScheme
from .wrapper import set_module

AUTOROUND_VERSION = "0.15.0"      # format version of the reference this writer mirrors


class QuantLinear(nn.Module):
    """Packed replacement of an nn.Linear (holder of the wire-format buffers; inference is out of scope)."""

    def __init__(self, in_features: int, out_features: int, scheme: QuantizationScheme, buffers: dict, bias=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.bits, self.group_size, self.sym, self.data_type = scheme.bits, scheme.group_size, scheme.sym, scheme.data_type
        for k, v in buffers.items():
            # g_idx is rebuilt by the loader (k // group_size); the reference's checkpoints do not carry it
            self.register_buffer(k, v, persistent=(k != "g_idx"))
        if bias is not None:
            self.register_buffer("bias", bias)
        else:
            self.bias = None

    def extra_repr(self):
        return f"in={self.in_features}, out={self.out_features}, bits={self.bits}, g={self.group_size}, {self.data_type}"


def packing_format_of(scheme: QuantizationScheme) -> str:
    """export_to_autoround/export.py:275-280, formats/backends/autoround.py:62-75."""
    if scheme.data_type == "int" and scheme.sym:
        return "auto_round:auto_gptq"
    if scheme.data_type in ("nv_fp", "mx_fp", "nv_fp4", "mx_fp4"):
        return "auto_round:llm_compressor"             # what v0.15.0 writes for FP4 (tests/golden/rtn_export_*.pt)
    return "auto_round"


@torch.no_grad()
def pack_linear(layer: nn.Linear, scheme: QuantizationScheme, device=None, out_device="cpu") -> QuantLinear:
    """Pack one tuned nn.Linear (weight = qdq weight, `.scale/.zp/.weight_global_scale` set by the tuner)."""
    if not hasattr(layer, "scale"):
        raise RuntimeError("pack_linear: layer has no `.scale` (was it quantised?)")
    dev = torch.device(device) if device is not None else (layer.weight.device if layer.weight.is_cuda else torch.device("cuda"))
    wq = layer.weight.data.to(dev).to(torch.bfloat16).contiguous()
    n, k = wq.shape
    name = scheme.qdq_name
    bias = None if layer.bias is None else layer.bias.detach().to(torch.float16).to(out_device)
    if name in ("int_sym", "int_asym"):
        if scheme.data_type == "int" and not scheme.sym and scheme.bits == 4:
            raise NotImplementedError("int4 asym uses the AWQ layout in the reference (autoround.py:68-75): out of scope")
        scale = layer.scale.to(dev).to(torch.float16).reshape(n, -1).contiguous()
        if name == "int_sym":
            qw, qz, st, gi = ops.pack_int(wq, scale, None, scheme.bits, scheme.group_size, zp_minus_one=True,
                                          zp_const=int(layer.zp))
            bufs = {"qweight": qw, "qzeros": qz, "scales": st, "g_idx": gi}
        else:
            zp = layer.zp.to(dev).to(torch.float32).reshape(n, -1).contiguous()
            qw, qz, st, _ = ops.pack_int(wq, scale, zp, scheme.bits, scheme.group_size, zp_minus_one=False)
            bufs = {"qweight": qw, "qzeros": qz, "scales": st}
    elif name == "nv_fp4":
        scale = layer.scale.to(dev).to(torch.float32).reshape(n, -1).contiguous()
        gs = layer.weight_global_scale.to(dev).to(torch.float32).reshape(1).contiguous()
        pk, sc = ops.pack_fp4_nv(wq, scale, gs)
        bufs = {"weight_packed": pk, "weight_scale": sc.view(torch.float8_e4m3fn), "weight_global_scale": gs.clone()}
    elif name == "mx_fp4":
        e = layer.scale.to(dev).to(torch.bfloat16).reshape(n, -1).contiguous()
        pk, sc = ops.pack_fp4_mx(wq, e)
        bufs = {"weight_packed": pk, "weight_scale": sc}
    else:
        raise NotImplementedError(name)
    bufs = {k_: v.to(out_device) for k_, v in bufs.items()}
    return QuantLinear(k, n, scheme, bufs, bias)


def pack_layer(name: str, model: nn.Module, scheme: QuantizationScheme, device=None, out_device="cpu"):
    """Replace `model.<name>` by its packed QuantLinear (immediate_pack, compressors/utils.py:534-554)."""
    layer = model.get_submodule(name)
    if type(layer) is not nn.Linear or not hasattr(layer, "scale"):
        return None
    ql = pack_linear(layer, scheme, device, out_device)
    set_module(model, name, ql)
    layer.weight = None                     # release the dense weight
    return ql


def extra_config_entry(scheme: QuantizationScheme) -> dict:
    """`extra_config[layer]` of a quantised layer outside the blocks (lm_head): every QuantizationScheme field of the
    resolved per-layer config (export_to_autoround/export.py:303-306).  For the weight-only schemes in scope the reference
    resolves the activation fields to their defaults (16 bit float, dynamic, symmetric, the weight's group size)."""
    return {"act_bits": 16, "act_data_type": "float", "act_dynamic": True, "act_group_size": scheme.group_size, "act_sym": True,
            "bits": scheme.bits, "data_type": scheme.data_type, "group_size": scheme.group_size, "rotation_config": None,
            "super_bits": None, "super_group_size": None, "sym": scheme.sym}


def build_quantization_config(scheme: QuantizationScheme, block_names, extra: Optional[dict] = None, iters=200,
                              nsamples=128, seqlen=2048, batch_size=8, tuning: Optional[dict] = None) -> dict:
    """Keys of export_to_autoround/export.py:286-336 after filter_quantization_config (export/utils.py:334-374): a
    hyper-parameter is written only when it differs from the reference's default (`iters` 200, `lr` / `minmax_lr` 1/iters,
    `enable_minmax_tuning` / `enable_quanted_input` True); nsamples / seqlen / batch_size are not part of the serialised
    config in v0.15.0 (pinned by tests/test_export_config.py against configs the reference wrote)."""
    cfg = {"bits": scheme.bits, "group_size": scheme.group_size, "sym": scheme.sym, "data_type": scheme.data_type}
    tuning = tuning or {}
    if iters == 0:
        cfg["enable_quanted_input"] = False            # RTN route of the reference; the default (True) is filtered out
    else:
        if iters != 200:
            cfg["iters"] = int(iters)
        default_lr = 1.0 / iters
        for key in ("lr", "minmax_lr"):
            if tuning.get(key) is not None and float(tuning[key]) != default_lr:
                cfg[key] = float(tuning[key])
        for key in ("enable_minmax_tuning", "enable_quanted_input"):
            if tuning.get(key) is False:
                cfg[key] = False
    cfg["static_attention_granularity"] = "tensor"
    cfg["static_kv_granularity"] = "tensor"
    cfg["autoround_version"] = AUTOROUND_VERSION
    cfg["block_name_to_quantize"] = block_names
    cfg["quant_method"] = "auto-round"
    cfg["packing_format"] = packing_format_of(scheme)
    if extra:
        cfg["extra_config"] = extra
    return cfg


def save_quantized(model: nn.Module, output_dir: str, quantization_config: dict, tokenizer=None):
    os.makedirs(output_dir, exist_ok=True)
    if hasattr(model, "config"):
        model.config.quantization_config = quantization_config
    if tokenizer is not None and hasattr(tokenizer, "save_pretrained"):
        try:
            tokenizer.save_pretrained(output_dir)
        except Exception:  # noqa: BLE001  (dummy tokenizers in tests)
            pass
    if hasattr(model, "save_pretrained"):
        model.save_pretrained(output_dir, safe_serialization=True)
    else:
        from safetensors.torch import save_file
        save_file({k: v.contiguous() for k, v in model.state_dict().items()}, os.path.join(output_dir, "model.safetensors"))
        with open(os.path.join(output_dir, "quantization_config.json"), "w") as f:
            json.dump(quantization_config, f, indent=2)
    return output_dir
