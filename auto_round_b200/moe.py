"""MoE expert un-fusing -- host-side mirror of auto_round/modeling/fused_moe/moe_experts_interface.py:173-260
(linear_loop_experts_forward) and replace_modules.py: HF >= 5 stores experts as fused 3-D parameters
(`gate_up_proj [E, 2I, H]`, `down_proj [E, H, I]`); the tuner needs one nn.Linear per expert projection so that
`wrapper_block` can wrap them (names `...experts.{e}.{gate,up,down}_proj`).

Data-dependent token counts per expert make this path ineligible for CUDA-graph replay (the loop falls back to the
eager launch sequence of the same kernels); expert-parallel grouped GEMMs are a "next" row.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _ExpertContainer(nn.Module):
    def __init__(self, gate: nn.Linear, up: nn.Linear, down: nn.Linear):
        super().__init__()
        self.gate_proj, self.up_proj, self.down_proj = gate, up, down


class LinearLoopExperts(nn.Module):
    """Per-expert nn.Linear containers "0".."E-1" + the reference's loop forward."""

    def __init__(self, fused: nn.Module):
        super().__init__()
        gu, dn = fused.gate_up_proj.data, fused.down_proj.data
        e, two_i, h = gu.shape
        inter = two_i // 2
        self.num_experts = e
        self.act_fn = fused.act_fn
        for i in range(e):
            gate = nn.Linear(h, inter, bias=False, device=gu.device, dtype=gu.dtype)
            up = nn.Linear(h, inter, bias=False, device=gu.device, dtype=gu.dtype)
            down = nn.Linear(inter, h, bias=False, device=gu.device, dtype=gu.dtype)
            gate.weight.data.copy_(gu[i, :inter])
            up.weight.data.copy_(gu[i, inter:])
            down.weight.data.copy_(dn[i])
            self.add_module(str(i), _ExpertContainer(gate, up, down))

    def forward(self, hidden_states, top_k_index, top_k_weights):
        shape3 = None
        if hidden_states.dim() == 3:
            shape3 = hidden_states.shape
            hidden_states = hidden_states.view(-1, shape3[-1])
            top_k_index = top_k_index.view(-1, top_k_index.size(-1))
            top_k_weights = top_k_weights.view(-1, top_k_weights.size(-1))
        ntok, k = hidden_states.size(0), top_k_index.size(-1)
        token_idx = torch.arange(ntok, device=hidden_states.device).unsqueeze(1).expand(-1, k).reshape(-1)
        weights = top_k_weights.reshape(-1).to(hidden_states.dtype)
        expert_ids = top_k_index.reshape(-1)
        selected = hidden_states[token_idx]
        out = torch.zeros_like(selected)
        for e in range(self.num_experts):
            idx = torch.nonzero(expert_ids == e, as_tuple=False).squeeze(-1)
            if idx.numel() == 0:
                continue
            x = selected.index_select(0, idx)
            ex = getattr(self, str(e))
            y = ex.down_proj(self.act_fn(ex.gate_proj(x)) * ex.up_proj(x))
            out.index_copy_(0, idx, y.to(out.dtype))
        out = out * weights.unsqueeze(-1)
        res = out.view(ntok, k, -1).sum(dim=1)      # deterministic (no index_add_ atomics)
        return res.view(shape3) if shape3 is not None else res


def is_fused_experts(m: nn.Module) -> bool:
    gu, dn = getattr(m, "gate_up_proj", None), getattr(m, "down_proj", None)
    return isinstance(gu, nn.Parameter) and isinstance(dn, nn.Parameter) and gu.dim() == 3 and dn.dim() == 3 \
        and hasattr(m, "act_fn") and gu.shape[0] == dn.shape[0]


def unfuse_experts(block: nn.Module) -> int:
    """Replace every fused-experts module of `block` in place; returns how many were replaced."""
    from .wrapper import set_module
    n = 0
    for name, m in list(block.named_modules()):
        if is_fused_experts(m):
            set_module(block, name, LinearLoopExperts(m))
            n += 1
    return n
