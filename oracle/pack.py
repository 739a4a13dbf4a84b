"""CPU oracle for the low-bit weight packers (integer/byte work, numpy).

TEST INFRASTRUCTURE ONLY (see oracle/qdq.py header).  Restates:
  * auto_round_extension/torch/qlinear_torch_zp.py:93-150  pack_248_bits, "gptq_zp" layout (stores zp-1)
  * auto_round_extension/torch/qlinear_torch.py:110-168    pack_248_bits, plain layout
  * auto_round_extension/torch/qlinear_torch.py:170-281    pack_3bits (32 values in 3 words)
  * auto_round/export/export_to_autoround/qlinear_fp.py:141-193, :235-265   FP4 pack (NV / MX)
Pinned by tests/golden/pack_*.pt (reference outputs) and by the literal nibble known-answers of the
reference's own tests (test/unit/test_cpu/export/test_qlinear_fp_helpers.py:174-221).
"""
from __future__ import annotations

import numpy as np
import torch

from .qdq import cast_to_fp4, recip0, to_groups, from_groups

E2M1_LUT = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]  # qlinear_fp.py:49-58


def int_codes(w_qdq: torch.Tensor, scale: torch.Tensor, zp, group_size: int) -> np.ndarray:
    """round(W/s + zp) as int32 [N,K]  (qlinear_torch.py:126-134).  fp math in torch to keep its
    dtype promotion (bf16/fp16 -> fp32 division); everything after this is integer."""
    n, k = w_qdq.shape
    g = k if group_size == -1 else group_size
    s_rep = scale.repeat_interleave(g, 1)[:, :k]
    if isinstance(zp, torch.Tensor):
        z_rep = zp.repeat_interleave(g, 1)[:, :k]
        codes = torch.round(w_qdq / s_rep + z_rep)
    else:
        codes = torch.round(w_qdq / s_rep + zp)
    return codes.to(torch.int32).numpy()


def _pack_rows_pow2(vals: np.ndarray, bits: int) -> np.ndarray:
    """vals int32 [R, C] -> [R, C*bits/32]: 32/bits consecutive columns per word, LSB first."""
    per = 32 // bits
    r, c = vals.shape
    v = vals.astype(np.uint32).reshape(r, c // per, per)
    sh = (np.arange(per, dtype=np.uint32) * bits)[None, None, :]
    # the reference sums shifted int32 lanes; with in-range codes sum == or, wraparound included
    return (v << sh).sum(axis=-1, dtype=np.uint32).astype(np.uint32).view(np.int32)


def _pack_rows_3bit(vals: np.ndarray) -> np.ndarray:
    """[R, C] -> [R, C*3/32]; per 32 inputs, 3 words (qlinear_torch.py:197-227)."""
    r, c = vals.shape
    v = vals.astype(np.uint32).reshape(r, c // 32, 32)
    out = np.zeros((r, c // 32, 3), dtype=np.uint32)
    for j in range(10):
        out[:, :, 0] |= v[:, :, j] << np.uint32(3 * j)
    out[:, :, 0] |= v[:, :, 10] << np.uint32(30)
    out[:, :, 1] |= (v[:, :, 10] >> np.uint32(2)) & np.uint32(1)
    for j in range(10):
        out[:, :, 1] |= v[:, :, 11 + j] << np.uint32(3 * j + 1)
    out[:, :, 1] |= v[:, :, 21] << np.uint32(31)
    out[:, :, 2] |= (v[:, :, 21] >> np.uint32(1)) & np.uint32(3)
    for j in range(10):
        out[:, :, 2] |= v[:, :, 22 + j] << np.uint32(3 * j + 2)
    return out.reshape(r, -1).view(np.int32)


def pack_int(w_qdq: torch.Tensor, scale: torch.Tensor, zp, bits: int, group_size: int, zp_minus_one: bool):
    """Returns dict(qweight [K*bits/32, N] i32, qzeros [K/g, N*bits/32] i32, scales [K/g, N] f16, g_idx [K] i32).

    zp_minus_one=True is the AutoGPTQ-compatible layout the reference uses for int-sym
    (`auto_round:auto_gptq`); False is the plain `auto_round` layout used for int-asym, bits != 4.
    """
    n, k = w_qdq.shape
    g = k if group_size == -1 else group_size
    codes = int_codes(w_qdq, scale, zp, group_size)                 # [N, K]
    if bits == 3:
        qweight = _pack_rows_3bit(codes).T.copy()                   # along K, then transpose
    else:
        qweight = _pack_rows_pow2(codes, bits).T.copy()
    ngroups = scale.shape[1]
    if isinstance(zp, torch.Tensor):
        z = zp.t().contiguous().to(torch.int32).numpy().copy()     # [K/g, N]
        if zp_minus_one:
            z = z - 1
        z = z[:, : (n // 32 * bits) * 32 // bits] if bits != 3 else z
        qzeros = _pack_rows_3bit(z) if bits == 3 else _pack_rows_pow2(z, bits)
    else:
        zi = int(zp) - 1 if zp_minus_one else int(zp)
        z = np.full((ngroups, n), zi, dtype=np.int32)
        qzeros = _pack_rows_3bit(z) if bits == 3 else _pack_rows_pow2(z, bits)
    return {
        "qweight": qweight,
        "qzeros": qzeros,
        "scales": scale.t().contiguous().to(torch.float16).numpy(),
        "g_idx": (np.arange(k) // g).astype(np.int32),
    }


def unpack_int(qweight: np.ndarray, bits: int) -> np.ndarray:
    """Inverse of the power-of-two qweight packing: [K*bits/32, N] -> codes [N, K] (round-trip checks)."""
    per = 32 // bits
    w = qweight.view(np.uint32).T                                   # [N, K/per]
    sh = (np.arange(per, dtype=np.uint32) * bits)[None, None, :]
    return ((w[:, :, None] >> sh) & np.uint32((1 << bits) - 1)).reshape(w.shape[0], -1).astype(np.int32)


def fp4_nibbles(x: torch.Tensor) -> np.ndarray:
    """Nearest-LUT index | sign<<3 per element (qlinear_fp.py:245-257: argmin keeps the FIRST minimum,
    torch.signbit keeps -0.0)."""
    lut = torch.tensor(E2M1_LUT, dtype=x.dtype)
    idx = torch.argmin(torch.abs(torch.abs(x).unsqueeze(-1) - lut), dim=-1)
    return (idx + (torch.signbit(x).to(torch.long) << 3)).numpy().astype(np.uint8)


def _two_per_byte(nib: np.ndarray) -> np.ndarray:
    n, k = nib.shape
    flat = nib.reshape(-1, 2)
    return (flat[:, 0] | (flat[:, 1] << 4)).astype(np.uint8).reshape(n, k // 2)


def pack_nvfp4(w_qdq: torch.Tensor, scale: torch.Tensor, global_scale: torch.Tensor, group_size: int = 16):
    """weight_packed u8 [N,K/2], weight_scale e4m3 bytes [N,K/g], weight_global_scale f32 [1]."""
    g, shape, pad = to_groups(w_qdq, group_size)
    gs = global_scale.reshape([1]).to(torch.float32)
    x = g.to(torch.float32) * recip0(scale.reshape(g.shape[0], -1).to(torch.float32) * recip0(gs))
    x = cast_to_fp4(x.clamp_(-6.0, 6.0))
    x = from_groups(x, shape, pad)
    return {
        "weight_packed": _two_per_byte(fp4_nibbles(x)),
        "weight_scale": scale.to(torch.float8_e4m3fn).view(torch.uint8).numpy(),
        "weight_global_scale": gs.numpy(),
    }


def pack_mxfp4(w_qdq: torch.Tensor, shared_exp: torch.Tensor, group_size: int = 32):
    """weight_packed u8 [N,K/2] of W/2^e, weight_scale u8 [N,K/32] = clamp(e+127, 0, 255)."""
    g, shape, pad = to_groups(w_qdq, group_size)
    e = shared_exp.reshape(g.shape[0], -1)
    x = from_groups(g / (2 ** e), shape, pad)
    return {
        "weight_packed": _two_per_byte(fp4_nibbles(x)),
        "weight_scale": (shared_exp + 127).clamp(0, 255).to(torch.uint8).numpy(),
    }
