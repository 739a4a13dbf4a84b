"""CPU oracle for the weight quant-dequant ("fake-quant") numerics of the AutoRound hot path.

TEST INFRASTRUCTURE ONLY -- a torch-CPU restatement of the reference's algorithm, used as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product path
(auto_round_b200/) never imports this package.

Each function cites the reference source it restates (paths relative to /root/reference).
The restatement keeps the reference's *operation order and dtypes* (fp16 scale cast, fp32 division,
half-to-even rounding, straight-through estimators built with .detach()) because bit-exactness of
`round(W/s + V)` depends on them.  Gradients come from torch autograd over these functions, exactly
as in the reference (there is no hand-written backward there).

Pinned by tests/golden/qdq_*.pt, generated from the unmodified reference by oracle/gen_golden.py.
"""
from __future__ import annotations

import math

import torch

E2M1_MAX = 6.0
E4M3_MAX = 448.0


# --------------------------------------------------------------------------------------------
# group reshape -- auto_round/data_type/utils.py:29-71 (reshape_pad_tensor_by_group_size),
#                  :74-102 (revert_tensor_by_pad)
# --------------------------------------------------------------------------------------------
def to_groups(w: torch.Tensor, group_size: int):
    """[N,K] -> [G,g] view; groups run along K inside a row.  Returns (groups, orig_shape, pad)."""
    shape = w.shape
    if w.dim() > 2:
        w = w.reshape(-1, shape[-1])
    if group_size == 0:
        return w.reshape(1, -1), shape, 0
    if group_size == -1 or w.shape[1] < group_size:
        return w, shape, 0
    k = w.shape[1]
    if k % group_size == 0:
        return w.reshape(-1, group_size), shape, 0
    pad = math.ceil(k / group_size) * group_size - k
    w = torch.nn.functional.pad(w, (0, pad), value=0.0)
    return w.reshape(-1, group_size), shape, pad


def from_groups(x: torch.Tensor, shape, pad: int):
    if pad == 0:
        return x.reshape(shape)
    rows = shape[0] if len(shape) <= 2 else int(math.prod(shape[:-1]))
    return x.reshape(rows, -1)[:, :-pad].reshape(shape)


# STE helpers -- auto_round/data_type/utils.py:314-347
def round_ste(x):
    return (x.round() - x).detach() + x


def floor_ste(x):
    return (x.floor() - x).detach() + x


def group_minmax(w: torch.Tensor, group_size: int):
    """weight_min / weight_max as WrapperLinear precomputes them -- auto_round/wrapper.py:154-167."""
    g, _, _ = to_groups(w, group_size)
    return torch.clamp(g.min(1)[0], max=0), torch.clamp(g.max(1)[0], min=0)


def _cast_ste(x, dtype):
    """value of x.to(dtype) held in x's own dtype, identity gradient.  NOT reference behaviour: used only by the
    `grad_fp32=True` variants below, which keep the forward values bit-identical but let autograd accumulate the
    scale gradient in fp32 instead of the scale tensor's fp16 (tests compare the CUDA kernels, which keep fp32,
    against this exact-arithmetic gradient; the default path stays bit-exact to the reference)."""
    return (x.to(dtype).to(x.dtype) - x).detach() + x


def _sym_scale_clip(scale, thr):
    # auto_round/data_type/int.py:231-232
    return torch.where(scale < 0, torch.clamp(scale, max=-thr), torch.clamp(scale, min=thr))


# --------------------------------------------------------------------------------------------
# int_sym -- auto_round/data_type/int.py:165-238 (quant_tensor_sym), "full range" symmetric
# --------------------------------------------------------------------------------------------
def int_sym(w, bits=4, group_size=128, v=0, min_scale=1.0, max_scale=1.0, wmin=None, wmax=None,
            scale_dtype=torch.float16, q_scale_thresh=1e-5, init_scale=None, grad_fp32=False):
    g, shape, pad = to_groups(w, group_size)
    maxq = int(2 ** (bits - 1))
    if init_scale is not None:  # int.py:201-216 (alg_ext optimized wrapper)
        ms = max_scale.unsqueeze(-1) if isinstance(max_scale, torch.Tensor) else max_scale
        scale = _sym_scale_clip((init_scale * ms).to(scale_dtype), q_scale_thresh)
    else:
        if wmin is None or wmax is None:
            wmin = torch.clamp(g.min(-1)[0], max=0)
            wmax = torch.clamp(g.max(-1)[0], min=0)
        lo = -(wmin * min_scale)
        hi = wmax * max_scale
        signed_max = (2 * (hi < lo).int() - 1) * torch.max(hi, lo)
        if grad_fp32:
            thr = float(torch.tensor(q_scale_thresh).to(scale_dtype))
            scale = _sym_scale_clip(_cast_ste(signed_max / maxq, scale_dtype), thr).unsqueeze(-1)
        else:
            scale = _sym_scale_clip((signed_max / maxq).to(scale_dtype), q_scale_thresh).unsqueeze(-1)
    q = torch.clamp(round_ste(g / scale + v), -maxq, maxq - 1)
    return from_groups((scale * q).to(g.dtype), shape, pad), scale, maxq


# rtn_int_sym -- auto_round/data_type/int.py:125-162 (same math, V=0, in-place rounding, no STE)
def rtn_int_sym(w, bits=4, group_size=128, min_scale=1.0, max_scale=1.0, scale_dtype=torch.float16,
                q_scale_thresh=1e-5):
    g, shape, pad = to_groups(w, group_size)
    maxq = int(2 ** (bits - 1))
    lo = -(torch.clamp(g.min(-1)[0], max=0) * min_scale)
    hi = torch.clamp(g.max(-1)[0], min=0) * max_scale
    signed_max = (2 * (hi < lo).int() - 1) * torch.max(hi, lo)
    scale = _sym_scale_clip((signed_max / maxq).to(scale_dtype), q_scale_thresh).unsqueeze(-1)
    q = g.div(scale).round_().clamp_(-maxq, maxq - 1)
    return from_groups(q.mul_(scale).to(g.dtype), shape, pad), scale, maxq


# --------------------------------------------------------------------------------------------
# int_asym -- auto_round/data_type/int.py:241-298 (quant_tensor_asym)
# --------------------------------------------------------------------------------------------
def int_asym(w, bits=4, group_size=128, v=0, min_scale=1.0, max_scale=1.0, wmin=None, wmax=None,
             scale_dtype=torch.float16, q_scale_thresh=1e-5, grad_fp32=False):
    g, shape, pad = to_groups(w, group_size)
    maxq = int(2 ** bits) - 1
    if wmin is None or wmax is None:
        wmin = torch.clamp(g.min(-1)[0], max=0)
        wmax = torch.clamp(g.max(-1)[0], min=0)
    if isinstance(min_scale, torch.Tensor):
        lo, hi = wmin * min_scale, wmax * max_scale
    else:
        lo, hi = wmin, wmax
    if grad_fp32:
        scale = torch.clamp(_cast_ste((hi - lo) / maxq, scale_dtype), min=float(torch.tensor(q_scale_thresh).to(scale_dtype)))
    else:
        scale = torch.clamp(((hi - lo) / maxq).to(scale_dtype), min=q_scale_thresh)
    zp = round_ste(-lo / scale).unsqueeze(-1)
    scale = scale.unsqueeze(-1)
    q = torch.clamp(round_ste(g / scale + v) + zp, 0, maxq)
    return from_groups((scale * (q - zp)).to(g.dtype), shape, pad), scale, zp


# --------------------------------------------------------------------------------------------
# E2M1 element rounding used by MXFP4 -- auto_round/data_type/mxfp.py:49-85 (quant_element),
# format row "mx_fp4": ebits 2, mbits 3, emax 2, max_norm 6.0 (mxfp.py:37)
# --------------------------------------------------------------------------------------------
def mx_quant_element(t, ebits=2, mbits=3, max_norm=6.0):
    pexp = floor_ste(torch.log2(torch.abs(t) + (t == 0).type(t.dtype)))
    pexp = pexp.clip(min=-(2.0 ** float(ebits - 1)) + 2)
    up = 2.0 ** float(mbits - 2)
    t = t / (2.0 ** pexp.float()) * up
    a = torch.abs(t)
    tie = ((a - 0.5) % 2 == torch.zeros_like(a)).type(t.dtype)  # half-to-even correction
    t = torch.sign(t) * (floor_ste(a + 0.5) - tie)
    t = t / up * (2.0 ** pexp.float())
    return torch.clamp(t, min=-max_norm, max=max_norm)


# mx_fp4 -- auto_round/data_type/mxfp.py:233-291 (quant_mx); min_scale is not used by the reference
def mx_fp4(w, group_size=32, v=0, max_scale=1.0, init_scale=1.0):
    ebits, mbits, emax, max_norm = 2, 3, 2, 6.0
    g, shape, pad = to_groups(w, group_size)
    init_scale = 1.0 if init_scale is None else init_scale
    dt = g.dtype
    g = g.to(torch.float32)
    amax, _ = torch.max(torch.abs(g), dim=-1, keepdim=True)
    if isinstance(max_scale, torch.Tensor):
        amax = amax * (init_scale * max_scale.unsqueeze(-1))
    else:
        amax = amax * (init_scale * max_scale)
    e = torch.where(amax == 0, torch.ones_like(amax), torch.log2(amax))
    e = (floor_ste(e) - emax).clamp(min=-127.0, max=127.0)
    s = torch.pow(2.0, e.float())
    t = torch.clamp(g / s + v, min=-max_norm, max=max_norm)
    out = mx_quant_element(t, ebits, mbits, max_norm) * s
    return from_groups(out, shape, pad).to(dt), e.to(dt), None


# --------------------------------------------------------------------------------------------
# NVFP4 -- auto_round/data_type/nvfp.py:26-39 (cast_to_fp4), :42-48 (get_reciprocal),
#          :56-64 (calculate_gparam), :67-80 (ref_nvfp4_quant), :83-98 (nv_fp4)
# --------------------------------------------------------------------------------------------
def cast_to_fp4(x):
    sgn = torch.sign(x)
    a = torch.abs(x)
    half_steps = round_ste(2.0 * a) / 2.0
    unit_steps = round_ste(a)
    two_steps = 2.0 * round_ste(a / 2.0)
    lt2 = a < 2.0
    lt4 = a < 4.0
    a = half_steps * lt2 + unit_steps * (~lt2) * lt4 + two_steps * (~lt2) * (~lt4)
    return a.clamp(-6, 6) * sgn


def recip0(x):
    """1/x with 1/0 := 0 (nvfp.py:42-48)."""
    if isinstance(x, torch.Tensor):
        return torch.where(x == 0, torch.zeros_like(x), 1.0 / x)
    return 0.0 if x == 0 else 1.0 / x


def nv_global_scale(w_or_amax):
    """448*6/amax(W) in fp32 (nvfp.py:56-64)."""
    if isinstance(w_or_amax, torch.Tensor):
        amax = w_or_amax.to(torch.float32).abs().max()
    else:
        amax = torch.tensor(float(w_or_amax), dtype=torch.float32).abs()
    return E4M3_MAX * E2M1_MAX * recip0(amax)


def e4m3_ste(x):
    # auto_round/data_type/utils.py:350-365
    return (x.to(torch.float8_e4m3fn).to(x.dtype) - x).detach() + x


def nv_fp4(w, group_size=16, v=0, global_scale=None, max_scale=1.0, init_scale=1.0):
    dt = w.dtype
    init_scale = 1.0 if init_scale is None else init_scale
    g, shape, pad = to_groups(w, group_size)
    if global_scale is None:
        global_scale = nv_global_scale(g)
    gs = global_scale.to(torch.float32)
    coeff = max_scale * init_scale
    if isinstance(coeff, torch.Tensor):
        coeff = coeff.view(-1, 1)
    vmax = torch.max(torch.abs(g), dim=-1, keepdim=True)[0].to(torch.float32) * coeff
    sc = torch.clamp(gs * (vmax * recip0(E2M1_MAX)), min=-E4M3_MAX, max=E4M3_MAX)
    sc = e4m3_ste(sc).to(torch.float32)
    inv = recip0(sc * recip0(gs))
    x = torch.clamp(g.to(torch.float32) * inv + v, -6.0, 6.0)
    out = cast_to_fp4(x) * recip0(inv)
    return from_groups(out, shape, pad).to(dt), sc, None


# --------------------------------------------------------------------------------------------
# alg_ext init-scale search (int) -- auto_round/data_type/int.py:24-86 (search_scales)
# utils.get_reciprocal -- auto_round/utils/common.py:903-922 (eps-thresholded, NOT the nvfp one)
# --------------------------------------------------------------------------------------------
def recip_eps(x):
    eps = 1e-5 if x.dtype == torch.float16 else 1e-30
    ok = x.abs() >= eps
    return torch.where(ok, 1.0 / torch.where(ok, x, torch.ones_like(x)), torch.zeros_like(x))


def search_scales_int(g: torch.Tensor, bits: int, qw=None, search_ratio=0.75):
    """Per-group scale grid search.  `g` is the [G,g] grouped weight (any float dtype)."""
    nmax = int(2.0 ** (bits - 1))
    pos = torch.abs(g).argmax(dim=-1, keepdim=True)
    gmax = torch.take_along_dim(g, pos, dim=-1)
    iscale = -nmax * recip_eps(gmax)
    scales = recip_eps(iscale)
    q = torch.round(iscale * g).clamp_(-nmax, nmax - 1)
    weight = 1.0 if qw is None else qw

    def err(s, q_):
        e = ((s * q_ - g).to(torch.float32)) ** 2
        if isinstance(weight, torch.Tensor):
            e = e * weight
        return e.sum(dim=-1)

    best = err(scales, q)
    if bits == 2:
        half, step = 90, 0.01
    else:
        span = nmax * search_ratio
        step = span / 200 * 2
        half = int(span / step)
    for i in range(-half, half + 1):
        if i == 0:
            continue
        isc = -(nmax - step * i) * recip_eps(gmax)
        q_ = torch.round(isc * g).clamp_(-nmax, nmax - 1)
        sc = recip_eps(isc)
        e = err(sc, q_)
        better = e < best
        if better.any():
            scales[better] = sc[better]
            best[better] = e[better]
    return scales


# name -> function, mirrors the registry lookup the reference does in data_type/utils.py:105-176
QDQ = {"int_sym": int_sym, "int_asym": int_asym, "rtn_int_sym": rtn_int_sym, "mx_fp4": mx_fp4, "nv_fp4": nv_fp4}
