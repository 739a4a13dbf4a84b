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
    if isinstance(max_scale, torch.Tensor):
        max_scale = max_scale.view(-1)
    if isinstance(init_scale, torch.Tensor):
        init_scale = init_scale.view(-1)
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


# --------------------------------------------------------------------------------------------
# optimized RTN (iters == 0, disable_opt_rtn unset) -- the scale search weighted by the imatrix
#   auto_round/data_type/int.py:89-122      opt_rtn_int_sym
#   auto_round/data_type/nvfp.py:331-408    search_nvfp4_scale, opt_rtn_nv_fp4
#   auto_round/data_type/mxfp.py:103-169    search_mx_scale (alg_ext init scale)
#   auto_round/data_type/gguf.py:437-484    _imatrix_handle_zero
# --------------------------------------------------------------------------------------------
def imatrix_handle_zero(qw: torch.Tensor, g: torch.Tensor, bits: int, group_size: int):
    """qw, g: [G,gs].  Groups whose importance holds zeros get a weight-derived or mean-filled importance."""
    if torch.min(qw) != 0:
        return qw
    qw = qw.reshape(-1, qw.shape[-1]).clone()
    zero_cnt = torch.sum(qw <= 1e-30, dim=-1)
    replace = zero_cnt > group_size // 2
    if torch.sum(replace) > 0:
        if bits <= 3:
            tmp = torch.abs(g)
        else:
            tmp = torch.abs(g) + torch.sqrt(torch.sum(torch.pow(g, 2), dim=-1, keepdim=True) / 32)
        tmp = tmp.to(qw.dtype)
        qw[replace, :] = tmp[replace, :]
    mean_replace = (zero_cnt > 0) & (zero_cnt <= group_size // 2)
    if torch.sum(mean_replace) > 0:
        tmp = (torch.sum(qw, dim=-1) / (qw.shape[1] - zero_cnt)).view(-1, 1).expand(-1, qw.shape[1])
        idx = qw == 0
        qw[idx] = tmp[idx]
    return qw.reshape(g.shape)


def imatrix_weights(imatrix, g: torch.Tensor, bits: int, group_size: int):
    """[K] importance -> [G,gs] per-element loss weights (int.py:107-115): pad K with 1e-5, broadcast over rows."""
    if imatrix is None:
        return 1.0
    im = imatrix.reshape(1, -1)
    k = im.shape[1]
    if group_size > 0 and k >= group_size and k % group_size:
        im = torch.nn.functional.pad(im, (0, math.ceil(k / group_size) * group_size - k), value=1e-5)
    im = im.reshape(1, -1)
    qw = im.expand(g.numel() // im.numel(), -1).reshape(g.shape)
    return imatrix_handle_zero(qw, g, bits, group_size)


def opt_rtn_int_sym(w, bits=4, group_size=128, imatrix=None, q_scale_thresh=1e-5):
    """Returns (qdq, scale [G,1] in w.dtype, maxq).  All range math stays in the weight dtype (bf16)."""
    g, shape, pad = to_groups(w, group_size)
    maxq = int(2.0 ** (bits - 1))
    qw = imatrix_weights(imatrix, g, bits, group_size)
    scale = search_scales_int(g, bits, qw=qw)
    scale = torch.where(scale < 0, torch.clamp(scale, max=-q_scale_thresh), torch.clamp(scale, min=q_scale_thresh))
    q = g.div(scale).round_().clamp_(-maxq, maxq - 1)
    out = q.mul_(scale).to(g.dtype)
    return from_groups(out, shape, pad), scale, maxq


def search_scales_nvfp4(g: torch.Tensor, qw=1.0):
    """[G,16] -> best per-group coefficient in {1.0, 0.50 .. 1.51} (nvfp.py:331-385).  The search derives its own
    per-tensor global scale from `g` (global_scale=None inside), NOT the layer's fused one."""
    x = g.float()
    q0, scale, _ = nv_fp4(x, group_size=16, v=0, max_scale=1.0)
    best = (((q0 - x) ** 2) * qw).sum(dim=-1)
    best_scale = torch.ones_like(scale)
    for sv in range(50, 152):
        c = sv / 100.0
        if c == 1.0:
            continue
        test = torch.full_like(scale, c)
        q, _, _ = nv_fp4(x, group_size=16, v=0, max_scale=test)
        loss = (((q - x) ** 2) * qw).sum(dim=-1)
        m = loss < best
        best[m] = loss[m]
        best_scale[m.view(-1, 1) if best_scale.dim() == 2 else m] = c
    return best_scale


def opt_rtn_nv_fp4(w, group_size=16, global_scale=None, max_scale=1.0, imatrix=None):
    g, shape, pad = to_groups(w, group_size)
    qw = imatrix_weights(imatrix, g, 4, group_size) if isinstance(imatrix, torch.Tensor) else 1.0
    init_scale = search_scales_nvfp4(g, qw)
    return nv_fp4(w, group_size=group_size, v=0, global_scale=global_scale, max_scale=max_scale, init_scale=init_scale) + (init_scale,)


def _qdq_mxfp_given_max(x, max_val, emax=2, max_norm=6.0):
    """mxfp.py:172-199 (qdq_mxfp): shared exponent from a GIVEN per-group max."""
    shared_exp = torch.where(max_val == 0, torch.ones_like(max_val), torch.log2(max_val))
    shared_exp = torch.floor(shared_exp) - emax
    shared_exp = torch.clamp(shared_exp, min=-127.0, max=127.0)
    t = x / (2 ** shared_exp)
    t = torch.clamp(t, min=-max_norm, max=max_norm)
    t = mx_quant_element(t)
    return t * (2 ** shared_exp)


def search_scales_mx(g: torch.Tensor, qw=None):
    """[G,32] -> per-group coefficient in {1, 0.5, 2} (mxfp.py:103-169)."""
    x = g.to(torch.float32)
    max_val, _ = torch.max(torch.abs(x), dim=-1, keepdim=True)
    scales = torch.ones_like(max_val)

    def loss_of(q):
        e = (q - x).pow(2)
        if isinstance(qw, torch.Tensor) or (qw is not None and qw != 1.0):
            e = e * qw
        return e.sum(dim=-1)

    best = loss_of(_qdq_mxfp_given_max(x, max_val))
    for c in (0.5, 2.0):
        loss = loss_of(_qdq_mxfp_given_max(x, max_val * c))
        m = loss < best
        scales[m] = c
        best[m] = loss[m]
    return scales


def opt_rtn_mx_fp4(w, group_size=32, imatrix=None):
    """mxfp.py:172-230 (quant_mx_opt_rtn): the {0.5,1,2} coefficient search, then the plain shared-exponent qdq."""
    g, _, _ = to_groups(w, group_size)
    qw = imatrix_weights(imatrix, g.to(torch.float32), 4, group_size) if isinstance(imatrix, torch.Tensor) else None
    coeff = search_scales_mx(g, qw)
    return mx_fp4(w, group_size=group_size, max_scale=coeff.view(-1)) + (coeff,)


# name -> function, mirrors the registry lookup the reference does in data_type/utils.py:105-176
QDQ = {"int_sym": int_sym, "int_asym": int_asym, "rtn_int_sym": rtn_int_sym, "mx_fp4": mx_fp4, "nv_fp4": nv_fp4,
       "opt_rtn_int_sym": opt_rtn_int_sym, "opt_rtn_nv_fp4": opt_rtn_nv_fp4, "opt_rtn_mx_fp4": opt_rtn_mx_fp4}
