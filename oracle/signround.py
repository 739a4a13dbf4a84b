"""CPU oracle for the block-wise SignRound tuning loop.

TEST INFRASTRUCTURE ONLY (see oracle/qdq.py header).  Restates, with torch-CPU autograd:
  * auto_round/wrapper.py:62-293, :517-565      WrapperLinear (tunable V / min_scale / max_scale, qdq + F.linear)
  * auto_round/wrapper.py:345-468               unwrapper (bake best params, attach scale/zp)
  * auto_round/algorithms/quantization/sign_round/quantizer.py:311-552   quantize_block (the 200-iter loop)
  * auto_round/algorithms/quantization/sign_round/quantizer.py:127-158   masked MSE loss
  * auto_round/algorithms/quantization/sign_round/sign_sgd.py:356-389    sign-SGD update
  * auto_round/compressors/utils.py:388-438     IndexSampler (python `random`, global state)
  * auto_round/compressors/utils.py:109-172     block_forward (autocast(bf16) around the HF layer)
Pinned by tests/golden/block_*.pt: loss curves / best iteration / final qdq weights / scales produced
by the unmodified reference on tiny random-init blocks (oracle/gen_golden.py), which this module must
reproduce bit-for-bit on CPU (same torch ops in the same order).
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import qdq as Q


@dataclass
class LayerScheme:
    bits: int = 4
    group_size: int = 128
    sym: bool = True
    data_type: str = "int"            # "int" | "mx_fp" | "nv_fp"
    scale_dtype: torch.dtype = torch.float16

    @property
    def qdq_name(self) -> str:
        if self.data_type == "int":
            return "int_sym" if self.sym else "int_asym"
        return {"mx_fp": "mx_fp4", "nv_fp": "nv_fp4"}[self.data_type]


class IndexSampler:
    """compressors/utils.py:388-438.  Consumes the GLOBAL python `random` state, like the reference."""

    def __init__(self, nsamples: int, batch_size: int):
        if batch_size <= 0 or batch_size > nsamples:
            raise ValueError("batch_size must be > 0 and <= nsamples")
        self.nsamples, self.batch_size, self.index = nsamples, batch_size, 0
        self.indices = list(range(nsamples))
        random.shuffle(self.indices)

    def next_batch(self):
        if self.index + self.batch_size > self.nsamples:
            random.shuffle(self.indices)
            self.index = 0
        out = self.indices[self.index: self.index + self.batch_size]
        self.index += self.batch_size
        return out


class ReplaySampler:
    """Feeds a recorded batch sequence (fixtures store what the reference's sampler produced)."""

    def __init__(self, batches):
        self.batches, self.i = [list(b) for b in batches], 0

    def next_batch(self):
        b = self.batches[self.i]
        self.i += 1
        return b


class TunableLinear(nn.Module):
    """WrapperLinear restated: same parameters, same per-forward in-place clamp, same qdq call."""

    def __init__(self, linear: nn.Linear, scheme: LayerScheme, global_scale=None, minmax_bound=(0.0, 1.0),
                 init_scale=None):
        super().__init__()
        self.linear, self.scheme, self.bound = linear, scheme, minmax_bound
        self.init_scale = init_scale          # alg_ext (SignRoundOptimizedWrapperLinear): searched per-group init scale
        w = linear.weight.data
        groups, _, _ = Q.to_groups(w, scheme.group_size)
        self.wmin = torch.clamp(groups.min(1)[0], max=0)
        self.wmax = torch.clamp(groups.max(1)[0], min=0)
        self.value = nn.Parameter(torch.zeros(groups.shape, dtype=torch.float32, device=w.device))
        nscale = groups.shape[0]
        self.min_scale = nn.Parameter(torch.ones(nscale, dtype=torch.float32, device=w.device))
        self.max_scale = nn.Parameter(torch.ones(nscale, dtype=torch.float32, device=w.device))
        self.q_scale_thresh = 1e-8 if scheme.scale_dtype == torch.float32 else 1e-5
        self.global_scale = None
        if scheme.data_type == "nv_fp":
            self.global_scale = global_scale if global_scale is not None else Q.nv_global_scale(w)
        self.params = {"value": self.value, "min_scale": self.min_scale, "max_scale": self.max_scale}

    def qdq(self, value, min_scale, max_scale):
        lo, hi = self.bound
        min_scale.data.clamp_(lo, hi)
        max_scale.data.clamp_(lo, hi)
        w, sc = self.linear.weight, self.scheme
        name = sc.qdq_name
        if name == "int_sym":
            out = Q.int_sym(w, sc.bits, sc.group_size, value, min_scale, max_scale, self.wmin, self.wmax,
                            sc.scale_dtype, self.q_scale_thresh, init_scale=self.init_scale)
        elif name == "int_asym":
            out = Q.int_asym(w, sc.bits, sc.group_size, value, min_scale, max_scale, self.wmin, self.wmax,
                             sc.scale_dtype, self.q_scale_thresh)
        elif name == "mx_fp4":
            out = Q.mx_fp4(w, sc.group_size, value, max_scale, init_scale=self.init_scale)
        else:
            out = Q.nv_fp4(w, sc.group_size, value, self.global_scale, max_scale,
                           init_scale=1.0 if self.init_scale is None else self.init_scale)
        wq, scale, zp = out
        return wq.to(w.dtype), scale, zp

    def forward(self, x):
        wq, _, _ = self.qdq(self.value, self.min_scale, self.max_scale)
        return F.linear(x, wq, self.linear.bias)

    @torch.no_grad()
    def bake(self, best):
        """unwrapper: qdq with the best params -> weight.data; attach scale / zp / weight_global_scale."""
        best = best or {}
        dev = self.linear.weight.device
        v = best.get("value", torch.tensor(0.0, device=dev))
        mn = best.get("min_scale", torch.tensor(1.0, device=dev))
        mx = best.get("max_scale", torch.tensor(1.0, device=dev))
        wq, scale, zp = self.qdq(v, mn, mx)
        lin = self.linear
        lin.weight.data.copy_(wq)
        n = wq.shape[0]
        lin.scale = scale.reshape(n, -1) if scale.numel() > 1 else scale.view(-1)
        if isinstance(zp, torch.Tensor):
            lin.zp = zp.reshape(n, -1) if zp.numel() > 1 else zp.view(-1)
        else:
            lin.zp = zp
        if self.global_scale is not None:
            lin.weight_global_scale = self.global_scale
        return lin


def search_init_scale(w, scheme: LayerScheme, imatrix=None, q_scale_thresh=1e-5):
    """search_optimized_init_scale (data_type/utils.py:223-254) on the grouped weight; the importance matrix is padded with
    1e-5 and broadcast over rows WITHOUT the zero repair of the RTN route (reshape_imatrix_for_weight, :269-282)."""
    g, _, _ = Q.to_groups(w, scheme.group_size)
    if imatrix is None:
        qw = torch.ones_like(g)
    else:
        im = imatrix.reshape(1, -1)
        k = im.shape[1]
        gs = scheme.group_size
        if gs > 0 and k >= gs and k % gs:
            im = F.pad(im, (0, (k + gs - 1) // gs * gs - k), value=1e-5)
        qw = im.reshape(1, -1).expand(g.numel() // im.numel(), -1).reshape(g.shape)
    name = scheme.qdq_name
    if name == "int_sym":
        s0 = Q.search_scales_int(g, scheme.bits, qw)
        return torch.where(s0 < 0, torch.clamp(s0, max=-q_scale_thresh), torch.clamp(s0, min=q_scale_thresh))
    if name == "mx_fp4":
        return Q.search_scales_mx(g, qw)
    if name == "nv_fp4":
        return Q.search_scales_nvfp4(g, qw)
    return None                                   # int asym keeps the plain wrapper (data_type/utils.py:197-201)


def outlier_suppressed_loss(pred, ref, mask):
    """SignRoundV2Quantizer._get_loss (sign_roundv2/quantizer.py:362-399): the numel/1000 largest |pred - ref| (taken on
    the bf16 difference, token mask NOT applied to the selection) are zeroed, then mean over ALL elements of the squared
    fp32 difference."""
    diff = torch.abs(pred - ref)
    flat = diff.view(-1)
    topk = max(1, int(flat.numel() / 1000))
    _, top = torch.topk(torch.abs(flat), topk)
    keep = torch.ones_like(flat, dtype=torch.bool)
    keep[top] = False
    keep = keep.view_as(diff)
    d = torch.abs(pred.to(torch.float32) - ref.to(torch.float32))
    if mask is not None:
        return torch.mean((d * mask * keep) ** 2)
    return torch.mean((d * keep) ** 2)


def wrap_block(block: nn.Module, scheme_of, nv_global_scales=None, alg_ext=False, imatrices=None):
    """wrapper_block (wrapper.py:774-828): every nn.Linear with bits <= 8 -> TunableLinear.  alg_ext: symmetric int / mx /
    nv layers get the optimized wrapper (searched init_scale, max_scale bound [0,2]; sign_roundv2/quantizer.py:101-125)."""
    wrapped = {}
    for name, mod in list(block.named_modules()):
        if type(mod) is nn.Linear:
            sc = scheme_of(name, mod)
            if sc is None or sc.bits > 8:
                continue
            gs = None if nv_global_scales is None else nv_global_scales.get(name)
            init, bound = None, (0.0, 1.0)
            if alg_ext and sc.qdq_name != "int_asym":
                thr = 1e-8 if sc.scale_dtype == torch.float32 else 1e-5
                init = search_init_scale(mod.weight.data, sc, (imatrices or {}).get(name), thr)
                bound = (0.0, 2.0)
            tl = TunableLinear(mod, sc, gs, bound, init)
            parent = block
            parts = name.split(".")
            for p in parts[:-1]:
                parent = getattr(parent, p)
            setattr(parent, parts[-1], tl)
            wrapped[name] = tl
    return wrapped


def unwrap_block(block: nn.Module, wrapped: dict, best: dict):
    for name, tl in wrapped.items():
        lin = tl.bake(best.get(name))
        parent = block
        parts = name.split(".")
        for p in parts[:-1]:
            parent = getattr(parent, p)
        setattr(parent, parts[-1], lin)


SHARED_KEYS = ("position_ids", "cache_position", "position_embeddings", "cu_seqlens")  # utils/common.py:676


def select_batch(inputs, others: dict, idx):
    """BlockForwardRunner._select_batch (algorithms/block_runner.py:368-422): per-sample lists are
    concatenated over `idx`; SHARED_KEYS entries are shared across samples (first cached copy)."""
    x = torch.cat([inputs[i] for i in idx], dim=0)
    sel = {}
    for key, val in others.items():
        if "positional_inputs" in key:
            continue
        if key in SHARED_KEYS:
            if isinstance(val, list) and len(val) >= 1:
                j = int(idx[0]) if (len(idx) == 1 and len(val) > 1) else 0
                sel[key] = val[j] if j < len(val) else val[0]
            else:
                sel[key] = val
        elif isinstance(val, list):
            vals = [val[i] for i in idx]
            sel[key] = vals[0] if len(vals) == 1 else torch.cat(vals, dim=0)
        elif isinstance(val, torch.Tensor):
            sel[key] = torch.index_select(val, 0, torch.tensor(list(idx), device=val.device))
        else:
            sel[key] = val
    return x, sel


def block_forward(block, hidden, others: dict, amp=True, amp_dtype=torch.bfloat16):
    kw = dict(others)
    if amp:
        with torch.autocast(device_type=hidden.device.type, dtype=amp_dtype):
            out = block(hidden, **kw)
    else:
        out = block(hidden, **kw)
    return out[0] if isinstance(out, (tuple, list)) else out


def masked_mse(pred, ref, mask, reduction="mean"):
    """quantizer.py:142-156 -- MSELoss('mean') over ALL elements; masked tokens contribute zeros.  reduction="sum" is what
    the loop switches to under gradient accumulation (quantizer.py:436-439)."""
    if mask is not None:
        return F.mse_loss((pred * mask).to(torch.float32), (ref * mask).to(torch.float32), reduction=reduction)
    return F.mse_loss(pred.to(torch.float32), ref.to(torch.float32), reduction=reduction)


@dataclass
class TuneResult:
    losses: list = field(default_factory=list)     # total_loss per iteration (loss.item()/num_elm)
    best_iter: int = 0
    best_loss: float = float("inf")
    best_params: dict = field(default_factory=dict)
    batches: list = field(default_factory=list)


class BlockTuner:
    """quantize_block as a stepper: construction = wrapper_block + optimizer / scheduler / sampler set-up
    (quantizer.py:311-436), `step(it)` = one iteration of the loop (quantizer.py:455-552), `finish()` = apply the best
    parameters and unwrap.  `tune_block` below is the plain loop over it; the bench times `step` on the host cores (the
    reference's CPU path) and on the GPU (the reference's eager ATen + cuBLAS path) without the set-up cost."""

    def __init__(self, block, inputs, others, fp_outputs, scheme_of, iters=200, batch_size=8, lr=None, minmax_lr=None,
                 token_masks=None, enable_minmax_tuning=True, nv_global_scales=None, amp=True, not_use_best_mse=False,
                 sampler=None, alg_ext=False, imatrices=None, outlier_loss=None, gradient_accumulate_steps=1):
        self.block, self.inputs, self.others, self.fp_outputs = block, inputs, others, fp_outputs
        # gradient accumulation (quantizer.py:436-452, 480-500): an iteration = batch_size * steps samples run as micro-batches
        # of batch_size with MSELoss(reduction="sum"); without a token mask num_elm is fixed before the loop
        self.accum, self.micro_bs, self.micro_losses = int(gradient_accumulate_steps), batch_size, []
        self.iters, self.token_masks, self.amp, self.alg_ext = iters, token_masks, amp, alg_ext
        self.not_use_best_mse = not_use_best_mse
        nsamples = len(inputs)
        self.wrapped = wrapped = wrap_block(block, scheme_of, nv_global_scales, alg_ext, imatrices)
        if outlier_loss is None:   # sign_roundv2/quantizer.py:334-352: symmetric schemes with bits < 4 (act quant is out of scope)
            outlier_loss = alg_ext and any(tl.init_scale is not None and tl.scheme.bits < 4 for tl in wrapped.values())
        self.outlier_loss = outlier_loss
        self.res = TuneResult()
        if not wrapped:
            return

        def lr_for(bits):
            if lr is not None:
                return lr
            return 2.0 / iters if (iters >= 1000 and bits <= 3) else 1.0 / iters

        # one lr tensor per param group, scaled by the same LinearLR factors (quantizer.py:374-429)
        self.groups = groups = []
        for tl in wrapped.values():
            base = lr_for(tl.scheme.bits)
            groups.append(([tl.value], torch.tensor(float(base))))
            if enable_minmax_tuning:
                mm = [tl.max_scale] if tl.scheme.data_type != "int" else [tl.min_scale, tl.max_scale]
                groups.append((mm, torch.tensor(float(minmax_lr if minmax_lr is not None else base))))
        if not enable_minmax_tuning:
            for tl in wrapped.values():
                tl.min_scale.requires_grad_(False)
                tl.max_scale.requires_grad_(False)
        gbs = min(nsamples, batch_size * self.accum)
        self.fixed_num_elm = 1
        if self.accum != 1 and not token_masks:
            self.fixed_num_elm = sum(int(inputs[i].numel()) for i in range(gbs))
        self.sampler = sampler if sampler is not None else IndexSampler(nsamples, gbs)
        self.best_loss = torch.finfo(torch.float).max

    def step(self, it):
        res, wrapped, token_masks = self.res, self.wrapped, self.token_masks
        gidx = self.sampler.next_batch()
        res.batches.append(list(gidx))
        num_elm = self.fixed_num_elm
        if token_masks:
            num_elm = sum(int(torch.count_nonzero(token_masks[i]).item()) for i in gidx)
        reduction = "sum" if self.accum != 1 else "mean"
        total = 0.0
        for start in range(0, len(gidx), self.micro_bs):
            idx = gidx[start:start + self.micro_bs]
            mask = torch.cat([token_masks[i] for i in idx], dim=0).unsqueeze(-1) if token_masks else None
            ref = torch.cat([self.fp_outputs[i] for i in idx], dim=0)
            x, sel = select_batch(self.inputs, self.others, idx)
            pred = block_forward(self.block, x, sel, self.amp)
            if self.outlier_loss:
                loss = outlier_suppressed_loss(pred, ref, mask)
            elif self.alg_ext:
                # SignRoundV2Quantizer._get_loss falls back to super()._get_loss WITHOUT forwarding valid_token_mask
                # (sign_roundv2/quantizer.py:399): plain MSE over every token; num_elm still counts valid tokens only
                loss = masked_mse(pred, ref, None, reduction)
            else:
                loss = masked_mse(pred, ref, mask, reduction)
            num_elm = 1 if num_elm <= 0 else num_elm
            total += loss.item() / num_elm
            self.micro_losses.append(float(loss.item()))
            (loss * 1000).backward()
        res.losses.append(total)
        if total < self.best_loss:
            self.best_loss = total
            if not self.not_use_best_mse:
                res.best_params = {n: {k: p.data.clone() for k, p in tl.params.items()} for n, tl in wrapped.items()}
                res.best_iter = it
        if self.not_use_best_mse and it == self.iters - 1:
            res.best_params = {n: {k: p.data.clone() for k, p in tl.params.items()} for n, tl in wrapped.items()}
            res.best_iter = it
        # sign-SGD step + zero_grad + LinearLR(1 -> 0 over iters)
        with torch.no_grad():
            for params, lr_t in self.groups:
                for p in params:
                    if p.grad is not None:
                        p.add_(torch.sign(p.grad), alpha=-lr_t.item())
                        p.grad = None
        # LinearLR(start 1 -> end 0 over `iters`), chainable form: lr *= 1 - 1/(iters - it), in the lr tensor's fp32
        iters = self.iters
        for _, lr_t in self.groups:
            lr_t.mul_(1.0 + (0.0 - 1.0) / (iters * 1.0 + it * (0.0 - 1.0)))
        return total

    def finish(self) -> "TuneResult":
        if self.wrapped:
            self.res.best_loss = self.best_loss
            with torch.no_grad():
                unwrap_block(self.block, self.wrapped, self.res.best_params)
        return self.res


def tune_block(block, inputs, others, fp_outputs, scheme_of, iters=200, batch_size=8, lr=None, minmax_lr=None,
               token_masks=None, enable_minmax_tuning=True, nv_global_scales=None, amp=True,
               not_use_best_mse=False, sampler=None, alg_ext=False, imatrices=None, outlier_loss=None,
               gradient_accumulate_steps=1) -> TuneResult:
    """quantize_block: `inputs`/`fp_outputs` are per-sample lists of [1,S,H]; `token_masks` per-sample [1,S]
    long tensors (1 = valid) or None.  Mutates `block` in place (qdq weights + scale/zp attributes)."""
    tuner = BlockTuner(block, inputs, others, fp_outputs, scheme_of, iters, batch_size, lr, minmax_lr, token_masks,
                       enable_minmax_tuning, nv_global_scales, amp, not_use_best_mse, sampler, alg_ext, imatrices, outlier_loss,
                       gradient_accumulate_steps)
    if not tuner.wrapped:
        return tuner.res
    for it in range(iters):
        tuner.step(it)
    res = tuner.finish()
    res.micro_losses = tuner.micro_losses
    return res


# --------------------------------------------------------------------------------------------
# quantize_layer_outside_block -- auto_round/algorithms/quantization/sign_round/quantizer.py:554-759 (lm_head with
# quant_lm_head=True).  Differences from the block loop: micro-batches of ONE sample with gradient accumulation over the
# `batch_size` samples of an iteration, MSELoss(reduction="sum"), the FP target is recomputed each step with the FP weight on
# the FP input, `num_elm` is fixed BEFORE the loop from the first `batch_size` samples (valid tokens when a mask exists,
# input elements otherwise), and the logged loss of an iteration is the sum of loss.item()/num_elm over its micro-batches.
# --------------------------------------------------------------------------------------------
def tune_layer(linear: nn.Linear, fp_inputs, q_inputs, scheme: LayerScheme, iters=200, batch_size=8, lr=None, minmax_lr=None,
               token_masks=None, enable_minmax_tuning=True, amp=True, amp_dtype=torch.bfloat16, not_use_best_mse=False,
               sampler=None) -> TuneResult:
    res = TuneResult()
    nsamples = len(fp_inputs)
    dt = linear.weight.dtype
    fp_inputs = [t.to(dt) for t in fp_inputs]
    q_inputs = None if q_inputs is None else [t.to(dt) for t in q_inputs]
    tl = TunableLinear(linear, scheme)
    base = lr if lr is not None else (2.0 / iters if (iters >= 1000 and scheme.bits <= 3) else 1.0 / iters)
    groups = [([tl.value], torch.tensor(float(base)))]
    if enable_minmax_tuning:
        groups.append(([tl.min_scale, tl.max_scale], torch.tensor(float(minmax_lr if minmax_lr is not None else base))))
    else:
        tl.min_scale.requires_grad_(False)
        tl.max_scale.requires_grad_(False)
    gas = batch_size                                            # gradient_accumulate_steps = batch_size * 1 (:660-662)
    gbs = min(nsamples, gas)
    num_elm = 1
    if gas != 1:
        whole = list(range(gbs))
        if token_masks:
            num_elm = sum(int(torch.count_nonzero(token_masks[i]).item()) for i in whole)
        else:
            src = q_inputs if q_inputs is not None else fp_inputs
            num_elm = sum(int(src[i].numel()) for i in whole)
    reduction = "sum" if gas != 1 else "mean"
    sampler = sampler if sampler is not None else IndexSampler(nsamples, gbs)
    best_loss = torch.finfo(torch.float).max
    micro_losses = []
    for it in range(iters):
        total = 0.0
        idx = sampler.next_batch()
        res.batches.append(list(idx))
        for i in idx:
            cur_in = (q_inputs if q_inputs is not None else fp_inputs)[i]
            org_in = fp_inputs[i]
            with torch.no_grad():
                target = linear(org_in)
            mask = token_masks[i].unsqueeze(-1) if token_masks else None
            ctx = torch.autocast(device_type="cpu", dtype=amp_dtype) if amp else _null()
            with ctx:
                out = tl(cur_in)
                if mask is not None:
                    loss = F.mse_loss((out * mask).to(torch.float32), (target * mask).to(torch.float32), reduction=reduction)
                else:
                    loss = F.mse_loss(out.to(torch.float32), target.to(torch.float32), reduction=reduction)
            num_elm = 1 if num_elm <= 0 else num_elm
            total += loss.item() / num_elm
            micro_losses.append(float(loss.item()))
            (loss * 1000).backward()
        res.losses.append(total)
        if total < best_loss:
            best_loss = total
            if not not_use_best_mse:
                res.best_params = {k: p.data.clone() for k, p in tl.params.items()}
                res.best_iter = it
        if not_use_best_mse and it == iters - 1:
            res.best_params = {k: p.data.clone() for k, p in tl.params.items()}
            res.best_iter = it
        with torch.no_grad():
            for params, lr_t in groups:
                for p in params:
                    if p.grad is not None:
                        p.add_(torch.sign(p.grad), alpha=-lr_t.item())
                        p.grad = None
        for _, lr_t in groups:
            lr_t.mul_(1.0 + (0.0 - 1.0) / (iters * 1.0 + it * (0.0 - 1.0)))
    res.best_loss = best_loss
    res.micro_losses = micro_losses
    with torch.no_grad():
        tl.bake(res.best_params)
    return res


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
