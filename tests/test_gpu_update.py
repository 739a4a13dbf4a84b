"""GPU parity of the fused per-layer update kernel (ar_fq_update, csrc/ar_qdq.cu) against the oracle:

    Gq (bf16 dL/dWq) --autograd through the reference's qdq function--> dV, d min/max_scale      (oracle/qdq.py, autograd)
    p <- p - lr * sign(grad), scales clamped                                                      (sign_sgd.py:369-389)
    Wq' = qdq(W; V', scales')                                                                     (wrapper.py:244-293)

Bars: V' bit-exact (dV = Gq*s*mask is the same fp32 product on both sides, only its sign is used); the scale gradients are
group sums whose order differs from torch's, so a scale parameter may step the other way when its gradient is ~0: at most
0.5 % of the groups; Wq' bit-exact to the oracle's qdq of the parameters the kernel wrote, and bit-exact to the oracle's own
Wq' on every group whose scale parameters agree.  Also: snapshot semantics, row-shard form (data-parallel rank), has_grad."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from auto_round_b200 import ops  # noqa: E402
from oracle import qdq as Q  # noqa: E402
from oracle import signround as S  # noqa: E402

DEV = torch.device("cuda", 0)

CASES = [  # name, qdq name, bits, g, n, k, with init (alg_ext)
    ("int_sym_w4g128", "int_sym", 4, 128, 32, 512, False), ("int_sym_w2g32", "int_sym", 2, 32, 16, 256, False),
    ("int_sym_w4g64", "int_sym", 4, 64, 16, 256, False), ("int_sym_w8g256", "int_sym", 8, 256, 16, 512, False),
    ("int_asym_w2g32", "int_asym", 2, 32, 16, 256, False), ("int_asym_w4g128", "int_asym", 4, 128, 16, 256, False),
    ("mx_fp4", "mx_fp4", 4, 32, 16, 256, False), ("nv_fp4", "nv_fp4", 4, 16, 16, 256, False),
    ("int_sym_w4g128_kpad", "int_sym", 4, 128, 16, 200, False),
    ("int_sym_w2g32_init", "int_sym", 2, 32, 16, 256, True), ("mx_fp4_init", "mx_fp4", 4, 32, 16, 256, True),
    ("nv_fp4_init", "nv_fp4", 4, 16, 16, 256, True),
    # per-row groups: group_size = -1, and a weight narrower than the group (data_type/utils.py:57-61)
    ("int_sym_w4_row", "int_sym", 4, -1, 24, 512, False), ("int_asym_w8_row", "int_asym", 8, -1, 16, 1024, False),
    ("int_sym_w4_k_lt_g", "int_sym", 4, 128, 16, 64, False),
]


def _oracle_qdq(qname, w, bits, g, v, mn, mx, init, gs):
    if qname == "int_sym":
        return Q.int_sym(w, bits, g, v, mn, mx, init_scale=init) if init is not None else Q.int_sym(w, bits, g, v, mn, mx)
    if qname == "int_asym":
        return Q.int_asym(w, bits, g, v, mn, mx)
    if qname == "mx_fp4":
        return Q.mx_fp4(w, g, v, mx, init_scale=init)
    return Q.nv_fp4(w, g, v, gs, mx, init_scale=1.0 if init is None else init)


def _case(qname, bits, g, n, k, with_init, seed):
    gen = torch.Generator().manual_seed(seed)
    w = (torch.randn(n, k, generator=gen) * 0.05).bfloat16()
    w[1, :4] = torch.tensor([0.25, -0.25, 0.1, 0.0]).bfloat16()
    w[2, 7] = 2.0
    grp, _, _ = Q.to_groups(w, g)
    groups, hi = grp.shape[0], (2.0 if with_init else 1.0)
    v = ((torch.rand(grp.shape, generator=gen) - 0.5) * 0.8).float()
    mn = (0.5 + 0.5 * torch.rand(groups, generator=gen)).float()
    mx = (0.4 * hi + 0.6 * hi * torch.rand(groups, generator=gen)).float()
    mx[:3] = hi                                     # at the upper clamp: a negative gradient must not push it out
    init = None
    if with_init:
        sc = S.LayerScheme(bits, g, True, {"int_sym": "int", "mx_fp4": "mx_fp", "nv_fp4": "nv_fp"}[qname])
        im = (torch.rand(k, generator=gen) ** 2 * 40 + 0.01).float()
        init = S.search_init_scale(w, sc, im, 1e-5)
    gs = Q.nv_global_scale(w).reshape(1) if qname == "nv_fp4" else None
    gq = (torch.randn(n, k, generator=gen) * 1e-3).bfloat16()
    return w, v, mn, mx, init, gs, gq, hi


def _oracle_step(qname, bits, g, w, v, mn, mx, init, gs, gq, lr_v, lr_s, hi):
    vr, mnr, mxr = v.clone().requires_grad_(), mn.clone().requires_grad_(), mx.clone().requires_grad_()
    wq, _, _ = _oracle_qdq(qname, w, bits, g, vr, mnr, mxr, init, gs)
    (wq.to(w.dtype) * gq).sum().backward()          # the weight gradient arrives as a bf16 tensor (autograd of F.linear)
    with torch.no_grad():
        v2 = v - lr_v * torch.sign(torch.nan_to_num(vr.grad))
        mx2 = (mx - lr_s * torch.sign(torch.nan_to_num(mxr.grad))).clamp_(0, hi)
        mn2 = mn if mnr.grad is None else (mn - lr_s * torch.sign(torch.nan_to_num(mnr.grad))).clamp_(0, hi)
        wq2, _, _ = _oracle_qdq(qname, w, bits, g, v2, mn2, mx2, init, gs)
    return vr.grad, mxr.grad, mnr.grad, v2, mn2, mx2, wq2.to(w.dtype)


def _device_state(spec, w, v, mn, mx, init, gs, gq):
    n, k = w.shape
    d = dict(w=w.to(DEV), v=v.reshape(n, spec.kpad).to(DEV).contiguous(), mx=mx.to(DEV), gq=gq.to(DEV),
             mn=mn.to(DEV) if spec.is_int else None, gs=None if gs is None else gs.to(DEV),
             init=None if init is None else init.float().reshape(-1).to(DEV), wmin=None, wmax=None)
    if spec.is_int:
        d["wmin"], d["wmax"] = ops.group_minmax(spec, d["w"])
    d["wq"] = torch.zeros(n, k, dtype=torch.bfloat16, device=DEV)
    d["best_v"], d["best_mx"] = torch.full_like(d["v"], -7.0), torch.full_like(d["mx"], -7.0)
    d["best_mn"] = None if d["mn"] is None else torch.full_like(d["mn"], -7.0)
    return d


def _run(spec, d, lr_tab, flag, hi, **kw):
    ops.fq_update(spec, d["w"], d["v"], d["mn"], d["mx"], d["wmin"], d["wmax"], d["gs"], kw.pop("gq", d["gq"]), d["wq"], lr_tab,
                  best_v=d["best_v"], best_min=d["best_mn"], best_max=d["best_mx"], flag=flag, it=1, clamp_hi=hi,
                  init_scale=d["init"], **kw)
    torch.cuda.synchronize()


@pytest.mark.parametrize("name,qname,bits,g,n,k,with_init", CASES)
def test_fused_update_matches_oracle(name, qname, bits, g, n, k, with_init):
    w, v, mn, mx, init, gs, gq, hi = _case(qname, bits, g, n, k, with_init, seed=len(name))
    lr_v, lr_s = 0.004, 0.003
    dv_ref, dmx_ref, dmn_ref, v2, mn2, mx2, wq2 = _oracle_step(qname, bits, g, w, v, mn, mx, init, gs, gq, lr_v, lr_s, hi)
    spec = ops.make_spec(qname, bits, g, n, k, 1e-5, hi)
    d = _device_state(spec, w, v, mn, mx, init, gs, gq)
    lr_tab = torch.tensor([9.0, 9.0, lr_v, lr_s], device=DEV)          # row `it` = 1 is the one that must be read
    flag = torch.ones(1, dtype=torch.int32, device=DEV)
    dbg = (torch.empty(n, spec.kpad, device=DEV), torch.empty(spec.groups, device=DEV), torch.empty(spec.groups, device=DEV))
    _run(spec, d, lr_tab, flag, hi, dbg=dbg)
    # pre-sign gradients: dV bit-exact, scale gradients within 2e-3 of the largest (fp32 sums in another order)
    ref_dv = torch.nan_to_num(dv_ref).reshape(n, spec.kpad)
    assert torch.allclose(dbg[0].cpu(), ref_dv, rtol=2e-5, atol=0), name
    fin = torch.isfinite(dmx_ref)
    assert float((dbg[2].cpu() - dmx_ref)[fin].abs().max() / dmx_ref[fin].abs().max().clamp_min(1e-20)) <= 2e-3
    # snapshot = the PRE-update parameters
    assert torch.equal(d["best_v"].cpu().reshape(v.shape), v) and torch.equal(d["best_mx"].cpu(), mx)
    if spec.is_int:
        assert torch.equal(d["best_mn"].cpu(), mn)
    # the step
    assert torch.equal(d["v"].cpu().reshape(v2.shape), v2), name
    same = d["mx"].cpu() == mx2
    if spec.is_int:
        same &= d["mn"].cpu() == mn2
    assert float(same.float().mean()) >= 0.995, (name, float(same.float().mean()))
    assert float(d["mx"].max()) <= hi and float(d["mx"].min()) >= 0.0
    # next iteration's fake-quant weight: bit-exact to the oracle's qdq of the parameters the kernel wrote ...
    mn_dev = mn if d["mn"] is None else d["mn"].cpu()
    wq_own, _, _ = _oracle_qdq(qname, w, bits, g, d["v"].cpu().reshape(v.shape), mn_dev, d["mx"].cpu(), init, gs)
    assert torch.equal(d["wq"].cpu(), wq_own.to(w.dtype)), name
    # ... and to the oracle's own step wherever the scale parameters agree
    gpr = spec.kpad // spec.group_size
    rows_ok = same.reshape(n, gpr).all(dim=1)
    assert torch.equal(d["wq"].cpu()[rows_ok], wq2[rows_ok])


def test_fused_update_flag_off_and_has_grad():
    name, qname, bits, g, n, k, with_init = CASES[0]
    w, v, mn, mx, init, gs, gq, hi = _case(qname, bits, g, n, k, with_init, seed=5)
    spec = ops.make_spec(qname, bits, g, n, k, 1e-5, hi)
    lr_tab = torch.tensor([9.0, 9.0, 0.004, 0.003], device=DEV)
    d = _device_state(spec, w, v, mn, mx, init, gs, gq)
    _run(spec, d, lr_tab, torch.zeros(1, dtype=torch.int32, device=DEV), hi)
    assert float(d["best_v"].min()) == -7.0 and float(d["best_mx"].min()) == -7.0        # no snapshot without the flag
    assert not torch.equal(d["v"].cpu().reshape(v.shape), v)                             # but the step happened
    d = _device_state(spec, w, v, mn, mx, init, gs, gq)
    _run(spec, d, lr_tab, torch.ones(1, dtype=torch.int32, device=DEV), hi, has_grad=torch.zeros(1, dtype=torch.int32, device=DEV))
    assert torch.equal(d["v"].cpu().reshape(v.shape), v) and torch.equal(d["mx"].cpu(), mx)   # layer without a gradient: not
    assert float(d["wq"].abs().max()) == 0.0                                                  # stepped, no new Wq ...
    assert torch.equal(d["best_v"].cpu().reshape(v.shape), v) and torch.equal(d["best_mx"].cpu(), mx)   # ... but snapshotted
    d = _device_state(spec, w, v, mn, mx, init, gs, gq)
    _run(spec, d, lr_tab, torch.zeros(1, dtype=torch.int32, device=DEV), hi, has_grad=torch.zeros(1, dtype=torch.int32, device=DEV))
    assert float(d["best_v"].min()) == -7.0 and torch.equal(d["v"].cpu().reshape(v.shape), v)  # no flag: nothing at all


@pytest.mark.parametrize("name,qname,bits,g,n,k,with_init", [CASES[0], CASES[4], CASES[7]])
def test_fused_update_row_shards_equal_full(name, qname, bits, g, n, k, with_init):
    """A data-parallel rank updates rows [r0, r1) from its reduce-scattered shard of dWq: the union of the shards must
    equal the full update bit-for-bit, and rows outside the shard stay untouched."""
    w, v, mn, mx, init, gs, gq, hi = _case(qname, bits, g, n, k, with_init, seed=11)
    spec = ops.make_spec(qname, bits, g, n, k, 1e-5, hi)
    lr_tab = torch.tensor([9.0, 9.0, 0.004, 0.003], device=DEV)
    flag = torch.ones(1, dtype=torch.int32, device=DEV)
    full = _device_state(spec, w, v, mn, mx, init, gs, gq)
    _run(spec, full, lr_tab, flag, hi)
    world = 4
    per = n // world
    sh = _device_state(spec, w, v, mn, mx, init, gs, gq)
    for r in range(world):
        r0, r1 = r * per, (r + 1) * per
        if r == 1:                                  # nothing outside the shard may change
            before = {key: sh[key].clone() for key in ("v", "mx", "wq")}
        shard = sh["gq"][r0:r1].contiguous()
        _run(spec, sh, lr_tab, flag, hi, gq=shard, row0=r0, row1=r1, gq_row0=r0)
        if r == 1:
            keep = torch.ones(n, dtype=torch.bool, device=DEV)
            keep[r0:r1] = False
            assert torch.equal(sh["v"][keep], before["v"][keep]) and torch.equal(sh["wq"][keep], before["wq"][keep])
    for key in ("v", "mx", "mn", "wq", "best_v", "best_mx"):
        if full[key] is not None:
            assert torch.equal(full[key], sh[key]), key


@pytest.mark.parametrize("qname,bits,g,n,k", [("int_sym", 4, -1, 24, 512), ("int_asym", 2, -1, 16, 256), ("int_sym", 8, 128, 8, 64)])
def test_per_row_groups_forward_bit_exact(qname, bits, g, n, k):
    """group_size = -1 / K < group_size: one group per row -- group min/max, fake-quant weight, scale and zero-point bit-exact
    to the oracle (RTN form and with V / min_scale / max_scale)."""
    gen = torch.Generator().manual_seed(n + k)
    w = (torch.randn(n, k, generator=gen) * 0.05).bfloat16()
    spec = ops.make_spec(qname, bits, g, n, k)
    assert spec.group_size == k and spec.groups == n
    wmin, wmax = ops.group_minmax(spec, w.to(DEV))
    rmin, rmax = Q.group_minmax(w, g)
    assert torch.equal(wmin.cpu(), rmin.reshape(-1)) and torch.equal(wmax.cpu(), rmax.reshape(-1))
    v = ((torch.rand(n, k, generator=gen) - 0.5) * 0.8).float()
    mn, mx = (0.5 + 0.5 * torch.rand(n, generator=gen)).float(), (0.5 + 0.5 * torch.rand(n, generator=gen)).float()
    fn = Q.int_sym if qname == "int_sym" else Q.int_asym
    ref_q, ref_s, ref_zp = fn(w, bits, g, v, mn, mx)
    wq, sc, zp = ops.qdq_fwd(spec, w.to(DEV), v.to(DEV), mn.to(DEV), mx.to(DEV), wmin, wmax, None, want_scale=True)
    assert torch.equal(wq.cpu(), ref_q.to(w.dtype))
    assert torch.equal(sc.float().cpu().reshape(-1), ref_s.float().reshape(-1))
    if qname == "int_asym":
        assert torch.equal(zp.cpu().reshape(-1), ref_zp.float().reshape(-1))
    ref_q0, ref_s0, _ = fn(w, bits, g)                                    # plain RTN (iters = 0 / unwrap without parameters)
    wq0, sc0, _ = ops.qdq_fwd(spec, w.to(DEV), want_scale=True)
    assert torch.equal(wq0.cpu(), ref_q0.to(w.dtype)) and torch.equal(sc0.float().cpu().reshape(-1), ref_s0.float().reshape(-1))


@pytest.mark.parametrize("name,qname,bits,g,n,k,with_init", [c for c in CASES if c[3] > 0 and c[2] <= 4 and c[5] >= c[3]])
def test_wire_form_decodes_to_the_same_weight(name, qname, bits, g, n, k, with_init):
    """Data-parallel exchange of the next fake-quant weight as 4-bit codes + per-group {a, off}: the decode kernel must
    rebuild the bf16 weight the plain path writes -- every value identical; for the fp4 formats the bit patterns too (their
    code carries the sign bit), for int sym a zero may come back as +0 where s * (-0) was -0 (16 codes are all in use;
    a zero's sign changes no product)."""
    w, v, mn, mx, init, gs, gq, hi = _case(qname, bits, g, n, k, with_init, seed=23)
    w[3, :8] = 0                                     # zeros (and -0 after rounding) must survive the wire
    spec = ops.make_spec(qname, bits, g, n, k, 1e-5, hi)
    assert ops.wire_supported(spec)
    lr_tab = torch.tensor([9.0, 9.0, 0.004, 0.003], device=DEV)
    flag = torch.ones(1, dtype=torch.int32, device=DEV)
    full = _device_state(spec, w, v, mn, mx, init, gs, gq)
    _run(spec, full, lr_tab, flag, hi)
    world = 4
    per = n // world
    seg = ops.wire_segment_bytes(spec, per)
    wire = torch.zeros(world * seg, dtype=torch.uint8, device=DEV)
    sh = _device_state(spec, w, v, mn, mx, init, gs, gq)
    for r in range(world):
        r0, r1 = r * per, (r + 1) * per
        ops.fq_update(spec, sh["w"], sh["v"], sh["mn"], sh["mx"], sh["wmin"], sh["wmax"], sh["gs"], sh["gq"][r0:r1].contiguous(), None,
                      lr_tab, best_v=sh["best_v"], best_min=sh["best_mn"], best_max=sh["best_mx"], flag=flag, it=1, clamp_hi=hi,
                      init_scale=sh["init"], row0=r0, row1=r1, gq_row0=r0, wire=wire[r * seg:(r + 1) * seg])
    out = torch.zeros(n, k, dtype=torch.bfloat16, device=DEV)
    ops.wq_decode(spec, wire, world, out)
    torch.cuda.synchronize()
    assert torch.equal(out, full["wq"]), name
    nz = full["wq"] != 0
    assert torch.equal(out.view(torch.int16)[nz], full["wq"].view(torch.int16)[nz]), name
    if qname in ("mx_fp4", "nv_fp4", "int_asym"):
        assert torch.equal(out.view(torch.int16), full["wq"].view(torch.int16)), name    # bit patterns, signed zeros included
    for key in ("v", "mx", "mn", "best_v"):
        if full[key] is not None:
            assert torch.equal(full[key], sh[key]), key
    assert seg * world < n * k * 2 * (0.5 if g >= 32 else 0.8)                          # what the exchange saves
