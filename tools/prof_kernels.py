"""Launch each hot kernel of the path once at Llama-3-8B shapes (after one warm launch) so that

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels python tools/prof_kernels.py

captures them in a few seconds.  Not the bench: a profiling driver."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auto_round_b200 import ops  # noqa: E402

dev = "cuda"
T = 16384
n, k = 14336, 4096            # gate_proj
spec = ops.make_spec("int_sym", 4, 128, n, k)
w = (torch.randn(n, k, device=dev) * 0.02).bfloat16()
x = torch.randn(T, k, device=dev).bfloat16()
dy = torch.randn(T, n, device=dev).bfloat16()
v = (torch.rand(n, k, device=dev) - 0.5)
mn = torch.ones(spec.groups, device=dev)
mx = torch.ones(spec.groups, device=dev)
wmin, wmax = ops.group_minmax(spec, w)
wq = torch.empty_like(w)
y = torch.empty(T, n, device=dev, dtype=torch.bfloat16)
dx = torch.empty(T, k, device=dev, dtype=torch.bfloat16)
dv = torch.empty(n, k, device=dev)
dmn = torch.empty(spec.groups, device=dev)
dmx = torch.empty(spec.groups, device=dev)
gq = torch.randn(n, k, device=dev)
flag = torch.ones(1, dtype=torch.int32, device=dev)
best = torch.empty_like(v)
lr = torch.tensor([0.005, 0.005], device=dev)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
pred = torch.randn(T, 4096, device=dev).bfloat16()
ref = torch.randn(T, 4096, device=dev).bfloat16()
mask = torch.ones(T, dtype=torch.uint8, device=dev)

for rep in range(2):          # rep 0 = warm, rep 1 = the launch to read in the report
    ops.qdq_fwd(spec, w, v, mn, mx, wmin, wmax, None, out_wq=wq)
    ops.gemm(x, wq, out=y)                                   # forward   (K-major / K-major)
    ops.gemm(dy, wq, False, True, out=dx)                    # grad-in   (K-major / MN-major)
    ops.fq_linear_bwd_dw(spec, dy, x, w, v, mn, mx, wmin, wmax, None, dv, dmn, dmx)   # grad-w, fused epilogue
    ops.qdq_bwd(spec, w, gq, v, mn, mx, wmin, wmax, None, dv=dv, dmin=dmn, dmax=dmx)
    ops.signsgd_step(v.view(-1), dv.view(-1), best.view(-1), flag, lr, 0, clamp_begin=v.numel())
    ops.mse_fwd_bwd(pred, ref, mask, 1.0 / pred.numel(), 1000.0, loss)
    wq2, scale, _ = ops.qdq_fwd(spec, w, v, mn, mx, wmin, wmax, None, want_scale=True)
    ops.pack_int(wq2, scale.reshape(n, -1).contiguous(), None, 4, 128, True, zp_const=8)
torch.cuda.synchronize()

# optimized-RTN kernels at the same layer: importance accumulation over one batch of tokens and the
# 201-candidate scale search; CUDA-event timings printed so that a plain run doubles as the measurement
imx = torch.zeros(k, dtype=torch.float32, device=dev)
nv = ops.make_spec("nv_fp4", 4, 16, n, k)
mxs = ops.make_spec("mx_fp4", 4, 32, n, k)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_im = timed(lambda: ops.imatrix_accum(x, imx))
qw = imx / 8.0 + 1e-3
t_int = timed(lambda: ops.search_scale_int(spec, w, qw))
t_nv = timed(lambda: ops.search_scale_nv(nv, w, qw))
t_mx = timed(lambda: ops.search_scale_mx(mxs, w, qw))
cand = len(ops.int_search_table(4))
print("imatrix_accum  [%d x %d] bf16: %.3f ms = %.2f TB/s" % (T, k, t_im, T * k * 2 / t_im / 1e9))
print("search_scale_int W4 g128 [%d x %d], %d candidates: %.3f ms = %.1f G candidate-weights/s" % (n, k, cand, t_int, n * k * cand / t_int / 1e6))
print("search_scale_nv  g16, %d candidates (+ absmax pass): %.3f ms" % (len(ops.NV_SEARCH_TABLE), t_nv))
print("search_scale_mx  g32, 3 candidates: %.3f ms" % t_mx)
print("done")
