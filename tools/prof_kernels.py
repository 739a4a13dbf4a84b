"""Launch each hot kernel of the path once at Llama-3-8B shapes (after one warm launch) so that

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels python tools/prof_kernels.py

captures them in a few seconds.  `--mix` instead launches the 18 GEMMs of ONE sign-SGD iteration (7 forward, 4 grad-in,
7 grad-w at T = 16384) after a warm pass: captured with `-k regex:gemm_kernel -s 18 -c 18` it gives the DRAM traffic per
launch that bench.py reports as `roofline.traffic` (tools/summarize_ncu.py writes profiles/r02_gemm_traffic.json).
Not the bench: a profiling driver."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from auto_round_b200 import ops  # noqa: E402

dev = "cuda"
T = 16384

if "--mix" in sys.argv:
    one_iter, flops, nlaunch = bench._gemm_mix(torch.device(dev), bench.CONFIGS["llama3_8b_w4a16"])
    one_iter()
    torch.cuda.synchronize()
    one_iter()
    torch.cuda.synchronize()
    print("mix done: %d launches, %.4e FLOP" % (nlaunch, flops))
    sys.exit(0)

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
gq = torch.empty(n, k, device=dev, dtype=torch.bfloat16)
flag = torch.ones(1, dtype=torch.int32, device=dev)
best_v, best_mn, best_mx = torch.empty_like(v), torch.empty_like(mn), torch.empty_like(mx)
lr = torch.tensor([0.005, 0.005], device=dev)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
pred = torch.randn(T, 4096, device=dev).bfloat16()
ref = torch.randn(T, 4096, device=dev).bfloat16()
mask = torch.ones(T, dtype=torch.uint8, device=dev)
nvs = ops.make_spec("nv_fp4", 4, 16, n, k)
nv_gs = ops.nv_global_scale(w)
nv_mx = torch.ones(nvs.groups, device=dev)
nv_v = torch.zeros(n, k, device=dev)

for rep in range(2):          # rep 0 = warm, rep 1 = the launch to read in the report
    ops.qdq_fwd(spec, w, v, mn, mx, wmin, wmax, None, out_wq=wq)
    ops.gemm(x, wq, out=y)                                   # forward   (K-major / K-major)
    ops.gemm(dy, wq, False, True, out=dx)                    # grad-in   (K-major / MN-major)
    ops.gemm(dy, x, True, True, out=gq)                      # grad-w    (MN-major / MN-major), bf16 dWq
    ops.fq_update(spec, w, v, mn, mx, wmin, wmax, None, gq, wq, lr, best_v=best_v, best_min=best_mn, best_max=best_mx, flag=flag)
    ops.fq_update(nvs, w, nv_v, None, nv_mx, None, None, nv_gs, gq, wq, lr, best_v=best_v, best_max=best_mx, flag=flag)
    ops.mse_fwd_bwd(pred, ref, mask, 1.0 / pred.numel(), 1000.0, loss)
    wq2, scale, _ = ops.qdq_fwd(spec, w, v, mn, mx, wmin, wmax, None, want_scale=True)
    ops.pack_int(wq2, scale.reshape(n, -1).contiguous(), None, 4, 128, True, zp_const=8)
    wq3, sc3, _ = ops.qdq_fwd(nvs, w, None, None, nv_mx, None, None, nv_gs, want_scale=True)
    ops.pack_fp4_nv(wq3, sc3.reshape(n, -1).contiguous(), nv_gs)
torch.cuda.synchronize()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# CUDA-event timings (a plain run doubles as the measurement): the HBM-bound kernels against their algorithmic bytes
P = n * k
t_upd = timed(lambda: ops.fq_update(spec, w, v, mn, mx, wmin, wmax, None, gq, wq, lr, best_v=best_v, best_min=best_mn, best_max=best_mx, flag=flag))
flag0 = torch.zeros(1, dtype=torch.int32, device=dev)
t_upd0 = timed(lambda: ops.fq_update(spec, w, v, mn, mx, wmin, wmax, None, gq, wq, lr, best_v=best_v, best_min=best_mn, best_max=best_mx, flag=flag0))
t_updnv = timed(lambda: ops.fq_update(nvs, w, nv_v, None, nv_mx, None, None, nv_gs, gq, wq, lr, best_v=best_v, best_max=best_mx, flag=flag0))
t_qdq = timed(lambda: ops.qdq_fwd(spec, w, v, mn, mx, wmin, wmax, None, out_wq=wq))
sc16 = scale.reshape(n, -1).contiguous()
t_pack = timed(lambda: ops.pack_int(wq2, sc16, None, 4, 128, True, zp_const=8))
sc3c = sc3.reshape(n, -1).contiguous()
t_packnv = timed(lambda: ops.pack_fp4_nv(wq3, sc3c, nv_gs))
print("fq_update int_sym g128, snapshot on : %.1f us = %.2f TB/s of 18 B/weight" % (t_upd * 1e3, P * 18 / t_upd / 1e9))
print("fq_update int_sym g128, snapshot off: %.1f us = %.2f TB/s of 14 B/weight" % (t_upd0 * 1e3, P * 14 / t_upd0 / 1e9))
print("fq_update nv_fp4 g16,   snapshot off: %.1f us = %.2f TB/s of 14 B/weight" % (t_updnv * 1e3, P * 14 / t_updnv / 1e9))
print("qdq_fwd   int_sym g128              : %.1f us = %.2f TB/s of 8 B/weight" % (t_qdq * 1e3, P * 8 / t_qdq / 1e9))
print("pack_int4 (3 launches)              : %.1f us = %.2f TB/s of 2.53 B/weight" % (t_pack * 1e3, P * 2.53 / t_pack / 1e9))
print("pack_fp4_nv                         : %.1f us = %.2f TB/s of 2.56 B/weight" % (t_packnv * 1e3, P * 2.5625 / t_packnv / 1e9))

# optimized-RTN kernels at the same layer
imx = torch.zeros(k, dtype=torch.float32, device=dev)
mxs = ops.make_spec("mx_fp4", 4, 32, n, k)
t_im = timed(lambda: ops.imatrix_accum(x, imx), 3)
qw = imx / 8.0 + 1e-3
t_int = timed(lambda: ops.search_scale_int(spec, w, qw), 3)
t_nv = timed(lambda: ops.search_scale_nv(nvs, w, qw), 3)
t_mx = timed(lambda: ops.search_scale_mx(mxs, w, qw), 3)
cand = len(ops.int_search_table(4))
print("imatrix_accum  [%d x %d] bf16: %.3f ms = %.2f TB/s" % (T, k, t_im, T * k * 2 / t_im / 1e9))
print("search_scale_int W4 g128 [%d x %d], %d candidates: %.3f ms = %.1f G candidate-weights/s" % (n, k, cand, t_int, n * k * cand / t_int / 1e6))
print("search_scale_nv  g16, %d candidates (+ absmax pass): %.3f ms" % (len(ops.NV_SEARCH_TABLE), t_nv))
print("search_scale_mx  g32, 3 candidates: %.3f ms" % t_mx)
print("done")
