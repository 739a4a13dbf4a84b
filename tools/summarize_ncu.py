"""Turn the ncu exports under gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_ncu.py r01        # -> profiles/r01_*.{md,csv}

Inputs (produced on the GPU box, see profiles/README.md for the exact commands):
    gpurun_out/launches_bench.csv      ncu --metrics gpu__time_duration.sum ... python bench.py ...
    gpurun_out/prof_gemm_raw.csv       ncu -i prof_gemm.ncu-rep --page raw --csv
    gpurun_out/prof_elem_raw.csv       ncu -i prof_elem.ncu-rep --page raw --csv
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]


def read_csv(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    return list(csv.reader(lines))


def to_ns(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


def launches(tag, name="launches_bench"):
    path = os.path.join(GO, f"{name}.csv")
    if not os.path.exists(path):
        return
    rows = read_csv(path)
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    seq = []
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        seq.append((short(r[ix["Kernel Name"]]), to_ns(r[ix["Metric Value"]], r[ix["Metric Unit"]])))
    tot = sum(t for _, t in seq)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in seq:
        agg[n][0] += 1
        agg[n][1] += t
    with open(os.path.join(OUT, f"{tag}_{name}.csv"), "w") as f:
        f.write("kernel,launches,total_ms,avg_us,share_pct\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{n[:160]}\",{c},{t / 1e6:.3f},{t / c / 1e3:.1f},{100 * t / tot:.2f}\n")
    # one steady-state sign-SGD iteration = launches between two consecutive ar::signsgd_kernel
    pos = [i for i, (n, _) in enumerate(seq) if "signsgd_kernel" in n]
    lines = [f"# {tag} ({name}): ncu launch list of `bench.py --steps 1 --warmup 1 --iters 10 --no-cpu-baseline` (first 6000 launches)",
             "", f"total device time {tot / 1e6:.1f} ms over {len(seq)} launches (cold-cache, serialised: compare SHARES)", ""]
    if len(pos) >= 5:
        it = seq[pos[3] + 1: pos[4] + 1]
        t_it = sum(t for _, t in it)
        ours = sum(t for n, t in it if n.startswith("ar::"))
        gemm = sum(t for n, t in it if n.startswith("ar::gemm_kernel"))
        lines += [f"## one sign-SGD iteration (Llama-3-8B block, 8x2048 tokens): {len(it)} launches, {t_it / 1e6:.2f} ms device time",
                  f"* our kernels (libar_b200.so): {100 * ours / t_it:.1f} % of the iteration, tcgen05 GEMMs alone {100 * gemm / t_it:.1f} %",
                  f"* the rest is ATen elementwise (RMSNorm / RoPE / SwiGLU / residual / casts of the HF block) and cuDNN SDPA",
                  "", "| share | us | n | kernel |", "|---|---|---|---|"]
        a2 = collections.defaultdict(lambda: [0, 0.0])
        for n, t in it:
            a2[n][0] += 1
            a2[n][1] += t
        for n, (c, t) in sorted(a2.items(), key=lambda kv: -kv[1][1])[:28]:
            lines.append(f"| {100 * t / t_it:.2f}% | {t / 1e3:.1f} | {c} | `{n[:110]}` |")
    with open(os.path.join(OUT, f"{tag}_{name}.md"), "w") as f:
        f.write("\n".join(lines) + "\n")


def raw(tag, name):
    path = os.path.join(GO, f"{name}_raw.csv")
    if not os.path.exists(path):
        return
    rows = read_csv(path)
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {tag}: `ncu --set full --clock-control none` on tools/prof_kernels.py ({name}), gate_proj shapes "
             "(N=14336, K=4096, T=16384 tokens)", ""]
    for r in rows[2:]:
        lines.append(f"## `{short(r[ix['Kernel Name']])[:120]}`")
        for k in KEYS:
            if k in ix:
                lines.append(f"* {k} = {r[ix[k]]} {units[ix[k]]}")
        lines.append("")
    with open(os.path.join(OUT, f"{tag}_ncu_{name}.md"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    raw(tag, "prof_gemm")
    raw(tag, "prof_elem")
    print(sorted(os.listdir(OUT)))
