"""CPU tier: the bench's bookkeeping that needs no GPU -- the algorithmic FLOP counts it divides by and the JSON
contract of the reference arm (`bench.py --impl reference`: same metric / unit / config as our arm, `impl`, `cpu_baseline`,
`e2e` with zero copy bytes), run on a two-step bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_flop_counts_match_the_survey():
    cfg = bench.CONFIGS["llama3_8b_w4a16"]
    p_block = sum(n * k for _, n, k, _ in cfg["linears"])
    p_qkv = sum(n * k for _, n, k, dx in cfg["linears"] if not dx)
    assert p_block == 218_103_808 and p_qkv == 25_165_824
    per_iter = 16384 * (6 * p_block - 2 * p_qkv)
    assert abs(per_iter - 2.0616e13) / 2.0616e13 < 1e-4
    step = bench.flops_per_step(cfg, 200)
    assert abs(32 * step - 1.393e17) / 1.393e17 < 2e-3                          # "Run total = 1.393e17 FLOP"
    q = bench.CONFIGS["qwen2_nvfp4"]
    assert sum(n * k for _, n, k, _ in q["linears"]) == 233_046_016
    assert bench.CONFIGS["mixtral_mxfp4"]["p_block"] == 4096 * 4096 * 2 + 1024 * 4096 * 2 + 24 * 14336 * 4096


def test_reference_arm_line_contract():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, p.stderr[-2000:]
    d = json.loads(lines[-1])
    ours = bench.CONFIGS["llama3_8b_w4a16"]
    assert d["impl"] == "reference" and d["metric"] == ours["metric"] and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["config"] == bench.line_config(ours, ours["iters"], 1, 2)           # the same workload description as our arm
    assert d["steps"] == 2 and d["warmup"] == 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and cb["steps_timed"] >= 2
    assert "EXTRAPOLATED" in cb["sample"] and cb["fit"]["layers"] == 7
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["ms_per_step"] > 0
