"""Differential fuzz of the oracle against the unmodified reference, LIVE (only where /root/reference exists: the build
container's CPU tier; skipped on the GPU box).  Complements the committed fixtures: fresh random inputs, many seeds, values
and autograd gradients of every fake-quant function plus the RTN / optimized-RTN routes must agree bit for bit."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/auto_round"), reason="reference tree not present (GPU box)")
def test_oracle_equals_live_reference_on_random_inputs():
    p = subprocess.run([sys.executable, "-m", "oracle.diff_fuzz", "--seeds", "6"], cwd=ROOT, capture_output=True, text=True,
                       timeout=900)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert line, p.stderr[-2000:]
    res = json.loads(line[-1])
    assert res["cases"] >= 130
    assert res["failures"] == [], res["failures"][:10]
