"""CPU: the JSON-line contract of bench.py's reference arm (`--impl reference`), at a tiny scale. The arm runs the
UNMODIFIED reference (oracle/_ref) on host cores; no GPU, none of the product's kernels."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.ref
def test_reference_arm_prints_one_contract_line():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref was not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.002",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:] + out.stderr[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["metric"].startswith("time-to-top-k PCs") and d["unit"] == "GB/s" and d["steps"] == 2
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["time_to_pcs_s"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
