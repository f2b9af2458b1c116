"""CPU: the JSON line bench.py prints keeps the driver's contract. The reference arm (`--impl reference`: the
unmodified reference on the host, no GPU involved) is run for real on the small workload; the GPU arm's line is checked
on the copy committed under profiles/ by the last measured run."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    from oracle import refpipe
    if not refpipe.have_ref():
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["value"] > 0 and d["unit"] == "anchored k-mers/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_committed_gpu_line_has_roofline_and_baseline():
    d = json.loads((ROOT / "profiles" / "r2ak_bench_configs1.json").read_text())
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["roofline"]["physical"]["dram_bytes_per_launch"] == d["roofline"]["traffic"] > 0      # measured live by the ncu child
    assert d["roofline"]["stage"]["frac"] < d["roofline"]["frac"] < 1.0
    assert d["cpu_baseline"]["sample"].startswith("the whole workload")                               # same size as our arm
    assert d["config"]["workload"].startswith("configs[1]") and "model" not in d["config"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["traffic"] is None or rf["traffic"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
