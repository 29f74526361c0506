"""CPU suite, part 4: the bench.py contract that can be checked without a GPU — the reference arm (`--impl reference`)
runs the reference's own CPU implementation on a bounded sample and prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["workload"].startswith("c2")


def test_default_workload_is_the_8k_frame_at_every_n():
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 2, 4, 8):
        spec = bench.workload_spec("auto", n)
        assert spec["name"] == "c4" and spec["width"] == 7680 and spec["depth"] == 10
