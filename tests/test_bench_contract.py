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


def test_parity_block_flags_a_single_wrong_pixel(port, S):
    """bench.py's `parity` block: per-row CRCs of the timed frame against the unmodified reference's (fullsize_c2.json).
    The oracle's own 1080p frame passes; one flipped bit in one pixel is one bad row."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    spec = bench.workload_spec("c2", 1)
    frame = port.render(S.default_scene(), S.default_camera(1920, 16.0 / 9.0).pod(), 8, want=("rgba8",))["rgba8"]
    ok = bench.parity_check(spec, {"device frame": frame, "host frame": frame[None]})
    assert ok["rows_checked"] == 2 * 1080 and ok["rows_bad"] == 0 and "fullsize_c2.json" in ok["source"]
    bad = frame.copy()
    bad[540, 960] ^= 0x100
    out = bench.parity_check(spec, {"device frame": bad})
    assert out["rows_bad"] == 1 and out["first_bad"] == [["device frame", 0, 540]]
    small = bench.parity_check(spec, {"device frame": frame[:100]})
    assert small["rows_bad"] is None and small["rows_checked"] == 0          # a scaled run is not compared, and says so


def test_both_arms_print_the_same_structural_config():
    sys.path.insert(0, ROOT)
    import importlib
    import bench
    S = importlib.import_module("ray-tracer-from-scratch_b200").scene
    spec = bench.workload_spec("auto", 8)
    scene = bench.build_scene(S, spec["scene"])
    pods = [S.default_camera(spec["width"], 16.0 / 9.0).pod()]
    cfg = bench.structural_config(spec, S, pods, scene, 8, 4)
    assert {k: cfg[k] for k in ("workload", "width", "height", "frames_per_step", "depth", "n_spheres", "n_walls", "band_rows")} == {
        "workload": spec["label"], "width": 7680, "height": 4320, "frames_per_step": 1, "depth": 10, "n_spheres": 10000, "n_walls": 64, "band_rows": 4}
    assert "L2 flush" in cfg["l2"] and set(cfg) == {"workload", "width", "height", "frames_per_step", "depth", "n_spheres", "n_walls", "band_rows", "l2", "arms"}
    assert bench.host_threads() >= 1
