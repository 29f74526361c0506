"""CPU suite, part 2: the host side of the C ABI and the Python mirror of the reference interface.

No compute call is made here (there is no GPU): the library must load, export every symbol of
include/rtx_b200.h, answer its host-only entry points, and REFUSE to create a context (no CPU fallback).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, fh, fh3, load_json


@pytest.fixture(scope="module")
def lib(renderer_mod):
    if not os.path.exists(renderer_mod.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return renderer_mod.load_library()


def test_library_exports_every_declared_symbol(lib, pkg):
    header = open(os.path.join(ROOT, "include", "rtx_b200.h")).read()
    declared = set(re.findall(r"\b(rtx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(pkg.abi.EXPORTS), declared ^ set(pkg.abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rtx_abi_version() == pkg.abi.ABI_VERSION


def test_struct_sizes_match_header(pkg, tmp_path):
    """ctypes mirror vs the C header itself: sizes and the offset of every field, as gcc lays them out."""
    a = pkg.abi
    assert C.sizeof(a.Vec3) == 24 and C.sizeof(a.MaterialPOD) == 64
    pairs = [("rtx_vec3", a.Vec3), ("rtx_material", a.MaterialPOD), ("rtx_object", a.ObjectPOD), ("rtx_camera_desc", a.CameraDesc),
             ("rtx_camera", a.CameraPOD), ("rtx_params", a.Params), ("rtx_outputs", a.Outputs), ("rtx_stats", a.Stats)]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "rtx_b200.h"', "int main(void) {"]
    for cname, ct in pairs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in pairs:
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)


def test_constants_match_header(pkg):
    """Every integer #define of the header that abi.py mirrors has the header's value (RTX_ORDER_*, RTX_ACCEL_*, RTX_MEM_*, ...)."""
    header = open(os.path.join(ROOT, "include", "rtx_b200.h")).read()
    defines = {k: int(v, 0) for k, v in re.findall(r"^#define\s+(RTX_[A-Z0-9_]+)\s+(-?(?:0x[0-9a-fA-F]+|\d+))\b", header, re.M)}
    mirrored = [k for k in defines if hasattr(pkg.abi, k)]
    assert len(mirrored) >= 20, mirrored
    for k in mirrored:
        assert getattr(pkg.abi, k) == defines[k], k
    for k in ("RTX_ORDER_AUTO", "RTX_ORDER_SCAN", "RTX_ORDER_COST", "RTX_ACCEL_GRID", "RTX_MEM_HOST_MAPPED", "RTX_FRAME_COPY", "RTX_MAX_IN_FLIGHT"):
        assert k in mirrored, k
    assert defines["RTX_ABI_VERSION"] == pkg.abi.ABI_VERSION


def test_no_cpu_fallback(lib, pkg):
    """Without a GPU the product path must fail loudly instead of computing anything on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    rc = lib.rtx_create(C.byref(ctx), 0)
    assert rc == pkg.abi.RTX_ERR_CUDA and not ctx.value
    assert b"CUDA" in lib.rtx_last_error(None) or b"device" in lib.rtx_last_error(None)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    pkg_dir = os.path.join(ROOT, "ray-tracer-from-scratch_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".sh")):
                text = open(os.path.join(base, f)).read()
                assert "liboracle" not in text and "libref_oracle" not in text and "oracle." not in text.replace("oracle.c", ""), f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f


def test_default_params_are_the_reference_literals(renderer_mod, port):
    p, q = renderer_mod.default_params(), port.default_params()
    for name, _ in p._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if hasattr(a, "tuple"):
            assert a.tuple() == b.tuple(), name
        else:
            assert a == b, name
    assert p.max_depth == 10 and p.reflect_offset == .0001 and p.sky_exponent == 0.25
    assert p.light_pos.tuple() == (0, 0, 0) and p.ground_color.tuple() == (0.025, 0.05, 0.075)


def test_camera_init_three_ways(renderer_mod, port, S):
    """Camera::init (scene.cpp:80-106): library host code == Python mirror == C oracle == golden (reference)."""
    for c in load_json("kat.json")["camera"]:
        cam = S.Camera()
        cam.position, cam.lookat, cam.vup = fh3(c["position"]), fh3(c["lookat"]), fh3(c["vup"])
        cam.vfov, cam.aspect_ratio, cam.image_width = fh(c["vfov"]), fh(c["aspect_ratio"]), fh(c["image_width"])
        cam.init()
        for pod in (cam.pod(), renderer_mod.camera_init(cam), port.camera_init(cam)):
            assert pod.image_top_left.tuple() == fh3(c["image_top_left"])
            assert pod.delta_x.tuple() == fh3(c["delta_x"]) and pod.delta_y.tuple() == fh3(c["delta_y"])
            assert (pod.width, pod.height) == (c["width"], c["height"])


def test_default_camera_quirks(S):
    """ASPECT_RATIO = 4/3 is integer 1 -> 640x640 (main.cpp:25); pi is 3.14 (scene.cpp:84)."""
    cam = S.default_camera()
    assert (int(cam.image_width), int(cam.image_height)) == (640, 640)
    assert cam.image_top_left == (-1.0, 0.99764273387050351, -0.99764273387050351)
    assert cam.u[0][1] == -0.0031225124690782585 and cam.u[1][2] == 0.0031225124690782585
    assert S.default_camera(1920, 16.0 / 9.0).image_height == 1080 and S.default_camera(7680, 16.0 / 9.0).image_height == 4320


def test_material_constructor_order_and_default_mat(S):
    m = S.Material((0, 1, 0), 0.5)
    assert (m.metallic, m.ambient, m.diffuse, m.specular, m.specular_exponent) == (.5, .1, .9, .4, 50)
    d = S.default_mat()     # DEFAULT_MAT quirk, scene.h:3
    assert (d.metallic, d.ambient, d.diffuse, d.specular, d.specular_exponent) == (.9, .9, .3, 30, 50)
    pod = S.Sphere().pod()
    assert pod.mat.metallic == .9 and pod.mat.specular == 30 and pod.a == 1.0


@pytest.mark.parametrize("height,band,ranks", [(4320, 4, 8), (4320, 4, 2), (2160, 4, 8), (54, 4, 3), (7, 16, 4), (1, 1, 1), (10, 3, 4)])
def test_band_map_partitions_rows(lib, renderer_mod, height, band, ranks):
    seen = []
    for r in range(ranks):
        rows = renderer_mod.global_rows(height, band, ranks, r)
        assert len(rows) == lib.rtx_local_rows(height, band, ranks, r)
        assert list(rows) == sorted(rows) and all((g // band) % ranks == r for g in rows)
        seen += list(rows)
        assert lib.rtx_global_row(len(rows), height, band, ranks, r) == -1
    assert sorted(seen) == list(range(height))
    assert lib.rtx_local_rows(height, band, ranks, ranks) == 0 and lib.rtx_local_rows(0, band, ranks, 0) == 0


def test_flythrough_cameras_keep_focal_length_one(S):
    """focal length must stay 1 because |d| leaks into sphere hit points (row I)."""
    cams = S.flythrough_cameras(256, 96, 16.0 / 9.0)
    assert len(cams) == 256
    for c in cams[::16]:
        assert abs(c.focal_length - 1.0) < 1e-12 and int(c.image_height) == 54


def test_status_strings(lib):
    assert lib.rtx_status_string(0) == b"ok" and b"invalid" in lib.rtx_status_string(1)


# ---- camera moves (scene.cpp:108-165): golden walks generated from the unmodified reference ---------------------------

def _walks():
    return load_json("camera_walks.json")["walks"]


def _camera_of(S, w):
    cam = S.Camera()
    c = w["camera"]
    cam.position, cam.lookat, cam.vup = fh3(c["position"]), fh3(c["lookat"]), fh3(c["vup"])
    cam.vfov, cam.aspect_ratio, cam.image_width = c["vfov"], float.fromhex(c["aspect_ratio"]), c["image_width"]
    return cam


def _same(got, exp):
    got, exp = np.asarray(got, np.float64), np.asarray(exp, np.float64)
    return np.array_equal(got.view(np.uint64), exp.view(np.uint64)) or np.array_equal(got, exp)   # (-0.0 == 0.0 is fine)


@pytest.mark.parametrize("idx", range(4))
def test_camera_walk_python_mirror_matches_reference(S, idx):
    """scene.Camera.forward/backward/left/right/rotate_*: bit for bit the reference's position / direction / vup after
    every step of the golden walks; init() is not re-run, so image_top_left stays that of init()."""
    w = _walks()[idx]
    cam = _camera_of(S, w)
    cam.init()
    top_left = tuple(cam.image_top_left)
    for (op, arg), state in zip(w["steps"], w["states"]):
        if op in "wsad":
            {"w": cam.forward, "s": cam.backward, "a": cam.left, "d": cam.right}[op]()
        elif op == "y":
            cam.rotate_left_right(arg)
        else:
            cam.rotate_up_down(arg)
        exp = [fh3(v) for v in state]
        assert _same([cam.position, cam.direction, cam.vup], exp), (op, arg)
    pod = cam.pod()
    assert _same([pod.position.x, pod.position.y, pod.position.z], fh3(w["final_pod"]["position"]))
    assert tuple(cam.image_top_left) == top_left and _same(top_left, fh3(w["final_pod"]["image_top_left"]))
    assert (pod.width, pod.height) == (w["final_pod"]["width"], w["final_pod"]["height"])


@pytest.mark.parametrize("idx", range(4))
def test_camera_walk_oracle_port_matches_reference(port, S, idx):
    w = _walks()[idx]
    states, pod = port.camera_walk(_camera_of(S, w), [(op, arg) for op, arg in w["steps"]])
    assert _same(states, [[fh3(v) for v in st] for st in w["states"]])
    assert _same([pod.position.x, pod.position.y, pod.position.z], fh3(w["final_pod"]["position"]))


@pytest.mark.parametrize("idx", range(4))
def test_camera_walk_cpp_facade_matches_reference(pkg, idx):
    """host/rtx_scene.hpp's Camera through examples/camera_walk.cpp (host code only: runs without a GPU)."""
    exe = os.path.join(os.path.dirname(os.path.abspath(pkg.__file__)), "rtx_camera_walk")
    if not os.path.exists(exe):
        subprocess.check_call([os.path.join(os.path.dirname(exe), "build.sh")])
    w = _walks()[idx]
    c = w["camera"]
    argv = [exe] + c["position"] + c["lookat"] + c["vup"] + [repr(float(c["vfov"])), c["aspect_ratio"], repr(float(c["image_width"]))]
    argv += [op if op in "wsad" else "%s:%s" % (op, float(arg).hex()) for op, arg in w["steps"]]
    out = subprocess.run(argv, capture_output=True, text=True, check=True).stdout.split("\n")
    got = [[float.fromhex(x) for x in line.split()] for line in out if line.strip()]
    exp = [[float.fromhex(x) for vec in st for x in vec] for st in w["states"]]
    assert _same(got, exp)


def test_camera_walk_reference_build_agrees_with_golden(ref, S):
    """Where oracle/_ref exists: the golden file is what the unmodified reference computes now."""
    for w in _walks():
        states, _ = ref.camera_walk(_camera_of(S, w), [(op, arg) for op, arg in w["steps"]])
        assert _same(states, [[fh3(v) for v in st] for st in w["states"]])


def test_hot_loop_canary(renderer_mod):
    """The trace kernel's hot loop in the BUILT library: 84 FFMA2 (12 entries x 2 chains x 7) and nothing ptxas should not have
    put there. An unrelated edit once made it re-load kernel parameters inside the loop (184 -> 220 instructions, +2 % frame
    time); tools/sass_ffma2.py is the check, this test runs it."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_ffma2.py"), renderer_mod.LIB_PATH],
                         capture_output=True, text=True, check=True).stdout
    m = re.search(r"hot loop: (\d+) instructions per iteration: \{'FFMA2': (\d+)", out)
    assert m, out
    assert int(m.group(2)) == 84 and int(m.group(1)) <= 190, out
