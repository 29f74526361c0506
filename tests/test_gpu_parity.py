"""GPU suite: the CUDA path, called through the C ABI, against the oracle and the committed golden fixtures.

Bars (BASELINE.json north_star): object-id / hit-mask bit-exact (we additionally require ray counts and RGBA8 to be
identical — the kernel takes every hit decision in the reference's own double arithmetic); radiance relative error
<= 1e-4 with denominator max(|ref|, 1e-3) — asserted here at the much tighter 1e-12, the only divergence being
libm-vs-CUDA pow (<= 2 ulp) and the front-to-back accumulation order of the reflection lerp.
"""
import glob
import hashlib
import sys

import numpy as np
import pytest

from conftest import fh3, load_golden_frame, load_json

pytestmark = pytest.mark.gpu

WANT = ("rgba8", "radiance_f64", "radiance_f32", "object_id", "hit_mask", "ray_count")
RAD_TOL = 1e-12          # asserted; the contract's tolerance is 1e-4
RAD_TOL_CONTRACT = 1e-4


def rel_err(got, exp):
    return np.abs(got - exp) / np.maximum(np.abs(exp), 1e-3)


def check_frame(got, exp, st=None):
    assert np.array_equal(got["object_id"], exp["object_id"])
    assert np.array_equal(got["hit_mask"], exp["hit_mask"])
    assert np.array_equal(got["ray_count"], exp["ray_count"])
    e64 = rel_err(got["radiance_f64"], exp["radiance"])
    assert np.nanmax(e64) < RAD_TOL, np.nanmax(e64)
    e32 = rel_err(got["radiance_f32"].astype(np.float64), exp["radiance"])
    assert np.nanmax(e32) < RAD_TOL_CONTRACT
    lsb = np.abs(unpack(got["rgba8"]).astype(int) - unpack(exp["rgba8"]).astype(int))
    assert lsb.max() <= 1                                  # the contract
    assert np.array_equal(got["rgba8"], exp["rgba8"])      # what we actually deliver
    if st is not None:
        assert st.total_rays == int(exp["ray_count"].astype(np.int64).sum())


def unpack(rgba):
    a = np.asarray(rgba, dtype=np.uint32)
    return np.stack([(a >> 24) & 255, (a >> 16) & 255, (a >> 8) & 255], axis=-1)


def render(gpu, renderer_mod, scene, pod, depth, **kw):
    gpu.set_scene(scene)
    planes, st = gpu.render([pod], renderer_mod.default_params(max_depth=depth, **kw), want=WANT)
    return {k: v[0] for k, v in planes.items()}, st


@pytest.fixture(scope="module")
def syn(S):
    return S.synthetic_scene()


def test_c1_default_frame_matches_reference_hashes(gpu, renderer_mod, S, ob):
    """640x640 default frame: the reference's own golden hashes (ids, ray counts, RGBA8 surface of main())."""
    g = load_json("c1_default_640.json")
    got, st = render(gpu, renderer_mod, S.default_scene(), S.default_camera().pod(), 10)
    assert hashlib.sha256(got["object_id"].tobytes()).hexdigest() == g["object_id_sha256"]
    assert hashlib.sha256(got["ray_count"].tobytes()).hexdigest() == g["ray_count_sha256"]
    assert hashlib.sha256(got["rgba8"].tobytes()).hexdigest() == g["rgba8_sha256"] == g["main_surface_sha256"]
    assert hashlib.sha256(ob.rgb8_bytes(got["rgba8"]).tobytes()).hexdigest() == g["rgb8_sha256"]
    assert st.total_rays == g["total_rays"] == 486688
    for px in g["pixels"]:
        assert rel_err(got["radiance_f64"][px["i"], px["j"]], np.array(fh3(px["rgb"]))).max() < RAD_TOL


@pytest.mark.parametrize("name", ["default_160x90_d8.npz", "synthetic_96x54_d10.npz", "synthetic_384x216_bands_d10.npz",
                                  "flythrough_96x54_k000.npz", "flythrough_96x54_k048.npz",
                                  "flythrough_96x54_k128.npz", "flythrough_96x54_k224.npz"])
def test_golden_frames(gpu, renderer_mod, S, syn, name):
    g = load_golden_frame(name)
    scene = syn if name.startswith("synthetic") else S.default_scene()
    if name.startswith("flythrough"):
        cam = S.flythrough_cameras(256, g["width"], 16.0 / 9.0)[int(name[-7:-4])]
    else:
        cam = S.default_camera(g["width"], 16.0 / 9.0)
    got, _ = render(gpu, renderer_mod, scene, cam.pod(), g["depth"])
    got = {k: v[g["rows"]] for k, v in got.items()}
    check_frame(got, g)


@pytest.mark.parametrize("width,aspect,depth", [(320, 16.0 / 9.0, 8), (97, 1.3, 2), (33, 1.0, 0), (1, 1.0, 10), (640, 1.0, 10)])
def test_default_scene_vs_oracle(gpu, renderer_mod, port, S, width, aspect, depth):
    scene, pod = S.default_scene(), S.default_camera(width, aspect).pod()
    got, st = render(gpu, renderer_mod, scene, pod, depth)
    check_frame(got, port.render(scene, pod, depth), st)


def test_synthetic_10k_vs_oracle(gpu, renderer_mod, port, S, syn):
    """Config C3's scene at a size the oracle finishes in seconds (depth 10, long chains, over-range radiance)."""
    pod = S.default_camera(128, 16.0 / 9.0).pod()
    got, st = render(gpu, renderer_mod, syn, pod, 10)
    exp = port.render(syn, pod, 10)
    check_frame(got, exp, st)
    assert exp["ray_count"].max() == 11                              # chains reach the cap
    assert np.nanmax(exp["radiance"]) > 1.0 and st.over_range_pixels > 0   # the wrap of main.cpp:345 is exercised
    over = (exp["radiance"] * 255 >= 256).any(axis=-1) | (exp["radiance"] < 0).any(axis=-1)
    assert st.over_range_pixels == int(over.sum())
    assert abs(st.max_luminance - exp["radiance"].mean(axis=-1).max()) < 1e-12


def test_synthetic_other_viewpoints(gpu, renderer_mod, port, S, syn):
    """Cameras inside the sphere cloud: origins inside spheres, back faces, grazing hits."""
    for pos, look in [((30.0, 0.0, 8.0), (29.0, 0.3, 8.1)), ((62.0, 30.0, 20.0), (63.0, 31.0, 20.5)), ((10.0, -20.0, -5.0), (10.0, -21.0, -5.0))]:
        cam = S.Camera()
        cam.aspect_ratio, cam.image_width, cam.vfov = 16.0 / 9.0, 64, 90
        cam.position, cam.lookat, cam.vup = pos, look, (0, 0, -1)
        pod = cam.pod()
        got, st = render(gpu, renderer_mod, syn, pod, 10)
        check_frame(got, port.render(syn, pod, 10), st)


def test_mixed_scene_order_and_ties(gpu, renderer_mod, port, S):
    """Walls and spheres interleaved in scene order; duplicated objects give exact distance ties, where the
    lowest scene index must win (strict '<' in main.cpp:77)."""
    M = S.Material
    scene = [
        S.Wall(M((.2, .3, .9), .3), (3.0, 2, 0), (0, -1, 0), 1, 1),
        S.Sphere(M((.9, .2, .1), .6), (2.0, 0.3, 0.2), .4),
        S.Sphere(M((.1, .9, .1), .2), (2.0, 0.3, 0.2), .4),            # exact duplicate of id 1
        S.Wall(M((.9, .9, .1), .5), (3.0, -3, 0), (0, 1, 0), 2, 2),
        S.Wall(M((.1, .9, .9), .5), (3.0, -3, 0), (0, 1, 0), 2, 2),    # exact duplicate of id 3
        S.Sphere(M((.5, .5, .5), .7), (4.0, -1.0, 1.0), 1.0),
        S.Sphere(),                                                    # DEFAULT_MAT, unit sphere around the camera: never hit from inside
        S.Wall(M((1, 1, 1)), (1, 0, 0), (0, 0, 1), 1, 1),               # normal || z: NaN basis, never hit (scene.cpp:18)
    ]
    pod = S.default_camera(96, 1.0).pod()
    got, st = render(gpu, renderer_mod, scene, pod, 6)
    exp = port.render(scene, pod, 6)
    check_frame(got, exp, st)
    ids = set(np.unique(exp["object_id"]))
    assert 1 in ids and 2 not in ids and 3 in ids and 4 not in ids and 6 not in ids and 7 not in ids


def test_empty_scene_and_walls_only_and_spheres_only(gpu, renderer_mod, port, S):
    pod = S.default_camera(48, 1.0).pod()
    for scene in ([], S.default_scene()[1:], S.default_scene()[:1], S.synthetic_scene(7, 0, seed=3), S.synthetic_scene(0, 5, seed=4)):
        got, st = render(gpu, renderer_mod, scene, pod, 4)
        check_frame(got, port.render(scene, pod, 4), st)


def test_params_are_honoured(gpu, renderer_mod, port, S):
    scene, pod = S.default_scene(), S.default_camera(80, 1.0).pod()
    kw = dict(light_pos=(0.5, -1.0, 2.0), ground_color=(.3, .2, .1), sky_low=(.9, .8, .7), sky_high=(.1, .2, .3),
              reflect_offset=.001, sky_exponent=0.5)
    p = port.default_params()
    p.max_depth = 5
    for k, v in kw.items():
        setattr(p, k, type(getattr(p, k))(*v) if isinstance(v, tuple) else v)
    got, st = render(gpu, renderer_mod, scene, pod, 5, **kw)
    check_frame(got, port.render(scene, pod, params=p), st)


def test_multi_frame_batch_equals_single_frames(gpu, renderer_mod, port, S):
    """Config C5's shape: several cameras in one launch."""
    scene = S.default_scene()
    cams = S.flythrough_cameras(256, 64, 16.0 / 9.0)
    pick = [0, 48, 100, 128, 224]
    gpu.set_scene(scene)
    planes, st = gpu.render([cams[k].pod() for k in pick], renderer_mod.default_params(), want=WANT)
    total = 0
    for f, k in enumerate(pick):
        exp = port.render(scene, cams[k].pod(), 10)
        check_frame({n: v[f] for n, v in planes.items()}, exp)
        total += exp["total_rays"]
    assert st.total_rays == total


@pytest.mark.parametrize("height_w,band,ranks", [((54, 96), 4, 2), ((54, 96), 4, 8), ((50, 64), 3, 4), ((7, 16), 16, 4)])
def test_row_band_sharding_reassembles_the_frame(gpu, renderer_mod, port, S, syn, height_w, band, ranks):
    """Config C4's shape: every rank renders its cyclic bands; the union is the full frame."""
    H, W = height_w
    cam = S.default_camera(W, W / H)
    pod = cam.pod()
    assert pod.height == H
    scene = syn[:400] + syn[10000:10016]
    exp = port.render(scene, pod, 10)
    gpu.set_scene(scene)
    seen = np.zeros(H, bool)
    total = 0
    for r in range(ranks):
        rows = renderer_mod.global_rows(H, band, ranks, r)
        if len(rows) == 0:
            continue
        planes, st = gpu.render([pod], renderer_mod.default_params(band_rows=band, n_ranks=ranks, rank=r), want=WANT)
        assert planes["rgba8"].shape == (1, len(rows), W)
        check_frame({n: v[0] for n, v in planes.items()}, {k: exp[k][rows] for k in ("radiance", "rgba8", "object_id", "hit_mask", "ray_count")})
        seen[rows] = True
        total += st.total_rays
    assert seen.all() and total == exp["total_rays"]


def test_unfused_quantise_equals_fused(gpu, renderer_mod, S, syn):
    pod = S.default_camera(96, 16.0 / 9.0).pod()
    gpu.set_scene(syn)
    a, sa = gpu.render([pod], renderer_mod.default_params(fuse_quantise=1), want=("rgba8", "radiance_f64"))
    b, sb = gpu.render([pod], renderer_mod.default_params(fuse_quantise=0), want=("rgba8",))
    assert np.array_equal(a["rgba8"], b["rgba8"])
    assert sb.launches == 2 and sa.launches == 1 and sb.surface_update_ms > 0
    assert sa.over_range_pixels == sb.over_range_pixels and sa.max_luminance == sb.max_luminance


def test_standalone_quantise_kernel(gpu, port):
    """main.cpp:338-347 as its own kernel: golden vectors from the reference, wrap semantics, ragged sizes."""
    q = load_json("kat.json")["quantise"]
    rgb = np.array([float.fromhex(x) for x in q["rgb"]]).reshape(-1, 3)
    assert list(gpu.quantise(rgb)) == q["rgba8"]
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 4, 5, 1023, 4096, 100003):
        x = rng.uniform(-0.5, 1.6, size=(n, 3))
        assert np.array_equal(gpu.quantise(x), port.quantise(x))
        x32 = x.astype(np.float32)
        assert np.array_equal(gpu.quantise(x32), port.quantise_mode(x32, 0))
        assert np.array_equal(gpu.quantise(x, 1), port.quantise_mode(x, 1))
        over = ((x * 255 >= 256) | (x * 255 < 0)).any(axis=-1).sum()
        gpu.quantise(x)
        assert gpu.last_stats.over_range_pixels == over
    assert gpu.quantise(np.zeros((0, 3))).size == 0


def test_rt_scene_shaped_call(renderer_mod, port, S):
    """The drop-in call with the reference's signature shape (main.cpp:124-125, call site main.cpp:329)."""
    cam = S.default_camera(120, 1.0)
    u = cam.init()
    scene = S.default_scene()
    frame_buffer = np.zeros((int(cam.image_height), int(cam.image_width), 3))
    renderer_mod.rt_scene(u, scene, cam, frame_buffer)
    exp = port.render(scene, cam.pod(), 10)
    assert rel_err(frame_buffer, exp["radiance"]).max() < RAD_TOL


def test_error_paths(gpu, renderer_mod, pkg, S):
    a = pkg.abi
    fresh = renderer_mod.Renderer(0)
    with pytest.raises(renderer_mod.RtxError) as e:
        fresh.render([S.default_camera(8, 1.0).pod()])
    assert e.value.status == a.RTX_ERR_NO_SCENE
    fresh.set_scene(S.default_scene())
    with pytest.raises(renderer_mod.RtxError) as e:
        fresh.render([S.default_camera(8, 1.0).pod()], renderer_mod.default_params(max_depth=300))
    assert e.value.status == a.RTX_ERR_INVALID
    with pytest.raises(renderer_mod.RtxError):
        fresh.render([S.default_camera(8, 1.0).pod(), S.default_camera(9, 1.0).pod()])
    with pytest.raises(renderer_mod.RtxError):
        fresh.render([S.default_camera(8, 1.0).pod()], renderer_mod.default_params(n_ranks=2, rank=2))
    bad = S.Sphere()
    bad.kind = 7
    with pytest.raises(renderer_mod.RtxError):
        fresh.set_scene([bad])
    fresh.close()


def test_ray_far_outside_scene_bound_falls_back_to_exact(gpu, renderer_mod, port, S):
    """Primary-ray overshoot can start a bounce beyond the FP32 screen's assumed origin
    bound: those lanes must fall back to exact tests. A wide FOV with tiny spheres far away provokes it."""
    M = S.Material
    scene = [S.Sphere(M((.9, .9, .9), .9), (40.0, 39.0, 39.0), 1.0), S.Sphere(M((.2, .9, .2), .9), (41.0, 35.0, 41.0), 1.5),
             S.Sphere(M((.9, .2, .2), .9), (38.0, 42.0, 36.0), 2.0), S.Wall(M((.3, .3, .9), .5), (60, -30, -30), (-1, 0, 0), 60, 60)]
    cam = S.Camera()
    cam.aspect_ratio, cam.image_width, cam.vfov = 1.0, 200, 100
    cam.position, cam.lookat, cam.vup = (0, 0, 0), (-1, -1, -1), (0, 0, -1)
    pod = cam.pod()
    got, st = render(gpu, renderer_mod, scene, pod, 8)
    check_frame(got, port.render(scene, pod, 8), st)


def test_cpp_headless_main_matches_reference_main_loop(tmp_path, port, S, renderer_mod):
    """The C++ host facade (host/rtx_scene.hpp) driving the C ABI like the reference's main loop (main.cpp:250-375):
    frame 0 at the start position, frame 1 after a 'w' key (Camera::forward, init() NOT re-run, as the reference)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(renderer_mod.LIB_PATH), "rtx_headless")
    assert os.path.exists(exe), "build() must produce rtx_headless"
    raw, ppm, png = tmp_path / "f.rgba", tmp_path / "f.ppm", tmp_path / "f.png"
    out = subprocess.run([exe, "--width", "320", "--frames", "2", "--keys", "xw", "--raw", str(raw), "--out", str(ppm), "--png", str(png)],
                         check=True, capture_output=True, text=True).stdout
    assert "microseconds for average raytracing" in out and "milliseconds for surface average update" in out
    cam = S.default_camera(320, 1.0)
    pod = cam.pod()
    pod.position = type(pod.position)(0.1, 0.0, 0.0)        # forward(): position + normalize(direction) * 0.1, direction = (1,0,0)
    exp = port.render(S.default_scene(), pod, 10, want=("rgba8",))["rgba8"]
    got = np.fromfile(raw, dtype=np.uint32).reshape(320, 320)
    assert np.array_equal(got, exp)
    data = open(ppm, "rb").read()
    assert data.startswith(b"P6\n320 320\n255\n") and len(data) == 15 + 320 * 320 * 3
    # the PNG (stored deflate) decodes to the same pixels
    import struct
    import zlib
    blob = open(png, "rb").read()
    assert blob[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat = 8, b""
    while pos < len(blob):
        n, tag = struct.unpack(">I4s", blob[pos:pos + 8])
        body = blob[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
        if tag == b"IDAT":
            idat += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(320, 1 + 320 * 3)
    assert (rows[:, 0] == 0).all() and rows[:, 1:].tobytes() == data[15:]
    # a longer scripted walk: forward, right, backward (Camera::forward/right/backward, scene.cpp:121-135), init() never re-run
    raw2 = tmp_path / "g.rgba"
    subprocess.run([exe, "--width", "200", "--frames", "4", "--keys", "xwds", "--raw", str(raw2), "--out", ""], check=True, capture_output=True)
    cam2 = S.default_camera(200, 1.0)
    pod2 = cam2.pod()
    pos_x = (0.0 + 0.1) - 0.1                               # forward then backward along direction (1,0,0)
    pod2.position = type(pod2.position)(pos_x, 0.1, 0.0)    # right_vec = normalize(cross(direction, vup)) = (0,1,0)
    exp2 = port.render(S.default_scene(), pod2, 10, want=("rgba8",))["rgba8"]
    assert np.array_equal(np.fromfile(raw2, dtype=np.uint32).reshape(200, 200), exp2)
    # moves interleaved with the mouse-look rotations (scene.cpp:137-165; keys j/l/i/k = rotate_left_right(+-0.05) /
    # rotate_up_down(+-0.05)): a rotation changes right_vec and so the later a/d moves; expected camera from the oracle's walk
    raw3 = tmp_path / "h.rgba"
    subprocess.run([exe, "--width", "160", "--frames", "9", "--keys", "xjjdikkaw", "--raw", str(raw3), "--out", ""], check=True, capture_output=True)
    steps = [("y", 0.05), ("y", 0.05), ("d", 0), ("p", 0.05), ("p", -0.05), ("p", -0.05), ("a", 0), ("w", 0)]
    states, pod3 = port.camera_walk(S.default_camera(160, 1.0), steps)
    assert abs(states[-1][0][1]) > 1e-4                     # the yaw really moved the camera sideways differently
    exp3 = port.render(S.default_scene(), pod3, 10, want=("rgba8",))["rgba8"]
    assert np.array_equal(np.fromfile(raw3, dtype=np.uint32).reshape(160, 160), exp3)


def test_scene_larger_than_shared_memory_streams_tiles(gpu, renderer_mod, port, S):
    """More entries than fit the 227 KB shared-memory tile (~13.8k): the kernel streams tiles through the same
    buffer in a CTA-synchronous loop (template STREAM). Same parity bar."""
    scene = S.synthetic_scene(16000, 40, seed=0x51)
    pod = S.default_camera(48, 16.0 / 9.0).pod()
    got, st = render(gpu, renderer_mod, scene, pod, 4)
    check_frame(got, port.render(scene, pod, 4), st)
    assert st.sphere_tests == st.total_rays * 16000


@pytest.mark.parametrize("flag", ["-DRTX_MBOX_CAP=1", "-DRTX_QUEUE_CAP=1"])
def test_overflow_paths_of_the_trace_kernel(tmp_path, renderer_mod, port, S, flag):
    """Two fixed-size buffers of the trace kernel have overflow paths that ordinary scenes rarely take:
    the cooperative drain's per-warp mailbox (overflow: the segment is redone the ordinary way) and a chain's queue of
    screen survivors (full: flushed early, in the middle of the scan). Build the library with 1-slot buffers to force
    those paths everywhere and check parity."""
    import os
    import subprocess
    pkg_dir = os.path.dirname(os.path.abspath(renderer_mod.__file__))
    root = os.path.dirname(pkg_dir)
    sources = sorted(glob.glob(os.path.join(pkg_dir, "csrc", "*.cu")))      # the same list, in the same order, as build.sh
    lib = os.path.join(pkg_dir, "test_builds", "librtx_b200_%s.so" % flag[2:])      # prebuilt by build.sh
    headers = [os.path.join(pkg_dir, "csrc", "rtx_device.cuh"), os.path.join(root, "include", "rtx_b200.h")]
    digest = hashlib.sha256(b"".join(open(f, "rb").read() for f in sources + headers)).hexdigest()
    stamp = os.path.join(pkg_dir, "test_builds", "SOURCES.sha256")
    if not (os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == digest):   # stale or absent
        lib = str(tmp_path / "librtx_b200_small_buffers.so")
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off",
                        flag, "-I" + os.path.join(root, "include"), "-I" + os.path.join(pkg_dir, "csrc"), "-shared"] + sources + ["-o", lib],
                       check=True)
    code = r"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, %r)
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
S = importlib.import_module("ray-tracer-from-scratch_b200").scene
from oracle import binding as ob
port = ob.load_port()
scene = S.synthetic_scene(3000, 24, seed=9)
pod = S.default_camera(96, 16.0 / 9.0).pod()
r = R.Renderer(0)
r.set_scene(scene)
got, st = r.render([pod], R.default_params(max_depth=10), want=("rgba8", "object_id", "ray_count"))
exp = port.render(scene, pod, 10)
assert np.array_equal(got["rgba8"][0], exp["rgba8"]) and np.array_equal(got["object_id"][0], exp["object_id"])
assert np.array_equal(got["ray_count"][0], exp["ray_count"]) and st.total_rays == exp["total_rays"]
print("ok")
""" % root
    env = dict(os.environ, RTX_B200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


@pytest.mark.parametrize("n_row,n_other", [(2400, 40), (3000, 11000)])
def test_ray_threading_thousands_of_spheres_flushes_the_survivor_queue(gpu, renderer_mod, port, S, n_row, n_other):
    """Thousands of nested spheres around the central line of sight: the central lanes find screen survivors in every
    hot-loop iteration, far more than a chain's queue holds, so the early flush runs many times per ray (second
    case: scene larger than shared memory, i.e. with tile streaming)."""
    rng = np.random.default_rng(n_row)
    mat = S.Material((0.9, 0.5, 0.2), 0.3)
    xs = [60.0 - 58.0 * i / n_row for i in range(n_row)]          # a cone of nested spheres around the view axis
    scene = [S.Sphere(mat, (x, 1e-3 * np.sin(i), 1e-3 * np.cos(i)), 0.15 * x * (0.8 + 0.1 * (i % 3))) for i, x in enumerate(xs)]
    for i in range(n_other):
        c = rng.uniform((4, -32, -8), (64, 32, 24))
        scene.append(S.Sphere(S.Material(tuple(rng.uniform(0.1, 1, 3)), float(rng.uniform(0, 0.8))), tuple(c), float(rng.uniform(0.1, 0.6))))
    scene.append(S.Wall(S.Material((0.2, 0.3, 0.9), 0.4), (70.0, -5.0, -5.0), (-1.0, 0.0, 0.0), 10.0, 10.0))
    pod = S.default_camera(49, 16.0 / 9.0).pod()                  # odd size: the centre pixel looks exactly along the axis
    got, st = render(gpu, renderer_mod, scene, pod, 3)
    check_frame(got, port.render(scene, pod, 3), st)
    assert (got["object_id"] >= 0).sum() > 8


def test_fused_frame_output_places_pixels_globally(gpu, renderer_mod, port, S, syn, pkg):
    """rtx_outputs.frame_rgba8 (the fused multi-GPU gather): every rank's kernel stores its pixels at their GLOBAL
    position of one shared row-major frame. Emulated on one GPU: 4 ranks x cyclic 3-row bands into one buffer
    allocated with rtx_buffer_alloc, and 5 frames sharded over 2 ranks with frame_offset/frame_stride."""
    import ctypes as C
    import torch
    a = pkg.abi
    scene = syn[:300] + syn[10000:10010]
    pod = S.default_camera(64, 64 / 50).pod()
    H, W = pod.height, pod.width
    assert H == 50
    exp = port.render(scene, pod, 8, want=("rgba8",))["rgba8"]
    gpu.set_scene(scene)
    ptr = gpu.buffer_alloc(H * W * 4)
    try:
        handle = gpu.buffer_export(ptr)
        assert len(handle) == 64
        view = torch.as_tensor(type("P", (), {"__cuda_array_interface__": {"shape": (H, W), "typestr": "<i4", "data": (ptr, False), "version": 3}})(),
                               device="cuda:0")
        view.fill_(-1)
        for r in range(4):
            o = a.Outputs()
            o.memory, o.frame_rgba8 = a.RTX_MEM_DEVICE, ptr
            gpu.render_raw([pod], renderer_mod.default_params(max_depth=8, band_rows=3, n_ranks=4, rank=r), o)
        assert np.array_equal(view.cpu().numpy().view(np.uint32), exp)
    finally:
        gpu.buffer_free(ptr)
    # frames sharded over 2 ranks
    cams = [c.pod() for c in S.flythrough_cameras(256, 48, 16.0 / 9.0)[::60]][:5]
    h, w = cams[0].height, cams[0].width
    gpu.set_scene(S.default_scene())
    ptr = gpu.buffer_alloc(5 * h * w * 4)
    try:
        view = torch.as_tensor(type("P", (), {"__cuda_array_interface__": {"shape": (5, h, w), "typestr": "<i4", "data": (ptr, False), "version": 3}})(),
                               device="cuda:0")
        view.fill_(-1)
        for r in range(2):
            o = a.Outputs()
            o.memory, o.frame_rgba8 = a.RTX_MEM_DEVICE, ptr
            gpu.render_raw(cams[r::2], renderer_mod.default_params(frame_offset=r, frame_stride=2), o)
        got = view.cpu().numpy().view(np.uint32)
        for f, pod_f in enumerate(cams):
            assert np.array_equal(got[f], port.render(S.default_scene(), pod_f, 10, want=("rgba8",))["rgba8"]), f
    finally:
        gpu.buffer_free(ptr)


def test_non_finite_objects_and_extreme_depth(gpu, renderer_mod, port, S):
    """Objects with NaN / inf geometry are never hit by the reference (comparisons with NaN are false) but must not
    disturb the others: the FP32 screen degrades to "everything goes to the exact test". Depth cap 254 = the
    largest the uint8 ray-count plane can carry."""
    M = S.Material
    nan, inf = float("nan"), float("inf")
    scene = S.default_scene() + [
        S.Sphere(M((1, 0, 0)), (nan, 0, 0), .5), S.Sphere(M((1, 0, 0)), (2, 0, 0), nan),
        S.Wall(M((1, 1, 0)), (3, nan, 0), (0, 1, 0), 1, 1), S.Wall(M((1, 1, 0)), (2.5, 1, 0), (0, 0, 0), 1, 1),   # zero normal -> NaN
        S.Sphere(M((0, 0, 1), .9), (2.0, 1.0, 0.5), 0.25), S.Sphere(M((0, 1, 1), .9), (2.0, -1.0, 0.5), -0.25),   # negative radius = same sphere
    ]
    pod = S.default_camera(64, 1.0).pod()
    got, st = render(gpu, renderer_mod, scene, pod, 254)
    exp = port.render(scene, pod, 254)
    check_frame(got, exp, st)
    assert set(np.unique(exp["object_id"])) >= {-1, 0, 7, 8} and not ({3, 4, 5, 6} & set(np.unique(exp["object_id"])))
    scene_inf = S.default_scene() + [S.Sphere(M((1, 0, 0)), (inf, 0, 0), .5)]
    got, st = render(gpu, renderer_mod, scene_inf, pod, 4)
    check_frame(got, port.render(scene_inf, pod, 4), st)


def test_one_row_and_one_column_frames(gpu, renderer_mod, port, S):
    scene = S.default_scene()
    for width, aspect in ((200, 200.0), (1, 0.01), (7, 7.0 / 3)):
        pod = S.default_camera(width, aspect).pod()
        got, st = render(gpu, renderer_mod, scene, pod, 10)
        assert got["rgba8"].shape == (pod.height, pod.width)
        check_frame(got, port.render(scene, pod, 10), st)


def test_small_scene_kernel_boundary(gpu, renderer_mod, port, S):
    """Scenes of up to 16 objects run the compact register-resident kernel (every object tested exactly), larger ones
    the screened kernel: both sides of the switch must match the oracle, for spheres-only, walls-only and mixed."""
    pod = S.default_camera(72, 16.0 / 9.0).pod()
    for n_s, n_w in ((16, 0), (17, 0), (0, 16), (0, 17), (9, 7), (9, 8), (1, 0), (0, 1)):
        scene = S.synthetic_scene(n_s, n_w, seed=100 + n_s * 31 + n_w)
        for g in scene:                      # pull the random objects in front of the default camera
            if g.kind == 0:
                g.center = (2.0 + g.center[0] / 16.0, g.center[1] / 12.0, g.center[2] / 12.0)
            else:
                g.position = (2.0 + g.position[0] / 16.0, g.position[1] / 12.0, g.position[2] / 12.0)
        got, st = render(gpu, renderer_mod, scene, pod, 10)
        exp = port.render(scene, pod, 10)
        check_frame(got, exp, st)
        assert exp["hit_mask"].any(), (n_s, n_w)


def test_small_scene_host_frames_are_traced_in_ranges(gpu, renderer_mod, port, S):
    """A small scene rendered into HOST memory is traced as four consecutive pixel ranges whose read-back overlaps the next
    range's kernel (rtx_render, `ranged`): same pixels as one launch. Cases: one frame, a pixel count that is not a
    multiple of four, a batch of frames (ranges cut across frame boundaries), and the device-memory path (one launch)."""
    scene = S.default_scene()
    gpu.set_scene(scene)
    for width, aspect, depth in ((640, 4.0 / 3.0, 6), (643, 2.0, 3)):
        pod = S.default_camera(width, aspect).pod()
        assert pod.width * pod.height >= 1 << 17
        planes, st = gpu.render([pod], renderer_mod.default_params(max_depth=depth), want=WANT)
        check_frame({k: v[0] for k, v in planes.items()}, port.render(scene, pod, depth), st)
        assert st.launches == 4
    pods = [c.pod() for c in S.flythrough_cameras(3, 321, 4.0 / 3.0)]
    planes, st = gpu.render(pods, renderer_mod.default_params(max_depth=5), want=WANT)
    assert st.launches == 4
    total = 0
    for f, pod in enumerate(pods):
        exp = port.render(scene, pod, 5)
        check_frame({k: v[f] for k, v in planes.items()}, exp)
        total += exp["total_rays"]
    assert st.total_rays == total
    small, st1 = gpu.render([S.default_camera(160, 1.0).pod()], renderer_mod.default_params(max_depth=5), want=("rgba8",))
    assert st1.launches == 1                                    # below the threshold: one launch, one copy


def test_grazing_spheres_never_lose_a_hit(gpu, renderer_mod, port, S):
    """Adversarial input for the conservative FP32 screen: spheres TANGENT to camera rays (distance from the centre
    to the ray = r * (1 + delta), |delta| from 1e-9 to 1e-5, both signs), far from the origin, so that hit / miss is
    decided in the last bits of the double discriminant. If the screen's error bound were too tight, true hits would
    be dropped and ids would differ from the reference arithmetic."""
    import math
    import random
    rng = random.Random(0xE)
    cam = S.default_camera(96, 1.0)
    pod = cam.pod()
    tl, dx, dy = pod.image_top_left.tuple(), pod.delta_x.tuple(), pod.delta_y.tuple()
    scene = []
    deltas = [0.0] + [s * 10.0 ** -k for k in range(5, 10) for s in (1, -1)]
    for n in range(330):
        i, j = rng.randrange(96), rng.randrange(96)
        centre = [tl[k] + dx[k] * j + dy[k] * i for k in range(3)]
        d = [-c for c in centre]                                   # position (0,0,0) - pixel centre
        L = math.sqrt(sum(x * x for x in d))
        dh = [x / L for x in d]
        a = [rng.uniform(-1, 1) for _ in range(3)]
        dot = sum(a[k] * dh[k] for k in range(3))
        nrm = [a[k] - dot * dh[k] for k in range(3)]
        nl = math.sqrt(sum(x * x for x in nrm))
        nrm = [x / nl for x in nrm]
        t = rng.uniform(20.0, 900.0)
        r = rng.uniform(0.05, 3.0)
        off = r * (1.0 + deltas[n % len(deltas)])
        c = tuple(t * dh[k] + off * nrm[k] for k in range(3))
        scene.append(S.Sphere(S.Material((rng.uniform(.1, 1), rng.uniform(.1, 1), rng.uniform(.1, 1)), rng.uniform(0, .8)), c, r))
    got, st = render(gpu, renderer_mod, scene, pod, 6)
    exp = port.render(scene, pod, 6)
    check_frame(got, exp, st)
    assert (exp["object_id"] >= 0).sum() > 50          # the construction does produce hits (and near-misses)


def test_mapped_host_surface_equals_staged_path(gpu, renderer_mod, port, S):
    """RTX_MEM_HOST_MAPPED: the kernels store straight into a pinned, mapped host surface (the SDL surface->pixels case,
    main.cpp:193,344) — no staging buffer, no copy. Same words as the staged RTX_MEM_HOST path, for the small-scene and
    the big kernel, all planes."""
    import ctypes as C
    from conftest import ROOT  # noqa: F401
    abi = renderer_mod.abi
    for scene, pod, depth in ((S.default_scene(), S.default_camera(200, 16.0 / 9.0).pod(), 8),
                              (S.synthetic_scene(900, 12), S.default_camera(120, 16.0 / 9.0).pod(), 6)):
        gpu.set_scene(scene)
        staged, st = gpu.render([pod], renderer_mod.default_params(max_depth=depth), want=WANT)
        n = pod.width * pod.height
        sizes = {"rgba8": 4, "radiance_f64": 24, "radiance_f32": 12, "object_id": 4, "hit_mask": 1, "ray_count": 1}
        ptrs = {k: gpu.host_alloc(n * b) for k, b in sizes.items()}
        try:
            o = abi.Outputs()
            o.memory = abi.RTX_MEM_HOST_MAPPED
            for k in sizes:
                setattr(o, k, ptrs[k])
            st2 = gpu.render_raw([pod], renderer_mod.default_params(max_depth=depth), o)
            assert st2.total_rays == st.total_rays and st2.d2h_ms < 0.05          # nothing is copied after the kernel
            for k, (field, dtype, tail) in renderer_mod._PLANES.items():
                if k not in sizes:
                    continue
                got = np.ctypeslib.as_array(C.cast(ptrs[k], C.POINTER(C.c_uint8)), shape=(n * sizes[k],)).view(dtype).reshape(staged[k][0].shape)
                assert np.array_equal(got, staged[k][0], equal_nan=True) if got.dtype.kind == "f" else np.array_equal(got, staged[k][0]), k
        finally:
            for p in ptrs.values():
                gpu.host_free(p)
    # pageable memory is refused in this mode (it cannot be mapped), with a message that says what to use
    buf = np.zeros(16, np.uint32)
    o = abi.Outputs()
    o.memory, o.rgba8 = abi.RTX_MEM_HOST_MAPPED, buf.ctypes.data
    with pytest.raises(renderer_mod.RtxError) as e:
        gpu.render_raw([S.default_camera(4, 1.0).pod()], renderer_mod.default_params(), o)
    assert e.value.status == abi.RTX_ERR_INVALID and "rtx_host_alloc" in str(e.value)


def test_async_frames_in_flight_equal_synchronous_frames(gpu, renderer_mod, port, S):
    """rtx_render_async / rtx_wait: RTX_MAX_IN_FLIGHT frames in flight, host and device outputs, stats returned in order; one
    call more without a wait is refused; a synchronous call drains what is in flight."""
    import torch
    abi = renderer_mod.abi
    scene = S.default_scene()
    gpu.set_scene(scene)
    cams = [c.pod() for c in S.flythrough_cameras(256, 160, 16.0 / 9.0)[::37]]
    p = renderer_mod.default_params(max_depth=10)
    exp = [port.render(scene, c, 10, want=("rgba8", "ray_count")) for c in cams]
    host = [np.zeros((cams[0].height, cams[0].width), np.uint32) for _ in cams]
    outs = []
    for h in host:
        o = abi.Outputs()
        o.memory, o.rgba8 = abi.RTX_MEM_HOST, h.ctypes.data
        outs.append(o)
    stats = []
    depth = abi.RTX_MAX_IN_FLIGHT
    assert len(cams) > depth
    for k, c in enumerate(cams):
        gpu.render_async([c], p, outs[k])
        if k >= depth - 1:
            stats.append(gpu.wait())
    for _ in range(depth - 1):
        stats.append(gpu.wait())
    for k in range(len(cams)):
        assert np.array_equal(host[k], exp[k]["rgba8"]), k
        assert stats[k].total_rays == int(exp[k]["ray_count"].astype(np.int64).sum()), k
    with pytest.raises(renderer_mod.RtxError):
        gpu.wait()                                             # nothing in flight
    for h in host:
        h[:] = 0
    for k in range(depth):
        gpu.render_async([cams[k]], p, outs[k])
    with pytest.raises(renderer_mod.RtxError):
        gpu.render_async([cams[depth]], p, outs[depth])        # one more call in flight is refused
    dev = torch.zeros((cams[depth].height, cams[depth].width), dtype=torch.int32, device="cuda")
    o = abi.Outputs()
    o.memory, o.rgba8 = abi.RTX_MEM_DEVICE, dev.data_ptr()
    st = gpu.render_raw([cams[depth]], p, o)                   # synchronous: completes the calls in flight first
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), exp[depth]["rgba8"]) and st.total_rays == stats[depth].total_rays
    for k in range(depth):
        assert np.array_equal(host[k], exp[k]["rgba8"]), k
    with pytest.raises(renderer_mod.RtxError):
        gpu.wait()


def test_frame_copy_mode_places_bands_and_frames(gpu, renderer_mod, port, S):
    """RTX_FRAME_COPY: staging + copy-engine transfers to the global position — cyclic bands of one frame (2-D copies,
    ragged last band) and whole frames of a camera path with frame_offset / frame_stride, into pinned host memory."""
    import ctypes as C
    abi = renderer_mod.abi
    scene = S.default_scene()
    gpu.set_scene(scene)
    pod = S.default_camera(100, 100.0 / 54).pod()                     # 54 rows: ragged for 4-row bands over 3 ranks
    exp = port.render(scene, pod, 6, want=("rgba8",))["rgba8"]
    H, W = pod.height, pod.width
    ptr = gpu.host_alloc(H * W * 4)
    try:
        frame = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(H, W))
        frame[:] = 0
        for rank in range(3):
            o = abi.Outputs()
            o.memory, o.frame_mode, o.frame_rgba8 = abi.RTX_MEM_DEVICE, abi.RTX_FRAME_COPY, ptr
            gpu.render_raw([pod], renderer_mod.default_params(max_depth=6, band_rows=4, n_ranks=3, rank=rank), o)
        assert np.array_equal(frame, exp)
    finally:
        gpu.host_free(ptr)
    cams = [c.pod() for c in S.flythrough_cameras(256, 96, 16.0 / 9.0)[::32]]      # 8 frames
    H, W = cams[0].height, cams[0].width
    ptr = gpu.host_alloc(len(cams) * H * W * 4)
    try:
        frames = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(len(cams), H, W))
        frames[:] = 0
        for rank in range(2):                                          # frame f -> rank f % 2, two chunks each, asynchronously
            mine = list(range(rank, len(cams), 2))
            for chunk in (mine[:2], mine[2:]):
                o = abi.Outputs()
                o.memory, o.frame_mode, o.frame_rgba8 = abi.RTX_MEM_DEVICE, abi.RTX_FRAME_COPY, ptr
                gpu.render_async([cams[f] for f in chunk], renderer_mod.default_params(max_depth=10, frame_offset=chunk[0], frame_stride=2), o)
            gpu.wait()
            gpu.wait()
        for f, c in enumerate(cams):
            assert np.array_equal(frames[f], port.render(scene, c, 10, want=("rgba8",))["rgba8"]), f
    finally:
        gpu.host_free(ptr)


def test_device_frame_buffer_alloc_store_read(gpu, renderer_mod, port, S):
    """rtx_buffer_alloc + frame_rgba8 (RTX_FRAME_STORE, one rank) + rtx_buffer_read: the `value` path of bench.py without torch."""
    abi = renderer_mod.abi
    scene = S.default_scene()
    gpu.set_scene(scene)
    pod = S.default_camera(120, 1.5).pod()
    n = pod.width * pod.height
    ptr = gpu.buffer_alloc(n * 4)
    try:
        o = abi.Outputs()
        o.memory, o.frame_rgba8 = abi.RTX_MEM_DEVICE, ptr
        gpu.render_raw([pod], renderer_mod.default_params(max_depth=7), o)
        got = gpu.buffer_read(ptr, np.zeros((pod.height, pod.width), np.uint32))
        assert np.array_equal(got, port.render(scene, pod, 7, want=("rgba8",))["rgba8"])
    finally:
        gpu.buffer_free(ptr)


def test_two_contexts_on_one_device_do_not_lower_each_others_shared_memory_limit(renderer_mod, S):
    """The opt-in to > 48 KB of dynamic shared memory belongs to (kernel, device) of the process, not to a context: a second
    context with a small scene once lowered it, and the first context's next big-scene launch failed (cudaErrorInvalidValue).
    Brute force and grid kernels, big scene / small scene / big scene again."""
    big, small = S.synthetic_scene(6000, 32, seed=3), S.synthetic_scene(40, 4, seed=3)
    pod = S.default_camera(64, 16.0 / 9.0).pod()
    a, b = renderer_mod.Renderer(0), renderer_mod.Renderer(0)
    try:
        a.set_scene(big)
        b.set_scene(small)
        for accel in (0, 1):
            first, _ = a.render([pod], renderer_mod.default_params(max_depth=4, accel=accel), want=("rgba8",))
            b.render([pod], renderer_mod.default_params(max_depth=4, accel=accel), want=("rgba8",))
            again, _ = a.render([pod], renderer_mod.default_params(max_depth=4, accel=accel), want=("rgba8",))
            assert np.array_equal(first["rgba8"], again["rgba8"])
    finally:
        a.close()
        b.close()
