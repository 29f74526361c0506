"""GPU suite, part 3: the EXTENSIONS (RTX_BOX, sun, Reinhard tone map) through the C ABI against their specification in
oracle/oracle.c (there is no reference code for them: tests/test_extensions_oracle.py pins what can be pinned).
Bars: ids / hit masks / ray counts / RGBA8 identical and radiance within 1e-12 for boxes and the sun (the kernel runs the
specification's double arithmetic operation for operation); the tone map within 1 LSB (CUDA's log/exp vs glibc's)."""
import numpy as np
import pytest

from test_gpu_parity import WANT, check_frame, render, unpack

pytestmark = pytest.mark.gpu


def _random_boxes(S, rng, n, lo=(4, -20, -8), hi=(50, 20, 16)):
    out = []
    for _ in range(n):
        mat = S.Material(tuple(rng.uniform(0.1, 1, 3)), float(rng.uniform(0, 0.8)))
        out.append(S.Box(mat, tuple(rng.uniform(lo, hi)), tuple(rng.uniform(0.3, 5.0, 3))))
    return out


def test_boxes_small_scene_kernel(gpu, renderer_mod, port, S):
    """<= 16 screen entries (a box counts six): the compact kernel. Reference default scene + one box resting between the walls."""
    scene = S.default_scene() + [S.Box(S.Material((0.9, 0.2, 0.2), 0.6), (2.5, -1.0, -0.8), (1.0, 1.2, 0.9))]
    for width, aspect, depth in ((160, 16.0 / 9.0, 8), (97, 1.0, 3)):
        pod = S.default_camera(width, aspect).pod()
        got, st = render(gpu, renderer_mod, scene, pod, depth)
        exp = port.render(scene, pod, depth)
        check_frame(got, exp, st)
        assert (exp["object_id"] == 3).sum() > 20           # the box is visible


def test_boxes_large_scene_and_mixed_order(gpu, renderer_mod, port, S):
    """Boxes interleaved with spheres and walls in scene order (ids and tie-breaks), big kernel with the FP32 screen."""
    rng = np.random.default_rng(21)
    syn = S.synthetic_scene(1500, 12, seed=4)
    boxes = _random_boxes(S, rng, 40)
    scene = []
    for k, g in enumerate(syn):
        scene.append(g)
        if k % 37 == 0 and boxes:
            scene.append(boxes.pop())
    scene += boxes
    pod = S.default_camera(128, 16.0 / 9.0).pod()
    got, st = render(gpu, renderer_mod, scene, pod, 6)
    exp = port.render(scene, pod, 6)
    check_frame(got, exp, st)
    box_ids = [k for k, g in enumerate(scene) if isinstance(g, S.Box)]
    assert np.isin(exp["object_id"], box_ids).sum() > 100


def test_camera_inside_a_box_and_touching_boxes(gpu, renderer_mod, port, S):
    """A room: the camera sits inside a big box (every primary ray hits a face from behind), with two boxes sharing a face
    (exact ties between faces of different objects: the lower scene id wins) and a sphere."""
    room = S.Box(S.Material((0.8, 0.8, 0.7), 0.2), (-3.0, -4.0, -2.0), (12.0, 8.0, 5.0))
    a = S.Box(S.Material((0.9, 0.1, 0.1), 0.5), (3.0, -1.0, -2.0), (1.0, 1.0, 1.0))
    b = S.Box(S.Material((0.1, 0.1, 0.9), 0.5), (3.0, 0.0, -2.0), (1.0, 1.0, 1.0))     # shares the face y = 0 with a
    scene = [room, a, b, S.Sphere(S.Material((0.2, 0.9, 0.2), 0.7), (4.0, 1.5, 0.0), 0.7)] + S.synthetic_scene(30, 2, seed=8)
    pod = S.default_camera(144, 16.0 / 9.0).pod()
    got, st = render(gpu, renderer_mod, scene, pod, 7)
    exp = port.render(scene, pod, 7)
    check_frame(got, exp, st)
    assert (exp["object_id"] == 0).sum() > 1000 and (exp["object_id"] >= 0).all()


def test_degenerate_boxes(gpu, renderer_mod, port, S):
    """Zero-thickness, negative and non-finite extents follow the specification (mostly: never hit), nothing crashes."""
    mat = S.Material((0.5, 0.5, 0.5), 0.3)
    scene = S.default_scene() + [S.Box(mat, (2.0, -0.5, -0.5), (0.0, 1.0, 1.0)), S.Box(mat, (2.2, 0.2, 0.2), (-1.0, 1.0, 1.0)),
                                 S.Box(mat, (2.0, 0.0, 0.0), (float("inf"), 1.0, 1.0)), S.Box(mat, (float("nan"), 0.0, 0.0), (1.0, 1.0, 1.0))]
    scene += S.synthetic_scene(20, 2, seed=5)
    scene.append(S.Wall(mat, (6.0, -3.0, -1.0), (-1, 0, 0), float("inf"), 2.0))      # a half-infinite strip: CAN be hit
    pod = S.default_camera(96, 1.0).pod()
    got, st = render(gpu, renderer_mod, scene, pod, 4)
    exp = port.render(scene, pod, 4)
    check_frame(got, exp, st)
    assert (exp["object_id"] == len(scene) - 1).sum() > 0 and (exp["object_id"] == 5).sum() > 0   # the strip and the unbounded box are visible


@pytest.mark.parametrize("which", ["default", "synthetic"])
def test_sun_matches_the_specification(gpu, renderer_mod, port, S, which):
    scene = S.default_scene() if which == "default" else S.synthetic_scene(800, 10, seed=2) + _random_boxes(S, np.random.default_rng(1), 5)
    pod = S.default_camera(120, 16.0 / 9.0).pod()
    kw = dict(sun_enabled=1) if which == "default" else dict(sun_enabled=1, sun_color=(0.6, 0.7, 1.1), sun_direction=(-2.0, 1.0, 0.5))
    p = port.default_params()
    p.max_depth = 6
    for k, v in kw.items():
        setattr(p, k, type(getattr(p, k))(*v) if isinstance(v, tuple) else v)
    got, st = render(gpu, renderer_mod, scene, pod, 6, **kw)
    exp = port.render(scene, pod, params=p)
    check_frame(got, exp, st)
    plain = port.render(scene, pod, 6)
    assert not np.array_equal(plain["radiance"], exp["radiance"])           # the sun really lights the scene


def _lsb(a, b):
    return np.abs(unpack(a).astype(int) - unpack(b).astype(int))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("pixels,frames,white,mode", [(4096, 1, 0.0, 1), (1001, 3, 0.0, 1), (777, 2, 3.0, 0), (1 << 18, 2, 1.5, 1)])
def test_standalone_tonemap(gpu, port, pkg, dtype, pixels, frames, white, mode):
    """rtx_tonemap (vector path for multiples of four pixels, scalar path otherwise) against orc_tonemap."""
    rng = np.random.default_rng(pixels)
    rgb = rng.lognormal(-1.0, 1.5, size=(frames, pixels, 3)).astype(dtype)
    rgb[0, : pixels // 20] = 0.0
    rgb[-1, 5] = (-1.0, 0.5, 0.25)                                        # a negative channel
    p = port.default_params()
    p.tonemap, p.tonemap_white, p.quantise_mode = pkg.abi.RTX_TONEMAP_REINHARD, white, mode
    got, lavg = gpu.tonemap(rgb, p)
    exp, lavg_exp = port.tonemap(rgb, p)
    # the same fixed-point sum up to a few ulps of log() (double radiance) / logf() (float radiance: processed in float)
    assert np.allclose(lavg, lavg_exp, rtol=1e-12 if rgb.dtype == np.float64 else 1e-6)
    d = _lsb(got, exp)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    got2, lavg2 = gpu.tonemap(rgb, p)                                      # integer atomics: run-to-run identical
    assert np.array_equal(got, got2) and np.array_equal(lavg, lavg2)


def test_render_with_tonemap_equals_trace_then_operator(gpu, renderer_mod, port, pkg, S):
    """rtx_params.tonemap inside rtx_render = the trace kernel's radiance followed by the operator (per frame)."""
    scene = S.synthetic_scene(600, 8, seed=6)
    pods = [c.pod() for c in S.flythrough_cameras(3, 96, 16.0 / 9.0)]
    gpu.set_scene(scene)
    params = renderer_mod.default_params(max_depth=5, tonemap=pkg.abi.RTX_TONEMAP_REINHARD, quantise_mode=pkg.abi.RTX_QUANT_SATURATE,
                                         tonemap_white=4.0, sun_enabled=1)
    planes, st = gpu.render(pods, params, want=("rgba8", "radiance_f64", "object_id"))
    p = port.default_params()
    p.max_depth, p.sun_enabled = 5, 1
    p.tonemap, p.tonemap_white, p.quantise_mode = pkg.abi.RTX_TONEMAP_REINHARD, 4.0, pkg.abi.RTX_QUANT_SATURATE
    rad = np.stack([port.render(scene, pod, params=p)["radiance"] for pod in pods])
    assert np.nanmax(np.abs(planes["radiance_f64"] - rad) / np.maximum(np.abs(rad), 1e-3)) < 1e-12
    exp, _ = port.tonemap(rad.reshape(len(pods), -1, 3), p)
    d = _lsb(planes["rgba8"].reshape(len(pods), -1), exp)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    assert st.launches == 3 and st.surface_update_ms > 0
    # without the operator the same call packs the radiance directly
    plain, _ = gpu.render(pods, renderer_mod.default_params(max_depth=5, sun_enabled=1, quantise_mode=pkg.abi.RTX_QUANT_SATURATE), want=("rgba8",))
    assert not np.array_equal(plain["rgba8"], planes["rgba8"])


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("H,W,ranks,band", [(54, 96, 4, 4), (50, 37, 3, 3)])
def test_row_sharded_tonemap_equals_single_gpu_bit_for_bit(gpu, renderer_mod, pkg, dtype, H, W, ranks, band):
    """rtx_tonemap_sums on every rank's rows, the integer sums added up (what the all-reduce does), rtx_tonemap_apply on
    every rank's rows: the assembled frames are identical to rtx_tonemap on the whole frames — the statistic is an
    integer sum, so neither the split nor the order matters. Ranks are emulated on one GPU; the second geometry has
    ragged bands and a pixel count that is not a multiple of four (scalar path)."""
    import torch
    SH = __import__("importlib").import_module("ray-tracer-from-scratch_b200.sharding")
    dev = torch.device("cuda", 0)
    F = 2
    g = torch.Generator(device="cpu").manual_seed(H * W)
    rad = torch.exp(torch.randn((F, H, W, 3), generator=g, dtype=torch.float64) * 1.5 - 1.0).to(getattr(torch, dtype)).to(dev)
    params = renderer_mod.default_params(tonemap=pkg.abi.RTX_TONEMAP_REINHARD, quantise_mode=pkg.abi.RTX_QUANT_SATURATE, tonemap_white=2.0)
    whole = torch.empty((F, H, W), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    gpu.tonemap_device(rad.data_ptr(), dtype == "float32", H * W, F, params, whole.data_ptr())
    # pass 1 on every rank's rows into ONE accumulator = local sums + integer all-reduce
    rows = [torch.as_tensor(renderer_mod.global_rows(H, band, ranks, r).astype("int64"), device=dev) for r in range(ranks)]
    local = [rad[:, rows[r]].contiguous() for r in range(ranks)]
    sums = torch.zeros(F, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for r in range(ranks):
        gpu.tonemap_sums_device(local[r].data_ptr(), dtype == "float32", local[r][0].numel() // 3, F, sums.data_ptr())
    assembled = torch.zeros((F, H, W), dtype=torch.int32, device=dev)
    for r in range(ranks):
        out = torch.empty(tuple(local[r].shape[:-1]), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        gpu.tonemap_apply_device(local[r].data_ptr(), dtype == "float32", local[r][0].numel() // 3, F, sums.data_ptr(), H * W, params, out.data_ptr())
        assembled[:, rows[r]] = out
    assert torch.equal(assembled, whole)
    # the one-process form of the product helper (world = 1) gives the same frames
    out1, sums1 = SH.tonemap_sharded(gpu, rad, H * W, params, 1)
    assert torch.equal(out1, whole) and torch.equal(sums1, sums)


def test_cpp_headless_main_with_extensions(tmp_path, renderer_mod, port, pkg, S):
    """The C++ facade end to end with the extensions on: Box in the scene, sun in the shading, and the tone map in the
    surface update (Renderer::update_surface -> rtx_tonemap), against the oracle's specification of the same pipeline."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(renderer_mod.__file__)), "rtx_headless")
    raw = tmp_path / "e.rgba"
    subprocess.run([exe, "--width", "192", "--aspect", "1.5", "--depth", "6", "--frames", "2", "--keys", "xd", "--sun", "1", "--tonemap", "1",
                    "--box", "1", "--raw", str(raw), "--out", ""], check=True, capture_output=True)
    scene = S.default_scene() + [S.Box(S.Material((0.9, 0.2, 0.2), 0.6), (2.5, -1.0, -0.8), (1.0, 1.2, 0.9))]
    cam = S.default_camera(192, 1.5)
    _, pod = port.camera_walk(cam, [("d", 0)])                       # frame 1: after one step to the right
    p = port.default_params()
    p.max_depth, p.sun_enabled = 6, 1
    p.tonemap, p.quantise_mode = pkg.abi.RTX_TONEMAP_REINHARD, pkg.abi.RTX_QUANT_SATURATE
    rad = port.render(scene, pod, params=p)["radiance"]
    exp, _ = port.tonemap(rad.reshape(1, -1, 3), p)
    got = np.fromfile(raw, dtype=np.uint32).reshape(pod.height, pod.width)
    d = _lsb(got.reshape(1, -1), exp)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    assert (port.render(scene, pod, params=p)["object_id"] == 3).sum() > 50     # the box is in view


def test_tonemap_error_paths(gpu, renderer_mod, pkg, S):
    gpu.set_scene(S.default_scene())
    pod = S.default_camera(32, 1.0).pod()
    with pytest.raises(renderer_mod.RtxError):
        gpu.render([pod], renderer_mod.default_params(tonemap=7))
    with pytest.raises(renderer_mod.RtxError):     # the statistic is global: not available on a row-sharded frame
        gpu.render([pod], renderer_mod.default_params(tonemap=pkg.abi.RTX_TONEMAP_REINHARD, n_ranks=2, rank=0, band_rows=4))
    with pytest.raises(renderer_mod.RtxError):
        gpu.tonemap(np.ones((1, 16, 3)), renderer_mod.default_params())      # params.tonemap not set
    with pytest.raises(renderer_mod.RtxError):
        gpu.tonemap(np.ones((1, 16, 3)), renderer_mod.default_params(tonemap=pkg.abi.RTX_TONEMAP_REINHARD, tonemap_key=0.0))
