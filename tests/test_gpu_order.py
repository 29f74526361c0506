"""GPU suite: the scheduling hint of the big kernel (rtx_params.pixel_order).

RTX_ORDER_COST hands out tiles of 256 pixels most expensive first, judged by the ray counts the previous call collected.
The reference traces every pixel independently (main.cpp:124-139: one recursive_ray_tracing per pixel, no state shared
between pixels), so WHEN a pixel is traced cannot show anywhere: every plane and every statistic must equal the scan-order
call's bit for bit — for frames whose pixel count is and is not a multiple of the tile, row bands, frame batches, calls in
flight, and a camera that moves between calls (the order then comes from a DIFFERENT frame than the one it is used for,
and a pixel the order lost would keep the previous frame's value and differ)."""
import numpy as np
import pytest

from test_gpu_kats import bits

pytestmark = pytest.mark.gpu

PLANES = ("rgba8", "radiance_f64", "object_id", "ray_count", "hit_distance")


def path(S, n, width, aspect):
    """n cameras of a walk through the synthetic scene (every frame differs from the one before)."""
    pods = []
    for k in range(n):
        cam = S.Camera()
        cam.aspect_ratio, cam.image_width, cam.vfov = aspect, width, 60
        cam.position, cam.lookat, cam.vup = (-4.0 + 1.5 * k, 0.4 * k, 3.0 + 0.25 * k), (30.0, 2.0 - k, 5.0), (0, 0, -1)
        cam.init()
        pods.append(cam.pod())
    return pods


def same(a, b):
    for name in a:
        x, y = a[name], b[name]
        if x.dtype == np.float64:
            assert ((bits(x) == bits(y)) | (np.isnan(x) & np.isnan(y))).all(), name
        else:
            assert np.array_equal(x, y), name


@pytest.fixture()
def own(renderer_mod):
    """A context of its own: the order is state of the context (the previous call's costs)."""
    r = renderer_mod.Renderer(0)
    yield r
    r.close()


@pytest.mark.parametrize("width,aspect", [(192, 16.0 / 9.0), (250, 1.37), (40, 2.0)])
def test_cost_order_changes_nothing_along_a_camera_path(own, renderer_mod, pkg, S, width, aspect):
    """192x108 = 81 whole tiles; 250x182 = 177 tiles + 188 pixels (the partial tile keeps the last place); 40x20 = 3 tiles + 32."""
    own.set_scene(S.synthetic_scene(1500, 16))
    pods = path(S, 5, width, aspect)
    scan = [own.render([pod], renderer_mod.default_params(max_depth=10, pixel_order=pkg.abi.RTX_ORDER_SCAN), want=PLANES) for pod in pods]
    for k, pod in enumerate(pods):          # frame k is handed out in the order of frame k-1's costs (frame 0: scan order)
        got, st = own.render([pod], renderer_mod.default_params(max_depth=10, pixel_order=pkg.abi.RTX_ORDER_COST), want=PLANES)
        same(scan[k][0], got)
        assert st.total_rays == scan[k][1].total_rays and st.over_range_pixels == scan[k][1].over_range_pixels
        assert st.max_luminance == scan[k][1].max_luminance
        assert st.launches == (5 if k == 0 else 4) and scan[k][1].launches == 1     # reset + trace + three order kernels; then trace + three
    # the same frame again and again: the order is now the frame's own
    for _ in range(3):
        got, st = own.render([pods[-1]], renderer_mod.default_params(max_depth=10, pixel_order=pkg.abi.RTX_ORDER_COST), want=PLANES)
        same(scan[-1][0], got)


def test_cost_order_with_row_bands_batches_and_changing_geometry(own, renderer_mod, pkg, S):
    own.set_scene(S.synthetic_scene(1200, 16, seed=4))
    cost = pkg.abi.RTX_ORDER_COST
    pods = path(S, 3, 160, 16.0 / 9.0)
    full = [own.render([pod], renderer_mod.default_params(max_depth=8, pixel_order=pkg.abi.RTX_ORDER_SCAN), want=("rgba8", "ray_count"))[0] for pod in pods]
    # every rank's rows twice in a row (second call ordered), other camera in between (changing the rank resets the order)
    for r in range(3):
        rows = renderer_mod.global_rows(pods[0].height, 2, 3, r)
        for k in (0, 1, 2, 2):
            part, _ = own.render([pods[k]], renderer_mod.default_params(max_depth=8, band_rows=2, n_ranks=3, rank=r, pixel_order=cost), want=("rgba8", "ray_count"))
            for name in part:
                assert np.array_equal(part[name][0], full[k][name][0][rows]), (name, r, k)
    # frame batches: one pixel space over all frames of the call
    for _ in range(3):
        batch, st = own.render(pods, renderer_mod.default_params(max_depth=8, pixel_order=cost), want=("rgba8", "ray_count"))
        for k in range(3):
            assert np.array_equal(batch["rgba8"][k], full[k]["rgba8"][0]) and np.array_equal(batch["ray_count"][k], full[k]["ray_count"][0])
    # another frame size in between, then back
    small = path(S, 1, 96, 1.5)[0]
    ref_small, _ = own.render([small], renderer_mod.default_params(max_depth=8, pixel_order=pkg.abi.RTX_ORDER_SCAN), want=("rgba8",))
    for pod, ref in ((small, ref_small), (pods[1], full[1]), (small, ref_small), (small, ref_small), (pods[1], full[1])):
        got, _ = own.render([pod], renderer_mod.default_params(max_depth=8, pixel_order=cost), want=("rgba8",))
        assert np.array_equal(got["rgba8"], ref["rgba8"])


def test_cost_order_with_calls_in_flight_and_other_kernels_in_between(own, renderer_mod, pkg, S):
    """rtx_render_async keeps several frames in flight on one stream: frame k+1's kernel starts after frame k's order has
    been built. Ray batches, the grid and a small scene in between leave the order alone."""
    syn = S.synthetic_scene(1000, 16, seed=8)
    own.set_scene(syn)
    pods = path(S, 6, 128, 16.0 / 9.0)
    cost = renderer_mod.default_params(max_depth=7, pixel_order=pkg.abi.RTX_ORDER_COST)
    scan = [own.render([pod], renderer_mod.default_params(max_depth=7, pixel_order=pkg.abi.RTX_ORDER_SCAN), want=("rgba8",))[0]["rgba8"] for pod in pods]
    depth = pkg.abi.RTX_MAX_IN_FLIGHT
    host = [np.zeros((pods[0].height, pods[0].width), np.uint32) for _ in pods]
    outs = []
    for h in host:
        o = pkg.abi.Outputs()
        o.memory, o.rgba8 = pkg.abi.RTX_MEM_HOST, h.ctypes.data
        outs.append(o)
    for k, pod in enumerate(pods):
        own.render_async([pod], cost, outs[k])
        if k >= depth - 1:
            own.wait()
    for _ in range(depth - 1):
        own.wait()
    for k in range(len(pods)):
        assert np.array_equal(host[k], scan[k][0]), k
    own.trace_rays([((0, 0, 0), (1, 0, 0))] * 40, cost)                                    # a ray batch: no order
    own.render([pods[0]], renderer_mod.default_params(max_depth=7, accel=1, pixel_order=pkg.abi.RTX_ORDER_COST), want=("rgba8",))   # grid kernel: no order
    got, st = own.render([pods[3]], cost, want=("rgba8",))
    assert np.array_equal(got["rgba8"], scan[3]) and st.launches == 4                      # the order of the last async frame is still there


def test_auto_order_from_2_pow_18_pixels_and_only_for_the_big_kernel(own, renderer_mod, pkg, S):
    syn = S.synthetic_scene(400, 8, seed=2)
    own.set_scene(syn)
    big, small = path(S, 2, 704, 16.0 / 9.0), path(S, 1, 640, 16.0 / 9.0)[0]                # 704x396 = 278 784 >= 2^18 > 640x360
    ref = [own.render([p], renderer_mod.default_params(max_depth=6, pixel_order=pkg.abi.RTX_ORDER_SCAN), want=("rgba8",)) for p in big]
    assert ref[0][1].launches == 1
    for k in (0, 1, 1):
        got, st = own.render([big[k]], renderer_mod.default_params(max_depth=6), want=("rgba8",))
        assert np.array_equal(got["rgba8"], ref[k][0]["rgba8"]) and st.launches >= 3
    _, st = own.render([small], renderer_mod.default_params(max_depth=6), want=("rgba8",))
    assert st.launches == 1                                                                  # below the threshold: scan order
    own.set_scene(S.default_scene())                                                         # 3 objects: trace_small_kernel
    _, st = own.render([big[0]], renderer_mod.default_params(max_depth=6, pixel_order=pkg.abi.RTX_ORDER_COST), want=("rgba8",))
    assert st.launches in (1, 4)                                             # one launch, or the four pixel ranges
    with pytest.raises(renderer_mod.RtxError):
        own.render([small], renderer_mod.default_params(pixel_order=3))
