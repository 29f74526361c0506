"""GPU suite: the uniform-grid extension (rtx_params.accel = RTX_ACCEL_GRID; README.md:17 names an
acceleration structure as the next step, the snapshot has none).

The grid only narrows WHICH objects reach the exact tests; the tests and the acceptance rule are the brute-force ones.
So the bar is not a tolerance: every plane — ids, hit distances and normals, ray counts, RGBA8 and the double radiance —
must equal the brute-force kernel's BIT FOR BIT, and through it the reference (full 4K frame of the 10 064-object scene
against the unmodified reference's per-row CRCs)."""
import numpy as np
import pytest

from conftest import fh, fh3, load_json
from test_gpu_fullsize import check_against_fullsize
from test_gpu_kats import bits, dust, geometry

pytestmark = pytest.mark.gpu

PLANES = ("rgba8", "radiance_f64", "object_id", "hit_mask", "ray_count", "hit_distance", "hit_normal")


def both(gpu, renderer_mod, scene, pods, depth, **kw):
    """The same frames brute force and through the grid; asserts that every plane is bitwise identical."""
    gpu.set_scene(scene)
    pods = pods if isinstance(pods, list) else [pods]
    a, sa = gpu.render(pods, renderer_mod.default_params(max_depth=depth, accel=0, **kw), want=PLANES)
    b, sb = gpu.render(pods, renderer_mod.default_params(max_depth=depth, accel=1, **kw), want=PLANES)
    for name in PLANES:
        x, y = a[name], b[name]
        if x.dtype == np.float64:
            same = (bits(x) == bits(y)) | (np.isnan(x) & np.isnan(y))
            assert same.all(), (name, int((~same).sum()))
        else:
            assert np.array_equal(x, y), name
    assert sa.total_rays == sb.total_rays and sa.over_range_pixels == sb.over_range_pixels
    return a, sa, sb


def test_grid_equals_brute_force_on_the_synthetic_scene(gpu, renderer_mod, S):
    syn = S.synthetic_scene()
    for width, depth in ((192, 10), (97, 3)):
        a, sa, sb = both(gpu, renderer_mod, syn, S.default_camera(width, 16.0 / 9.0).pod(), depth)
        assert (a["object_id"] >= 10000).any() and (a["ray_count"] > 3).any()        # walls and deep chains occur
    # other viewpoints: inside the cloud of spheres, looking along each axis, from far away
    for pos, look in (((30, 0, 5), (29, 0, 5)), ((34, 3, 8), (34, 4, 8)), ((34, 3, 30), (34, 3, 31)), ((-300, 10, 4), (-301, 10, 4)),
                      ((34.5, -2.25, 6.125), (34.5, -2.25, 7.125))):
        cam = S.Camera()
        cam.aspect_ratio, cam.image_width, cam.vfov = 1.5, 96, 70
        cam.position, cam.lookat, cam.vup = pos, look, (0, 0, -1) if look[2] == pos[2] else (0, 1, 0)
        cam.init()
        both(gpu, renderer_mod, syn, cam.pod(), 8)


def test_grid_with_walls_boxes_ties_and_odd_objects(gpu, renderer_mod, S):
    M = S.Material
    nan, inf = float("nan"), float("inf")
    rng = np.random.default_rng(5)
    cloud = [S.Sphere(M(tuple(rng.uniform(.1, 1, 3)), float(rng.uniform(0, .8))), tuple(rng.uniform(-6, 6, 3) + (9, 0, 0)), float(rng.uniform(.05, .7)))
             for _ in range(300)]
    odd = [
        S.Wall(M((.2, .3, .9), .3), (7.0, 2, 0), (0, -1, 0), 3, 3),
        S.Sphere(M((.9, .2, .1), .6), (6.0, 0.3, 0.2), .4), S.Sphere(M((.1, .9, .1), .2), (6.0, 0.3, 0.2), .4),   # exact duplicates: tie
        S.Box(M((.8, .8, .2), .4), (8, -4, -2), (1.5, 2, 1)),
        S.Sphere(M((1, 0, 0)), (nan, 0, 0), .5), S.Sphere(M((1, 0, 0)), (5, 0, 0), nan), S.Sphere(M((1, 0, 0)), (inf, 0, 0), .5),
        S.Sphere(M((0, 1, 1), .9), (5.0, -1.0, 0.5), -0.25),                                                       # negative radius
        S.Sphere(M((.3, .3, .3), .5), (9, 0, -1005), 1000.0),                                                     # a "ground" sphere: fills the grid -> always list
        S.Sphere(M((.5, .5, .9), .5), (9, 0, 0), 0.0),                                                            # zero radius
    ]
    scene = cloud[:150] + odd + cloud[150:]
    a, sa, sb = both(gpu, renderer_mod, scene, S.default_camera(128, 1.0).pod(), 12)
    ids = set(np.unique(a["object_id"]))
    assert 150 + 8 in ids            # the ground sphere is seen
    assert 150 + 2 not in ids        # the duplicate with the higher index never wins
    # sun + saturating pack go through the same shading code
    both(gpu, renderer_mod, scene, S.default_camera(64, 1.0).pod(), 5, sun_enabled=1, quantise_mode=1)


def test_grid_degenerate_grids(gpu, renderer_mod, S):
    """One sphere, coplanar spheres (a flat grid), collinear spheres, all spheres at one point, no spheres at all (walls only)."""
    M = S.Material
    m = M((.7, .4, .2), .5)
    pod = S.default_camera(64, 1.0).pod()
    walls = [S.Wall(M((.2, .3, .9), .3), (3.0 + k, 2, -1), (0, -1, 0), 1, 1) for k in range(20)]
    for spheres in ([S.Sphere(m, (3, 0, 0), .5)],
                    [S.Sphere(m, (3 + (k % 5), (k // 5) - 2.0, 0.0), .3) for k in range(25)],
                    [S.Sphere(m, (2 + k, 0, 0), .3) for k in range(25)],
                    [S.Sphere(m, (4, 0.5, 0.25), .3 + .01 * k) for k in range(25)],
                    []):
        both(gpu, renderer_mod, spheres + walls, pod, 6)


def test_grid_rays_from_outside_far_away_and_parallel_to_axes(gpu, renderer_mod, S):
    """rtx_trace_rays through the grid: origins far outside the scene bound (exact fallback), rays along the axes and
    along cell boundaries, rays that miss the grid box, a zero direction."""
    syn = S.synthetic_scene(2000, 16)
    rays = [((0, 0, 0), (1, 0, 0)), ((0, 0, 0), (0, 1, 0)), ((0, 0, 0), (0, 0, 1)), ((34, 0, -50), (0, 0, 1)), ((34, 0, 50), (0, 0, -1)),
            ((1e6, 0, 0), (-1, 0, 0)), ((-1e6, 3, 2), (1, 0, 0)), ((34, 0, 8), (0, 0, 0)), ((34, 0, 8), (1e-300, 0, 0)), ((34, 0, 8), (1e300, 1e300, 0)),
            ((200, 200, 200), (1, 1, 1)), ((4, -32, -8), (1, 1, .5)), ((64, 32, 24), (-1, -1, -.5))]
    rng = np.random.default_rng(9)
    rays += [(tuple(rng.uniform(-10, 80, 3)), tuple(rng.normal(size=3))) for _ in range(3000)]
    gpu.set_scene(syn)
    a, sa = gpu.trace_rays(rays, renderer_mod.default_params(max_depth=10, accel=0))
    b, sb = gpu.trace_rays(rays, renderer_mod.default_params(max_depth=10, accel=1))
    for name in a:
        x, y = a[name], b[name]
        if x.dtype == np.float64:
            assert ((bits(x) == bits(y)) | (np.isnan(x) & np.isnan(y))).all(), name
        else:
            assert np.array_equal(x, y), name
    assert sa.total_rays == sb.total_rays and (a["object_id"] >= 0).sum() > 50


def test_grid_reference_kats(gpu, renderer_mod, S):
    """The crafted reference vectors (det == 0, denominator == 0, back-face pass-through, ties, NaN) through the grid."""
    for c in load_json("kat_crafted.json")["cases"]:
        scene = [geometry(S, o) for o in c["objects"]] + dust(S)
        gpu.set_scene(scene)
        got, _ = gpu.trace_rays([(fh3(c["o"]), fh3(c["d"]))], renderer_mod.default_params(max_depth=c["depth"], accel=1))
        exp = c["closest"]
        assert got["object_id"][0] == exp["index"], c["what"]
        d = fh(exp["distance"])
        assert got["hit_distance"][0] == d or (np.isnan(d) and np.isnan(got["hit_distance"][0])), (c["what"], got["hit_distance"][0])
        want = np.array(fh3(c["rgb"]))
        if np.isnan(want).any():
            assert np.isnan(got["radiance_f64"][0]).all()
        else:
            assert (np.abs(got["radiance_f64"][0] - want) / np.maximum(np.abs(want), 1e-3)).max() < 1e-12, c["what"]


def test_grid_scene_change_rebuilds_the_grid(gpu, renderer_mod, S):
    pod = S.default_camera(64, 16.0 / 9.0).pod()
    for seed in (1, 2):
        both(gpu, renderer_mod, S.synthetic_scene(500, 8, seed=seed), pod, 6)


def test_grid_larger_than_shared_memory_reads_from_l2(gpu, renderer_mod, S):
    both(gpu, renderer_mod, S.synthetic_scene(20000, 32, seed=21), S.default_camera(96, 16.0 / 9.0).pod(), 6)


def test_grid_c3_full_4k_frame_matches_reference(gpu, renderer_mod, S):
    """Config C3 through the grid: every row of the 4K frame against the UNMODIFIED reference's per-row CRCs."""
    g = load_json("fullsize_c3.json")
    gpu.set_scene(S.synthetic_scene())
    planes, st = gpu.render([S.default_camera(3840, 16.0 / 9.0).pod()], renderer_mod.default_params(max_depth=10, accel=1),
                            want=("rgba8", "ray_count", "object_id"))
    check_against_fullsize(planes, st, g, np.array(g["rows"]))
    assert st.total_rays == g["total_rays"] and st.over_range_pixels == sum(g["row_over_range"])


def test_grid_row_bands_and_frame_batches(gpu, renderer_mod, S):
    syn = S.synthetic_scene(1500, 16)
    pod = S.default_camera(80, 16.0 / 9.0).pod()
    gpu.set_scene(syn)
    full, _ = gpu.render([pod], renderer_mod.default_params(max_depth=6, accel=1), want=("rgba8",))
    for r in range(3):
        rows = renderer_mod.global_rows(pod.height, 2, 3, r)
        part, _ = gpu.render([pod], renderer_mod.default_params(max_depth=6, accel=1, band_rows=2, n_ranks=3, rank=r), want=("rgba8",))
        assert np.array_equal(part["rgba8"][0], full["rgba8"][0][rows])
    cams = [c.pod() for c in S.flythrough_cameras(256, 64, 16.0 / 9.0)[::64]]
    both(gpu, renderer_mod, syn, cams, 5)
