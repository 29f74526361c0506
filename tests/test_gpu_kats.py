"""GPU suite: the reference's FUNCTION-LEVEL known answers replayed on the CUDA path.

tests/golden/kat.json (90 intersect / 30 find_closest_hit / 30 recursive_ray_tracing vectors) and
tests/golden/kat_crafted.json (the branches random rays never reach: det == 0 with its "/ a" quirk, scene.cpp:62-66;
denominator == 0 walls, scene.cpp:8-11; the back-face pass-through, main.cpp:111-113; exact ties; a zero direction)
were produced by the UNMODIFIED reference (oracle/_ref; make_golden.py, make_crafted.py). Each ray goes through
rtx_trace_rays — the C-ABI door for recursive_ray_tracing(scene, ray, depth) / find_closest_hit(scene, ray) — twice:

  * against the scene as it is (<= 16 entries: trace_small_kernel, exact tests only), and
  * against the same scene followed by 24 far-away dust spheres (trace_kernel: FP32 screen, queues, exact stage),

and distance, normal and index must equal the reference's BIT FOR BIT; radiance within 1e-12 (contract: 1e-4).
"""
import numpy as np
import pytest

from conftest import fh, fh3, load_json

pytestmark = pytest.mark.gpu

DBL_MAX = 1.7976931348623157e308
RAD_TOL = 1e-12


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def same_bits(a, b):
    """Bitwise equality of doubles, except that any NaN equals any NaN and -0 equals +0."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))))


def rad_close(got, exp):
    got, exp = np.asarray(got, dtype=np.float64), np.asarray(exp, dtype=np.float64)
    if np.isnan(exp).any():
        return bool(np.array_equal(np.isnan(got), np.isnan(exp)))
    return bool((np.abs(got - exp) / np.maximum(np.abs(exp), 1e-3)).max() < RAD_TOL)


def geometry(S, o):
    """A scene object from its kat.json / kat_crafted.json record (material optional: intersect() does not read it)."""
    m = o.get("mat")
    mat = S.Material(fh3(m["color"]), fh(m["metallic"]), fh(m["ambient"]), fh(m["diffuse"]), fh(m["specular"]),
                     fh(m["specular_exponent"])) if m else S.Material((1, 1, 1))
    if o["kind"] == 0:
        return S.Sphere(mat, fh3(o["p"]), fh(o["a"]))
    return S.Wall(mat, fh3(o["p"]), fh3(o["n"]), fh(o["a"]), fh(o["b"]))


def dust(S, n=24):
    """Spheres no test ray comes near: they only push the scene over the small-kernel limit (16 entries)."""
    return [S.Sphere(S.Material((1, 1, 1)), (900.0 + 3 * k, 800.0 - 2 * k, 700.0 + k), 1e-6) for k in range(n)]


def both_kernels(S, scene):
    return (("small kernel", scene), ("big kernel", list(scene) + dust(S)))


def test_intersect_kats_on_gpu(gpu, renderer_mod, S):
    """SceneGeometry::intersect (scene.cpp:4-78) of a single object: find_closest_hit over a one-object scene accepts it
    iff distance > 0 (main.cpp:77) and then reports exactly intersect()'s distance and normal."""
    kat = load_json("kat.json")["intersect"]
    hits = 0
    for c in kat:
        g = geometry(S, c)
        want_d, want_n = fh(c["distance"]), fh3(c["normal"])
        for which, scene in both_kernels(S, [g]):
            gpu.set_scene(scene)
            got, st = gpu.trace_rays([(fh3(c["o"]), fh3(c["d"]))], renderer_mod.default_params(max_depth=0))
            if want_d > 0:
                assert got["object_id"][0] == 0, (which, c)
                assert same_bits(got["hit_distance"][0], want_d), (which, c, got["hit_distance"][0])
                assert same_bits(got["hit_normal"][0], want_n), (which, c, got["hit_normal"][0])
            else:   # miss, or a "hit" behind the origin that find_closest_hit drops
                assert got["object_id"][0] == -1, (which, c)
                assert got["hit_distance"][0] == DBL_MAX and not got["hit_normal"][0].any(), (which, c)
            assert st.total_rays == 1
        hits += want_d > 0
    assert hits >= 30            # the fixture really exercises both outcomes


def test_closest_and_trace_kats_on_gpu(gpu, renderer_mod, S):
    """find_closest_hit (main.cpp:67-84) and recursive_ray_tracing (main.cpp:89-119) on the default scene, all 30 + 30
    reference vectors in ONE batched call per kernel."""
    kat = load_json("kat.json")
    assert {c["scene"] for c in kat["closest"]} == {"default"} and {c["scene"] for c in kat["trace"]} == {"default"}
    rays = [(fh3(c["o"]), fh3(c["d"])) for c in kat["closest"]]
    assert rays == [(fh3(c["o"]), fh3(c["d"])) for c in kat["trace"]]          # the generator pairs them
    for which, scene in both_kernels(S, S.default_scene()):
        gpu.set_scene(scene)
        got, st = gpu.trace_rays(rays, renderer_mod.default_params(max_depth=10))
        for k, (c, t) in enumerate(zip(kat["closest"], kat["trace"])):
            assert got["object_id"][k] == c["index"], (which, k)
            assert same_bits(got["hit_distance"][k], fh(c["distance"])), (which, k)
            assert same_bits(got["hit_normal"][k], fh3(c["normal"])), (which, k)
            assert rad_close(got["radiance_f64"][k], fh3(t["rgb"])), (which, k, got["radiance_f64"][k], fh3(t["rgb"]))
        assert st.total_rays == int(got["ray_count"].sum())


def test_crafted_branches_on_gpu(gpu, renderer_mod, S):
    """det == 0 (distance 4 where the geometry says 2), denominator == 0, back-face pass-through, exact ties, NaN rays."""
    cases = load_json("kat_crafted.json")["cases"]
    for c in cases:
        scene = [geometry(S, o) for o in c["objects"]]
        ray = (fh3(c["o"]), fh3(c["d"]))
        exp = c["closest"]
        for which, sc in both_kernels(S, scene):
            tag = (c["what"], which)
            gpu.set_scene(sc)
            got, st = gpu.trace_rays([ray], renderer_mod.default_params(max_depth=c["depth"]))
            assert got["object_id"][0] == exp["index"], tag
            assert same_bits(got["hit_distance"][0], fh(exp["distance"])), (tag, got["hit_distance"][0])
            assert same_bits(got["hit_normal"][0], fh3(exp["normal"])), (tag, got["hit_normal"][0])
            assert rad_close(got["radiance_f64"][0], fh3(c["rgb"])), (tag, got["radiance_f64"][0], fh3(c["rgb"]))
            # every object on its own: intersect() as the reference returns it
            for k, (g, e) in enumerate(zip(scene, c["intersect"])):
                gpu.set_scene([g] if which == "small kernel" else [g] + dust(S))
                one, _ = gpu.trace_rays([ray], renderer_mod.default_params(max_depth=0))
                d = fh(e["distance"])
                if d > 0:
                    assert one["object_id"][0] == 0 and same_bits(one["hit_distance"][0], d), (tag, k, one["hit_distance"][0])
                    assert same_bits(one["hit_normal"][0], fh3(e["normal"])), (tag, k)
                else:
                    assert one["object_id"][0] == -1, (tag, k)
    # the quirks themselves, spelled out (values from the reference, see the fixture):
    by = {c["what"].split(":")[0].split(",")[0] + "|" + str(c["depth"]): c for c in cases}
    assert fh(cases[0]["closest"]["distance"]) == 4.0                          # tangent sphere at x = 2: "/ a", not "/ 2a"
    assert "back-face pass-through (SURVEY §8(a) row M)|10" in by
    gpu.set_scene(S.default_scene())
    got, _ = gpu.trace_rays([((2.4, 5, .4), (.1, -1, .05))], renderer_mod.default_params(max_depth=10))
    assert got["ray_count"][0] == 4                                            # wall, same wall again, other wall, sky


def test_ray_batch_equals_camera_rays(gpu, renderer_mod, S, port):
    """rtx_trace_rays on the primary rays of a small frame == rtx_render of that frame (same kernels, same pixels)."""
    scene = S.synthetic_scene(600, 8)
    pod = S.default_camera(48, 16.0 / 9.0).pod()
    gpu.set_scene(scene)
    frame, st = gpu.render([pod], renderer_mod.default_params(max_depth=6), want=("radiance_f64", "object_id", "ray_count", "rgba8"))
    rays = []
    pos = np.array(pod.position.tuple())
    tl, dx, dy = (np.array(v.tuple()) for v in (pod.image_top_left, pod.delta_x, pod.delta_y))
    for i in range(pod.height):
        for j in range(pod.width):
            centre = (tl + dx * float(j)) + dy * float(i)                          # main.cpp:132, same association
            rays.append((tuple(pos), tuple(pos - centre)))
    got, st2 = gpu.trace_rays(rays, renderer_mod.default_params(max_depth=6), want=("radiance_f64", "object_id", "ray_count", "rgba8"))
    assert np.array_equal(got["object_id"], frame["object_id"][0].ravel())
    assert np.array_equal(got["ray_count"], frame["ray_count"][0].ravel())
    assert np.array_equal(got["rgba8"], frame["rgba8"][0].ravel())
    assert np.array_equal(bits(got["radiance_f64"]), bits(frame["radiance_f64"][0].reshape(-1, 3)))
    assert st2.total_rays == st.total_rays
