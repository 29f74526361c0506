#!/usr/bin/env python
"""Crafted known answers FROM THE UNMODIFIED REFERENCE (oracle/_ref) for the branches random rays never reach:

  * Sphere::intersect's det == 0 branch with its "/ a instead of / 2a" quirk (scene.cpp:62-66),
  * Wall::intersect with denominator == 0: -inf, NaN and +inf quotients (scene.cpp:8-11),
  * the back-face wall pass-through of recursive_ray_tracing (main.cpp:111-113 + vec.cpp:51-57),
  * exact ties between objects (strict '<' in main.cpp:77: the lowest index wins),
  * a zero-length direction (NaN everywhere downstream).

Each case: a scene (list of objects), a ray, the reference's intersect() of every object, find_closest_hit and
recursive_ray_tracing(depth). tests/test_gpu_parity.py::test_reference_kats_on_gpu replays them through rtx_trace_rays.

    python tests/golden/make_crafted.py        # writes tests/golden/kat_crafted.json
"""
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

S = importlib.import_module("ray-tracer-from-scratch_b200").scene


def hx(x):
    return float(x).hex()


def hx3(t):
    return [hx(v) for v in t]


def obj_json(g):
    pod = g.pod()
    m = pod.mat
    return {"kind": g.kind, "p": hx3(pod.p.tuple()), "n": hx3(pod.n.tuple()), "a": hx(pod.a), "b": hx(pod.b),
            "mat": {"color": hx3(m.color.tuple()), "ambient": hx(m.ambient), "metallic": hx(m.metallic), "diffuse": hx(m.diffuse),
                    "specular": hx(m.specular), "specular_exponent": hx(m.specular_exponent)}}


def main():
    ref = ob.load_reference()
    default = S.default_scene()
    green = S.Material((0, 1, 0), 0.5)
    tangent = [S.Sphere(green, (2, .5, 0), .5)]
    twins = [S.Sphere(S.Material((1, 0, 0), 0.3), (3, 0, 0), .5), S.Sphere(S.Material((0, 0, 1), 0.7), (3, 0, 0), .5)]
    wall_and_sphere_tie = [S.Wall(S.Material((1, 1, 0)), (2, -1, -1), (-1, 0, 0), 2, 2), S.Sphere(green, (3, 0, 0), 1.0)]
    cases = [
        ("sphere det == 0, |d| = 1: distance 4 instead of 2 (scene.cpp:65 divides by a)", tangent, (0, 0, 0), (1, 0, 0), 10),
        ("sphere det == 0, |d| = 2", tangent, (0, 0, 0), (2, 0, 0), 10),
        ("sphere det == 0 seen from the other side", tangent, (4, 0, 0), (-1, 0, 0), 3),
        ("wall denominator == 0, numerator < 0: t = -inf (scene.cpp:8-11)", default, (0, 0, .5), (1, 0, 0), 10),
        ("wall denominator == 0, numerator == 0: t = NaN", default, (0, 2, .5), (1, 0, 0), 10),
        ("wall denominator == 0, numerator > 0: t = +inf, hit point NaN", default, (0, 3, .5), (1, 0, 0), 10),
        ("back-face pass-through (SURVEY §8(a) row M): id1 at 3.0, id1 again at 1.006e-4, id2, sky", default, (2.4, 5, .4), (.1, -1, .05), 10),
        ("back-face pass-through with the depth budget running out inside it", default, (2.4, 5, .4), (.1, -1, .05), 1),
        ("exact tie between two identical spheres: lowest index wins (main.cpp:77)", twins, (0, 0, 0), (1, 0, 0), 4),
        ("exact tie between a wall (t = 2) and a sphere (distance 2): lowest index wins", wall_and_sphere_tie, (0, 0, 0), (1, 0, 0), 4),
        ("exact tie, order swapped", wall_and_sphere_tie[::-1], (0, 0, 0), (1, 0, 0), 4),
        ("zero direction: NaN distance, NaN sky", default, (0, 0, 0), (0, 0, 0), 10),
        ("origin inside the sphere: no hit from inside (scene.cpp:70-72 picks the negative root)", default, (1.5, 0, 0), (1, 0, 0), 10),
        ("straight up: z-parallel ray (the screen's u axis degenerates -> exact fallback)", default, (1.5, 0, -3), (0, 0, 1), 10),
    ]
    out = {"source": "oracle/_ref (unmodified reference), tests/golden/make_crafted.py", "cases": []}
    for what, scene, o, d, depth in cases:
        per_object = []
        for g in scene:
            dist, nrm, hit = ref.intersect(g, o, d)
            per_object.append({"distance": hx(dist), "normal": hx3(nrm), "hit": bool(hit)})
        dist, nrm, idx = ref.find_closest_hit(scene, o, d)
        rgb = ref.trace_ray(scene, o, d, depth)
        out["cases"].append({"what": what, "objects": [obj_json(g) for g in scene], "o": hx3(o), "d": hx3(d), "depth": depth,
                             "intersect": per_object, "closest": {"distance": hx(dist), "normal": hx3(nrm), "index": idx},
                             "rgb": hx3(rgb)})
        print("%-90s closest id %2d dist %-22r rgb %s" % (what[:90], idx, dist, tuple(round(c, 6) for c in rgb)))
    assert float.fromhex(out["cases"][0]["intersect"][0]["distance"]) == 4.0, "the det == 0 branch was not reached"
    with open(os.path.join(HERE, "kat_crafted.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
