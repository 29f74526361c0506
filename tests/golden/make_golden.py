#!/usr/bin/env python
"""Generates the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Every value is produced by oracle/_ref/libref_oracle.so, i.e. /root/reference/{main,scene,vec}.cpp compiled
as they are (oracle/build_ref.sh). The reference has no tests or golden data of its own, so
these vectors — plus the ones the survey captured by the same means (repeated in
survey_vectors.json) — are what pins the C oracle and the CUDA path. Re-run after rebuilding oracle/_ref:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there read only the files written here.
"""
import hashlib
import importlib
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402

S = importlib.import_module("ray-tracer-from-scratch_b200").scene


def hx(v):
    return float(v).hex()


def hx3(t):
    return [hx(t[0]), hx(t[1]), hx(t[2])]


def frame_fixture(ref, scene, cam_pod, depth, path, rows=None):
    r = ref.render(scene, cam_pod, depth, rows=rows, threads=0)
    np.savez_compressed(path, radiance=r["radiance"], rgba8=r["rgba8"], object_id=r["object_id"],
                        hit_mask=r["hit_mask"], ray_count=r["ray_count"],
                        rows=np.arange(cam_pod.height, dtype=np.int32) if rows is None else np.asarray(rows, np.int32),
                        meta=np.array([cam_pod.width, cam_pod.height, depth, r["total_rays"]], dtype=np.int64))
    return r


def main():
    ref = ob.load_reference()
    rng = random.Random(0xB200)

    # ---- C1: the reference's own default frame (640x640, depth 10) -----------------------------------
    scene = S.default_scene()
    cam = S.default_camera()
    pod = ref.camera_init(cam)
    c1 = ref.render(scene, pod, 10)
    ids = c1["object_id"]
    table = []
    for (i, j) in [(0, 0), (319, 320), (320, 320), (300, 0), (300, 639), (250, 250), (260, 420), (400, 320), (215, 320),
                   (639, 639)]:
        o = pod.position.tuple()
        c = [pod.image_top_left.tuple()[k] + pod.delta_x.tuple()[k] * j + pod.delta_y.tuple()[k] * i for k in range(3)]
        d = [o[k] - c[k] for k in range(3)]
        dist, _, idx = ref.find_closest_hit(scene, o, d)
        table.append({"i": i, "j": j, "id": int(ids[i, j]), "distance": hx(dist) if idx >= 0 else None,
                      "rgb": hx3(c1["radiance"][i, j]),
                      "rgb8": [int(x) for x in ob.rgb8_bytes(c1["rgba8"][i, j])]})
    main_surface = ref.run_main(1)
    c1_doc = {
        "source": "oracle/_ref (unmodified reference), default scene/camera main.cpp:146-163, depth 10",
        "width": pod.width, "height": pod.height,
        "camera": {"image_top_left": hx3(pod.image_top_left.tuple()), "delta_x": hx3(pod.delta_x.tuple()),
                   "delta_y": hx3(pod.delta_y.tuple())},
        "id_histogram": {str(k): int((ids == k).sum()) for k in (-1, 0, 1, 2)},
        "total_rays": c1["total_rays"],
        "rgb8_sha256": hashlib.sha256(ob.rgb8_bytes(c1["rgba8"]).tobytes()).hexdigest(),
        "rgba8_sha256": hashlib.sha256(c1["rgba8"].tobytes()).hexdigest(),
        "radiance_sha256": hashlib.sha256(c1["radiance"].tobytes()).hexdigest(),
        "object_id_sha256": hashlib.sha256(ids.tobytes()).hexdigest(),
        "ray_count_sha256": hashlib.sha256(c1["ray_count"].tobytes()).hexdigest(),
        "main_surface_sha256": hashlib.sha256(main_surface.tobytes()).hexdigest(),
        "pixels": table,
    }
    with open(os.path.join(HERE, "c1_default_640.json"), "w") as f:
        json.dump(c1_doc, f, indent=1)

    # ---- C2-shaped small frame: default scene, 16:9, depth 8 --------------------------------------------
    cam2 = S.default_camera(image_width=160, aspect_ratio=16.0 / 9.0)
    frame_fixture(ref, scene, ref.camera_init(cam2), 8, os.path.join(HERE, "default_160x90_d8.npz"))

    # ---- C3-shaped small frame: synthetic 10k spheres + 64 walls, depth 10 -----------------------------------
    syn = S.synthetic_scene()
    cam3 = S.default_camera(image_width=96, aspect_ratio=16.0 / 9.0)
    r3 = frame_fixture(ref, syn, ref.camera_init(cam3), 10, os.path.join(HERE, "synthetic_96x54_d10.npz"))
    # a cyclic band subset of the survey's 384x216 probe frame (every 6th 4-row band)
    cam3b = S.default_camera(image_width=384, aspect_ratio=16.0 / 9.0)
    pod3b = ref.camera_init(cam3b)
    rows = np.array([r for r in range(pod3b.height) if (r // 4) % 6 == 1], dtype=np.int32)
    frame_fixture(ref, syn, pod3b, 10, os.path.join(HERE, "synthetic_384x216_bands_d10.npz"), rows=rows)

    # ---- C5-shaped: four fly-through frames (incl. wall back faces and occluded centres) --------------------
    cams = S.flythrough_cameras(256, image_width=96, aspect_ratio=16.0 / 9.0)
    for k in (0, 48, 128, 224):
        frame_fixture(ref, scene, ref.camera_init(cams[k]), 10, os.path.join(HERE, "flythrough_96x54_k%03d.npz" % k))

    # ---- function-level known answers ---------------------------------------------------------------------
    kat = {"source": "oracle/_ref (unmodified reference)", "intersect": [], "closest": [], "trace": [], "out_color": [],
           "reflect": [], "shading": [], "camera": [], "quantise": {}}

    def rv(lo, hi):
        return (rng.uniform(lo, hi), rng.uniform(lo, hi), rng.uniform(lo, hi))

    # fixed cases from SURVEY.md §8(c) first, then random rays against default and synthetic objects
    fixed = [(scene[0], (0, 0, 0), (1, .2, .1)), (scene[1], (0, 0, 0), (1, .8, .1)), (scene[1], (0, 0, 0), (1, .9, .1)),
             (scene[2], (0, 0, 0), (1, -.8, .3)), (scene[1], (2.4, 5, .4), (.1, -1, .05)),
             (scene[0], (1.5, 0, 0), (1, 0, 0)),                       # origin inside the sphere: no hit from inside
             (scene[0], (0, .5, 0), (1, 0, 0)),                        # grazing, det ~ 0
             (S.Wall(S.Material((1, 1, 1)), (1, 0, 0), (0, 0, 1), 1, 1), (0, 0, 1), (1, 0, -1)),  # normal || z: NaN basis
             (S.Wall(), (0, 0, 0), (1, 0, 0)), (S.Sphere(), (3, 0, 0), (-1, 0, 0))]               # DEFAULT_MAT ctor paths
    cases = list(fixed)
    for _ in range(40):
        g = rng.choice(scene)
        cases.append((g, rv(-1, 1), rv(-1, 1)))
    for _ in range(40):
        g = syn[rng.randrange(len(syn))]
        o = rv(-2, 2)
        p = g.center if g.kind == 0 else g.position
        aim = [p[k] - o[k] + rng.uniform(-.4, .4) for k in range(3)]
        cases.append((g, o, tuple(aim)))
    for g, o, d in cases:
        dist, nrm, hit = ref.intersect(g, o, d)
        pod_g = g.pod()
        kat["intersect"].append({
            "kind": g.kind, "p": hx3(pod_g.p.tuple()), "n": hx3(pod_g.n.tuple()), "a": hx(pod_g.a), "b": hx(pod_g.b),
            "o": hx3(o), "d": hx3(d), "distance": hx(dist), "normal": hx3(nrm), "hit": hit})
    for _ in range(30):
        o, d = rv(-1, 1), (rng.uniform(.2, 1), rng.uniform(-1, 1), rng.uniform(-.5, .8))
        dist, nrm, idx = ref.find_closest_hit(scene, o, d)
        kat["closest"].append({"scene": "default", "o": hx3(o), "d": hx3(d), "distance": hx(dist), "normal": hx3(nrm), "index": idx})
        rgb = ref.trace_ray(scene, o, d, 10)
        kat["trace"].append({"scene": "default", "o": hx3(o), "d": hx3(d), "depth": 10, "rgb": hx3(rgb)})
    for v in [(1, 0, .5), (1, 0, -.5), (1, 0, 0), (0, 0, 1), (0, 0, 0)] + [rv(-1, 1) for _ in range(20)]:
        kat["out_color"].append({"v": hx3(v), "rgb": hx3(ref.out_color(v))})
    for v, n in [((1, .2, .1), (-.5, .1, .05))] + [(rv(-1, 1), rv(-1, 1)) for _ in range(20)]:
        kat["reflect"].append({"v": hx3(v), "n": hx3(n), "out": hx3(ref.reflect(v, n))})
    for pos, n, view in [((1, .2, .1), (-.5, .1, .05), (-1, -.2, -.1))] + [(rv(-2, 2), rv(-1, 1), rv(-1, 1)) for _ in range(20)]:
        kat["shading"].append({"pos": hx3(pos), "n": hx3(n), "view": hx3(view),
                               "diffuse": hx(ref.diffuse(pos, n)), "specular": hx(ref.specular(pos, n, view))})
    cam_cases = [S.default_camera(), S.default_camera(1920, 16.0 / 9.0), S.default_camera(3840, 16.0 / 9.0),
                 S.default_camera(7680, 16.0 / 9.0), S.default_camera(641, 1.5)] + [cams[k] for k in (0, 17, 48, 128, 224)]
    for c in cam_cases:
        p = ref.camera_init(c)
        kat["camera"].append({"position": hx3(c.position), "lookat": hx3(c.lookat), "vup": hx3(c.vup), "vfov": hx(c.vfov),
                              "aspect_ratio": hx(c.aspect_ratio), "image_width": hx(c.image_width),
                              "image_top_left": hx3(p.image_top_left.tuple()), "delta_x": hx3(p.delta_x.tuple()),
                              "delta_y": hx3(p.delta_y.tuple()), "width": p.width, "height": p.height})
    qin = [0.0, 1.0, 0.5, 0.999999, 1.0 / 255, 339.0 / 255, 1.33277, -0.2, -51.0 / 255, 256.0 / 255, 3.0, 1e6, -1e6, 1e12,
           -1e12, float("nan"), float("inf"), float("-inf"), 8421504.7, 2147483647.0 / 255, 2147483648.0 / 255, 1e-9, -1e-9]
    qin += [rng.uniform(-2, 3) for _ in range(64)]
    while len(qin) % 3:
        qin.append(0.25)
    qa = np.array(qin, dtype=np.float64).reshape(-1, 3)
    kat["quantise"] = {"rgb": [hx(x) for x in qa.ravel()], "rgba8": [int(x) for x in ref.quantise(qa)]}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---- the survey's own vectors, for the record; checked by tests/test_oracle.py ----
    survey = {
        "source": "SURVEY.md §8(c) and A.2: captured by the survey from the unmodified reference",
        "c1": {"id_histogram": {"-1": 333324, "0": 40280, "1": 14151, "2": 21845}, "total_rays": 486688,
               "rgb8_sha256": "e9c24e4b1a27fb8cffada0c7f8648edf5ab88cf153f2826ca33f7f410a77c3cd",
               "pixels": [[0, 0, -1, None, [0.16826695192825547, 0.24083667483082413, 0.50027889161027472], [42, 61, 127]],
                          [319, 320, 0, 1.0000073126433018, [0.14730342986288231, 0.88910188201527995, 0.27311033813195718], [37, 226, 69]],
                          [320, 320, 0, 1.0000073126433018, [0.012500000000000001, 0.72477086761940834, 0.037499999999999999], [3, 184, 9]],
                          [300, 0, 2, 3.0070885078880427, [0.12988983590042288, 0.53786185722625734, 0.26677812214560831], [33, 137, 68]],
                          [300, 639, 1, 2.0047256719253612, [0.12988983590042288, 0.17033436643682495, 0.63430561293504073], [33, 43, 161]],
                          [250, 250, 0, 1.1966790504503475, [0.080155023357626975, 0.37965398812348733, 0.24869273576640979], [20, 96, 63]],
                          [260, 420, -1, None, [0.21780005794232343, 0.29487279048253467, 0.51829093016084482], [55, 75, 132]],
                          [400, 320, 0, 1.1137433615022749, [0.012500000000000001, 0.38190111297048546, 0.037499999999999999], [3, 97, 9]],
                          [215, 320, 0, 1.243018237786498, [0.073621757544865774, 0.32363688412193209, 0.24631700274358753], [18, 82, 62]],
                          [639, 639, -1, None, [0.025, 0.05, 0.075], [6, 12, 19]]]},
        "synthetic": {"sphere0": {"center": [58.07738439454419, -21.911292446286652, 5.0785864877254099], "radius": 0.43230631669471309,
                                  "color": [0.82487999305303117, 0.91969074258959527, 0.44755495920589849], "metallic": 0.72066738682676201},
                      "wall0": {"position": [37.269910916545875, -3.1405171796379889, -7.2978479189889995], "phi": 1.8409633588487833,
                                "nz": -0.25338659597007884, "length": 5.4004321898364713, "width": 3.3380435017516565,
                                "color": [0.11199935765508733, 0.98415786595095878, 0.10857096580500576], "metallic": 0.57260323870629082},
                      "probe_384x216_d10": {"total_rays": 187615, "chain_histogram": {"1": 29975, "2": 28271, "3": 12106, "4": 5836, "5": 3153,
                                                                                      "6": 1659, "7": 900, "8": 462, "9": 271, "10": 139, "11": 172}}},
    }
    with open(os.path.join(HERE, "survey_vectors.json"), "w") as f:
        json.dump(survey, f, indent=1)
    print("synthetic 96x54 rays:", r3["total_rays"])
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
