#!/usr/bin/env python
"""Randomized parity soak (GPU box): random scenes, cameras, depths, frame sizes and parameters, CUDA path vs the C
oracle. Ids, hit masks, ray counts and RGBA8 must be identical, radiance within 1e-12 relative.

    python tools/fuzz_parity.py [seconds=120] [seed=1]          (FUZZ_BIG=1: larger scenes and frames;
                                                                 FUZZ_EXT=1: also the extensions — boxes and the sun;
                                                                 FUZZ_ACCEL=1: half of the cases through the uniform grid,
                                                                 rtx_params.accel = RTX_ACCEL_GRID — same oracle, same bar)
"""
import importlib
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray-tracer-from-scratch_b200")
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
from oracle import binding as ob  # noqa: E402

S = pkg.scene


def random_scene(rng):
    scale = rng.choice([1.0, 1.0, 10.0, 100.0, 0.05])
    big = os.environ.get("FUZZ_BIG") == "1"
    n_s = rng.choice([17, 150, 400, 2000, 6000] if big else [0, 1, 3, 8, 15, 16, 17, 40, 150, 400])
    n_w = rng.choice([0, 2, 16, 30, 100] if big else [0, 0, 1, 2, 5, 16, 30])
    M = S.Material
    scene = []

    def mat():
        return M((rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0, 1)), rng.uniform(0, 1), rng.uniform(0, .3), rng.uniform(0, 1),
                 rng.uniform(0, 1), rng.choice([1, 2, 8, 50, 200.5]))
    for _ in range(n_s):
        c = (rng.uniform(-6, 12) * scale, rng.uniform(-8, 8) * scale, rng.uniform(-6, 6) * scale)
        scene.append(S.Sphere(mat(), c, rng.choice([0.05, 0.3, 1.0, 2.5, 2.5, 60.0 if rng.random() < 0.02 else 0.3]) * scale * rng.uniform(0.5, 1.5)))
    for _ in range(n_w):
        p = (rng.uniform(-6, 12) * scale, rng.uniform(-8, 8) * scale, rng.uniform(-6, 6) * scale)
        n = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.choice([0.0, 0.0, rng.uniform(-1, 1)]))
        if abs(n[0]) + abs(n[1]) < 1e-3:
            n = (1.0, 0.0, n[2])
        scene.append(S.Wall(mat(), p, n, rng.uniform(0.2, 8) * scale, rng.uniform(0.2, 8) * scale))
    if os.environ.get("FUZZ_EXT") == "1":
        for _ in range(rng.choice([0, 1, 2, 5, 12, 40])):
            p = (rng.uniform(-6, 12) * scale, rng.uniform(-8, 8) * scale, rng.uniform(-6, 6) * scale)
            size = tuple(rng.choice([0.1, 0.5, 2.0, 6.0, 30.0]) * scale * rng.uniform(0.5, 1.5) for _ in range(3))
            scene.append(S.Box(mat(), p, size))
    rng.shuffle(scene)                      # spheres and walls interleaved in scene order
    if scene and rng.random() < 0.3:        # exact duplicates: ties must go to the lower index
        scene.insert(rng.randrange(len(scene)), scene[rng.randrange(len(scene))])
    return scene, scale


def random_camera(rng, scale):
    cam = S.Camera()
    cam.image_width = rng.choice([64, 97, 160, 320] if os.environ.get("FUZZ_BIG") == "1" else [1, 7, 33, 64, 97, 160])
    cam.aspect_ratio = rng.choice([1.0, 16.0 / 9.0, 0.5, 3.7])
    cam.vfov = rng.choice([20, 60, 90, 120])
    cam.position = (rng.uniform(-4, 10) * scale, rng.uniform(-6, 6) * scale, rng.uniform(-4, 4) * scale)
    off = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-.7, .7))
    f = rng.choice([1.0, 1.0, 0.5, 2.0])
    cam.lookat = tuple(cam.position[k] + off[k] * f for k in range(3))
    cam.vup = rng.choice([(0, 0, -1), (0, 0, 1), (0, 1, 0)])
    if int(cam.image_width / cam.aspect_ratio) < 1:
        cam.aspect_ratio = 1.0
    return cam


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    oracle = ob.load_port()
    r = R.Renderer(0)
    want = ("rgba8", "radiance_f64", "object_id", "hit_mask", "ray_count")
    t0, n, worst = time.time(), 0, 0.0
    while time.time() - t0 < seconds:
        scene, scale = random_scene(rng)
        cam = random_camera(rng, scale)
        pod = cam.pod()
        depth = rng.choice([0, 1, 3, 10, 10, 25])
        p = oracle.default_params()
        p.max_depth = depth
        kw = {}
        if rng.random() < 0.4:
            kw = dict(light_pos=(rng.uniform(-5, 5) * scale, rng.uniform(-5, 5) * scale, rng.uniform(-5, 5) * scale),
                      reflect_offset=rng.choice([1e-4, 1e-3, 1e-6]), sky_exponent=rng.choice([0.25, 0.5, 1.0, 2.2]))
            for k, v in kw.items():
                setattr(p, k, type(getattr(p, k))(*v) if isinstance(v, tuple) else v)
        if os.environ.get("FUZZ_EXT") == "1" and rng.random() < 0.5:
            ext = dict(sun_enabled=1, sun_color=(rng.uniform(0, 2), rng.uniform(0, 2), rng.uniform(0, 2)),
                       sun_direction=(rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1) or 0.5))
            for k, v in ext.items():
                setattr(p, k, type(getattr(p, k))(*v) if isinstance(v, tuple) else v)
            kw.update(ext)
        accel = 1 if os.environ.get("FUZZ_ACCEL") == "1" and rng.random() < 0.5 else 0
        r.set_scene(scene)
        got, st = r.render([pod], R.default_params(max_depth=depth, accel=accel, **kw), want=want)
        exp = oracle.render(scene, pod, params=p)
        tag = "case %d: %d objs, scale %g, %dx%d, depth %d, accel %d, %s" % (n, len(scene), scale, pod.width, pod.height, depth, accel, kw)
        for k, e in (("object_id", "object_id"), ("hit_mask", "hit_mask"), ("ray_count", "ray_count"), ("rgba8", "rgba8")):
            if not np.array_equal(got[k][0], exp[e]):
                bad = np.argwhere(got[k][0] != exp[e])
                print("MISMATCH", k, tag, "first at", bad[0], "count", len(bad))
                sys.exit(1)
        with np.errstate(invalid="ignore"):
            err = np.abs(got["radiance_f64"][0] - exp["radiance"]) / np.maximum(np.abs(exp["radiance"]), 1e-3)
        err = np.nanmax(err) if err.size else 0.0
        same_nan = np.array_equal(np.isnan(got["radiance_f64"][0]), np.isnan(exp["radiance"]))
        if not (err < 1e-12) or not same_nan or st.total_rays != exp["total_rays"]:
            print("MISMATCH radiance/rays", tag, err, same_nan, st.total_rays, exp["total_rays"])
            sys.exit(1)
        worst = max(worst, float(err))
        n += 1
    print("fuzz: %d random cases identical (ids, masks, ray counts, RGBA8); worst radiance rel err %.2e; %.0f s" % (n, worst, time.time() - t0))


if __name__ == "__main__":
    main()
