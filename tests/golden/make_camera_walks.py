#!/usr/bin/env python
"""Generates camera_walks.json FROM THE UNMODIFIED REFERENCE (oracle/_ref: Camera::init + forward/backward/left/right/
rotate_left_right/rotate_up_down of /root/reference/scene.cpp:80-165, driven by oracle/ref_harness.cpp::ref_camera_walk).

    python tests/golden/make_camera_walks.py

Each walk: the camera inputs, the steps (op, arg) and position / direction / vup after every step as hex doubles.
"""
import importlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding as ob  # noqa: E402

S = importlib.import_module("ray-tracer-from-scratch_b200").scene


def main():
    ref = ob.load_reference()
    rng = random.Random(0xCA3)
    walks = []
    cams = [("default", S.default_camera(640, 1.0))] + [("orbit%d" % k, S.flythrough_cameras(256, 1920, 16.0 / 9.0)[k]) for k in (7, 100, 201)]
    scripted = {
        "default": [("w", 0), ("w", 0), ("d", 0), ("s", 0), ("a", 0), ("y", 0.05), ("d", 0), ("p", -0.05), ("w", 0), ("y", -0.3), ("a", 0),
                    ("p", 1.2), ("p", 1.2), ("w", 0), ("p", -2.0), ("p", -2.0), ("d", 0), ("y", 3.5), ("s", 0)],
    }
    for name, cam in cams:
        steps = scripted.get(name)
        if steps is None:
            steps = []
            for _ in range(24):
                op = rng.choice("wsadyp")
                steps.append((op, round(rng.uniform(-2.0, 2.0), 3) if op in "yp" else 0))
        states, pod = ref.camera_walk(cam, steps)
        walks.append({
            "name": name,
            "camera": {"position": [float(x).hex() for x in cam.position], "lookat": [float(x).hex() for x in cam.lookat],
                       "vup": [float(x).hex() for x in cam.vup], "vfov": cam.vfov, "aspect_ratio": float(cam.aspect_ratio).hex(),
                       "image_width": cam.image_width},
            "steps": [[op, float(arg)] for op, arg in steps],
            "states": [[[float(x).hex() for x in vec] for vec in st] for st in states],
            "final_pod": {"position": [float(x).hex() for x in (pod.position.x, pod.position.y, pod.position.z)],
                          "image_top_left": [float(x).hex() for x in (pod.image_top_left.x, pod.image_top_left.y, pod.image_top_left.z)],
                          "width": pod.width, "height": pod.height},
        })
    with open(os.path.join(HERE, "camera_walks.json"), "w") as f:
        json.dump({"source": "oracle/_ref (unmodified reference), ref_camera_walk", "walks": walks}, f, indent=1)
    print("wrote camera_walks.json:", len(walks), "walks")


if __name__ == "__main__":
    main()
