#!/usr/bin/env python
"""Full-size golden statistics (BASELINE.json sizes) FROM THE UNMODIFIED REFERENCE, small enough to commit.

For config C3 (3840x2160, synthetic 10k scene, depth 10) the whole frame is rendered on the CPU:
  * RGBA8 surface words by oracle/_ref (the reference's recursive_ray_tracing + its quantise),
  * ray counts and primary object ids by the C port (bit-identical to the reference on every case both have
    been run on — tests/test_oracle.py — and about twice as fast, which matters for 1.9e11 object tests),
and reduced to per-row CRC32s / per-row ray totals. For C4 (7680x4320) every 16th 4-row band is rendered.
The GPU suite recomputes the same reductions from the CUDA frame (tests/test_gpu_fullsize.py).

For C5 (the 256-frame 1080p orbit of the default scene) three whole frames are rendered: k = 48 and k = 224 (a wall
occludes the sphere) and k = 128 (the far side of the orbit, where the walls are seen from BEHIND: the back-face
pass-through of main.cpp:111-113).

    python tests/golden/make_fullsize.py [c3] [c4] [c2] [c5]      # ~25 min (c3) + ~8 min (c4) on 8 cores; c5 seconds
"""
import importlib
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

S = importlib.import_module("ray-tracer-from-scratch_b200").scene


def crc_rows(a):
    return [zlib.crc32(np.ascontiguousarray(row).tobytes()) for row in a]


def run(name, width, depth, scene, rows, chunk=64, cam=None, extra=None):
    ref, port = ob.load_reference(), ob.load_port()
    pod = ref.camera_init(cam if cam is not None else S.default_camera(width, 16.0 / 9.0))
    objs = S.flatten(scene)
    out = {"source": "rgba8: oracle/_ref (unmodified reference); ray_count/object_id: oracle.c (pinned to the reference)",
           "width": pod.width, "height": pod.height, "depth": depth, "rows": [int(r) for r in rows],
           "rgba8_crc": [], "ray_count_crc": [], "object_id_crc": [], "row_rays": [], "row_over_range": []}
    t0 = time.time()
    for k in range(0, len(rows), chunk):
        rr = np.asarray(rows[k:k + chunk], dtype=np.int32)
        a = ref.render(objs, pod, depth, rows=rr, want=("radiance", "rgba8"))
        b = port.render(objs, pod, depth, rows=rr, want=("rgba8", "ray_count", "object_id"))
        assert np.array_equal(a["rgba8"], b["rgba8"]), "port and reference disagree"
        out["rgba8_crc"] += crc_rows(a["rgba8"])
        out["ray_count_crc"] += crc_rows(b["ray_count"])
        out["object_id_crc"] += crc_rows(b["object_id"])
        out["row_rays"] += [int(x) for x in b["ray_count"].astype(np.int64).sum(axis=1)]
        v = a["radiance"] * 255
        out["row_over_range"] += [int(x) for x in (~((v >= 0) & (v < 256)).all(axis=-1)).sum(axis=1)]
        print("%s: %d/%d rows, %.0f s" % (name, k + len(rr), len(rows), time.time() - t0), flush=True)
    out["total_rays"] = int(sum(out["row_rays"]))
    if extra is not None:
        out.update(extra)
        return out
    with open(os.path.join(HERE, "fullsize_%s.json" % name), "w") as f:
        json.dump(out, f)
    print(name, "total rays", out["total_rays"], "in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4"]
    if "c2" in which:
        run("c2", 1920, 8, S.default_scene(), list(range(1080)), chunk=270)
    if "c3" in which:
        run("c3", 3840, 10, S.synthetic_scene(), list(range(2160)), chunk=48)
    if "c5" in which:
        cams = S.flythrough_cameras(256, 1920, 16.0 / 9.0)
        frames = {}
        for k in (48, 128, 224):
            frames[str(k)] = run("c5[%d]" % k, 1920, 10, S.default_scene(), list(range(1080)), chunk=270, cam=cams[k], extra={"frame": k})
        with open(os.path.join(HERE, "fullsize_c5.json"), "w") as f:
            json.dump({"source": frames["48"]["source"], "n_frames": 256, "width": 1920, "height": 1080, "depth": 10, "frames": frames}, f)
        print("c5 rays", {k: v["total_rays"] for k, v in frames.items()})
    if "c4" in which:
        run("c4", 7680, 10, S.synthetic_scene(), [r for r in range(4320) if (r // 4) % 16 == 5], chunk=24)
