#!/usr/bin/env python
"""Developer probe (GPU box): what the scheduling hint (rtx_params.pixel_order) does to the drain tail.
C3 (4K) and C4 (8K) frames of the synthetic scene, whole and as one rank's rows of an 8-way split: scan order against
cost order, the same camera repeated and a camera that moves between frames. RTX_B200_LIB selects the library."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray-tracer-from-scratch_b200")
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
S = pkg.scene
abi = pkg.abi
HAS_ORDER = hasattr(abi, "RTX_ORDER_COST") and os.environ.get("PROBE_NO_ORDER") != "1"


def moving(width, k, step=(0.15, 0.05, 0.0)):
    cam = S.default_camera(width, 16.0 / 9.0)
    cam.position = tuple(c + s * k for c, s in zip(cam.position, step))
    cam.init()
    return cam.pod()


def run(r, tag, pods, **kw):
    line = []
    for pod in pods:
        _, st = r.render([pod], R.default_params(max_depth=10, **kw), want=("rgba8",))
        line.append("%.3f/%.3f" % (st.raytracing_ms, st.drain_ms))
    print("%-34s raytracing/drain ms: %s" % (tag, "  ".join(line)), flush=True)


def main():
    r = R.Renderer(0)
    r.set_scene(S.synthetic_scene())
    orders = (("scan", abi.RTX_ORDER_SCAN), ("cost", abi.RTX_ORDER_COST)) if HAS_ORDER else (("prev", None),)
    for width in [int(w) for w in (sys.argv[1:] or ["3840", "7680"])]:
        same = [S.default_camera(width, 16.0 / 9.0).pod()] * 5
        walk = [moving(width, k) for k in range(5)]
        key, fwd = [], []                                                  # the reference's own key moves (main.cpp:262-306, init() not re-run)
        cam_r, cam_f = S.default_camera(width, 16.0 / 9.0), S.default_camera(width, 16.0 / 9.0)
        for k in range(5):
            key.append(cam_r.pod())
            fwd.append(cam_f.pod())
            cam_r.right()
            cam_f.forward()
        for name, order in orders:
            kw = {} if order is None else {"pixel_order": order}
            run(r, "%d %s same camera" % (width, name), same, **kw)
            if os.environ.get("PROBE_QUICK") == "1":
                continue
            run(r, "%d %s moving camera" % (width, name), walk, **kw)
            run(r, "%d %s key 'd' once per frame" % (width, name), key, **kw)
            run(r, "%d %s key 'w' once per frame" % (width, name), fwd, **kw)
            run(r, "%d %s same camera, rank 0 of 8" % (width, name), same, band_rows=4, n_ranks=8, rank=0, **kw)
    r.close()


if __name__ == "__main__":
    main()
