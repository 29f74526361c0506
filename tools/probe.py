#!/usr/bin/env python
"""Developer probe (GPU box): FP32 peak microbenchmark and raw kernel timings of the benchmark configurations."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray-tracer-from-scratch_b200")
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
S = pkg.scene


def main():
    out = {}
    r = R.Renderer(0)
    for v, name in ((0, "ffma_scalar"), (1, "ffma_packed_f32x2"), (2, "ffma2_3distinct"), (3, "ffma2_shared_b"), (4, "ffma2_2distinct"), (5, "mixed_8ffma2_4ffma"), (6, "mixed_8ffma2_8ffma"), (7, "hfma2_f16x2")):
        t, mhz = r.ffma_peak(v)
        out[name] = {"tflops": t, "sm_mhz_est": mhz}
        print(name, "%.2f TFLOP/s" % t, "~%.0f MHz" % mhz, flush=True)
    which = sys.argv[1:] or ["c2", "c3"]
    if "c2" in which:
        r.set_scene(S.default_scene())
        pod = S.default_camera(1920, 16.0 / 9.0).pod()
        for it in range(4):
            _, st = r.render([pod], R.default_params(max_depth=8), want=("rgba8",))
            print("c2", st.as_dict(), flush=True)
        out["c2"] = st.as_dict()
    if "c5" in which:
        r.set_scene(S.default_scene())
        pods = [c.pod() for c in S.flythrough_cameras(256, 1920, 16.0 / 9.0)]
        import torch
        buf = torch.empty((256, 1080, 1920), dtype=torch.int32, device="cuda")
        o = pkg.abi.Outputs()
        o.memory, o.rgba8 = pkg.abi.RTX_MEM_DEVICE, buf.data_ptr()
        for it in range(3):
            st = r.render_raw(pods, R.default_params(max_depth=10), o)
            print("c5", st.as_dict(), flush=True)
        out["c5"] = st.as_dict()
    for name, ns in (("big12k", 12500), ("big20k", 20000), ("big40k", 40000)):
        if name in which:
            r.set_scene(S.synthetic_scene(ns, 64))
            pod = S.default_camera(1920, 16.0 / 9.0).pod()
            for it in range(3):
                _, st = r.render([pod], R.default_params(max_depth=10), want=("rgba8",))
            d = st.as_dict()
            d["tflops_algorithmic"] = (st.sphere_tests * 20 + st.wall_tests * 33) / (st.raytracing_ms * 1e-3) / 1e12
            d["mrays_s"] = st.total_rays / (st.raytracing_ms * 1e-3) / 1e6
            print(name, d, flush=True)
            out[name] = d
    if "c3" in which or "c3small" in which:
        t0 = time.time()
        syn = S.synthetic_scene()
        r.set_scene(syn)
        print("synthetic scene built+uploaded in %.1fs" % (time.time() - t0), flush=True)
        w = 3840 if "c3" in which else 960
        pod = S.default_camera(w, 16.0 / 9.0).pod()
        for it in range(3):
            _, st = r.render([pod], R.default_params(max_depth=10), want=("rgba8",))
            d = st.as_dict()
            flops = st.sphere_tests * 20 + st.wall_tests * 33
            d["tflops_algorithmic"] = flops / (st.raytracing_ms * 1e-3) / 1e12
            d["mrays_s"] = st.total_rays / (st.raytracing_ms * 1e-3) / 1e6
            print("c3", d, flush=True)
        out["c3"] = d
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
