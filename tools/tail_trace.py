#!/usr/bin/env python
"""Developer tool (GPU box): where does the drain tail of trace_kernel go?

Needs a library built with -DRTX_TAIL_TRACE (tools/build_variant.sh tail RTX_TAIL_TRACE=1). Renders the 4K frame of the
10k-object scene once, pulls the per-warp log of every scan segment executed after the pixel pool ran dry, and prints
  * per segment ordinal (1st, 2nd ... scan after dry): warps still active, mean live chains, share in cooperative mode,
    mean / max duration of the scan part and of the drain + shading part,
  * the distribution of warp finish times relative to the moment the pool ran dry.
    RTX_B200_LIB=.../librtx_b200_tail.so python tools/tail_trace.py [width]
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ray-tracer-from-scratch_b200")
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
S, abi = pkg.scene, pkg.abi

width = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
lib = R.load_library(R.LIB_PATH)
r = R.Renderer(0)
r.set_scene(S.synthetic_scene())
pod = S.default_camera(width, 16.0 / 9.0).pod()
import torch
frame = torch.empty((pod.height, pod.width), dtype=torch.int32, device="cuda")
o = abi.Outputs()
o.memory, o.rgba8 = abi.RTX_MEM_DEVICE, frame.data_ptr()
p = R.default_params(max_depth=10)
for _ in range(2):
    lib.rtx_debug_tail_clear()
    st = r.render_raw([pod], p, o)
NREC = 32
log = np.zeros(160 * 32 * NREC * 4, np.uint64)
dry = C.c_uint64()
assert lib.rtx_debug_tail_log(log.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(dry)) == 0
log = log.reshape(160, 32, NREC, 4)[:148, :16]
dry = dry.value
print("raytracing %.3f ms, drain %.3f ms, exit spread %.3f ms" % (st.raytracing_ms, st.drain_ms, st.exit_spread_ms))
t0 = log[..., 0].astype(np.float64)
valid = log[..., 0] > 0
live = (log[..., 1] & 0xFF).astype(np.int64)
coop = ((log[..., 1] >> 8) & 1).astype(np.int64)
scan = (log[..., 2].astype(np.float64) - t0) * 1e-3
rest = (log[..., 3].astype(np.float64) - log[..., 2].astype(np.float64)) * 1e-3
start = (t0 - float(dry)) * 1e-3
end = (log[..., 3].astype(np.float64) - float(dry)) * 1e-3
print("seg  warps  live(mean)  coop%%   start us(mean)  scan us mean/max   drain+shade us mean/max   end us mean/max")
for k in range(NREC):
    v = valid[..., k]
    if not v.any():
        break
    print("%3d  %5d  %9.1f  %5.1f  %12.1f   %8.1f %8.1f   %10.1f %8.1f   %8.1f %8.1f" % (
        k, v.sum(), live[..., k][v].mean(), 100.0 * coop[..., k][v].mean(), start[..., k][v].mean(),
        scan[..., k][v].mean(), scan[..., k][v].max(), rest[..., k][v].mean(), rest[..., k][v].max(),
        end[..., k][v].mean(), end[..., k][v].max()))
last = np.where(valid, end, 0).max(axis=-1)       # per warp: end of its last logged segment
print("warp finish after dry (us): mean %.1f  p50 %.1f  p90 %.1f  p99 %.1f  max %.1f" % (
    last.mean(), np.percentile(last, 50), np.percentile(last, 90), np.percentile(last, 99), last.max()))
per_sm = last.max(axis=1)
print("SM finish after dry (us): mean %.1f  min %.1f  max %.1f" % (per_sm.mean(), per_sm.min(), per_sm.max()))
tot_scan = np.where(valid, scan, 0).sum()
tot_rest = np.where(valid, rest, 0).sum()
print("warp-time after dry: scan %.1f ms, drain+shade %.1f ms (sum over %d warps)" % (tot_scan * 1e-3, tot_rest * 1e-3, 148 * 16))
