#!/usr/bin/env python
"""Developer probe (GPU box): the HBM-side kernels alone — standalone quantise and the tone-map extension on a
device-resident 8K radiance frame, L2 flushed before every launch. Prints ms and GB/s (algorithmic bytes)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray-tracer-from-scratch_b200")
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
abi = pkg.abi


def main():
    W, H = (7680, 4320) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
    npx = W * H
    dev = torch.device("cuda:0")
    r = R.Renderer(0)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    out = torch.empty(npx, dtype=torch.int32, device=dev)
    ptm = R.default_params(tonemap=abi.RTX_TONEMAP_REINHARD, quantise_mode=abi.RTX_QUANT_SATURATE)
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        rad = torch.rand(npx * 3, dtype=dt, device=dev) * 1.3
        eb = 4 if dt == torch.float32 else 8
        for kind, bpp, call in (("quantise", 3 * eb + 4, lambda: r.quantise_device(rad.data_ptr(), dt == torch.float32, npx, out.data_ptr())),
                                ("tonemap", 6 * eb + 4, lambda: r.tonemap_device(rad.data_ptr(), dt == torch.float32, npx, 1, ptm, out.data_ptr()))):
            ts = []
            for _ in range(5):
                flush.add_(1)
                torch.cuda.synchronize()
                ts.append(call().surface_update_ms)
            ms = sorted(ts[1:])[len(ts[1:]) // 2]
            print("%s %s %dx%d: %.4f ms, %.0f GB/s (%d B/px algorithmic)" % (kind, name, W, H, ms, npx * bpp / (ms * 1e-3) / 1e9, bpp), flush=True)
        del rad


if __name__ == "__main__":
    main()
