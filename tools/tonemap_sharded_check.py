#!/usr/bin/env python
"""Multi-process check (GPU box, under torchrun): the row-sharded tone map over NCCL against the single-GPU operator.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/tonemap_sharded_check.py

Every rank traces its cyclic row bands of one frame into a device radiance buffer, `sharding.tonemap_sharded` does
local sums -> int64 all-reduce (NCCL SUM) -> local map + pack, one all-gather + rtx_unpermute_bands assembles the frame
on rank 0, which then renders the whole frame alone with rtx_params.tonemap and compares bit for bit.
"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray-tracer-from-scratch_b200")
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
abi, S = pkg.abi, pkg.scene


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    r = R.Renderer(local)
    scene = S.synthetic_scene(2000, 16, seed=3)
    r.set_scene(scene)
    pod = S.default_camera(640, 16.0 / 9.0).pod()
    H, W, band = pod.height, pod.width, 4
    tm = dict(tonemap=abi.RTX_TONEMAP_REINHARD, quantise_mode=abi.RTX_QUANT_SATURATE, tonemap_white=3.0, sun_enabled=1)
    rows = R.local_rows(H, band, world, rank)
    rpr = SH.rows_per_rank(H, band, world)
    rad = torch.zeros((1, rows, W, 3), dtype=torch.float64, device=dev)
    o = abi.Outputs()
    o.memory, o.radiance_f64 = abi.RTX_MEM_DEVICE, rad.data_ptr()
    torch.cuda.synchronize()
    r.render_raw([pod], R.default_params(max_depth=6, sun_enabled=1, band_rows=band, n_ranks=world, rank=rank), o)
    out, sums = SH.tonemap_sharded(r, rad, H * W, R.default_params(**tm), world)
    packed = torch.zeros((rpr, W), dtype=torch.int32, device=dev)
    packed[:rows] = out[0]
    gathered = SH.all_gather_blocks(packed, world)
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    if rank == 0:
        frame = torch.empty((H, W), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        r.unpermute_bands(gathered.data_ptr(), frame.data_ptr(), H, W, 4, band, world, rpr)
        whole = torch.empty((1, H, W), dtype=torch.int32, device=dev)
        o2 = abi.Outputs()
        o2.memory, o2.rgba8 = abi.RTX_MEM_DEVICE, whole.data_ptr()
        torch.cuda.synchronize()
        r.render_raw([pod], R.default_params(max_depth=6, **tm), o2)
        same = bool(torch.equal(frame, whole[0]))
        ok[0] = 1 if same else 0
        print("tonemap_sharded over NCCL, %d ranks, %dx%d: %s (sum %d)" % (world, W, H, "identical to the single-GPU frame" if same else "MISMATCH",
                                                                         int(sums[0].item())), flush=True)
    dist.broadcast(ok, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
