"""CPU suite, part 3: the multi-GPU host logic with world_size 2 over gloo (no GPU).

Each rank renders its cyclic row bands (with the oracle standing in for the kernel — this is a test of the
sharding/gather plumbing, not of compute), the product's all_gather_blocks() gathers them, and rank 0 reassembles
the frame with the band map the library exports (rtx_local_rows / rtx_global_row). Frame sharding (config C5) is
checked the same way.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, band, result_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("ray-tracer-from-scratch_b200")
    R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
    SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
    from oracle import binding as ob
    S, oracle = pkg.scene, ob.load_port()
    scene = S.synthetic_scene(200, 8, seed=11)
    pod = S.default_camera(W, W / H).pod()
    assert pod.height == H

    # ---- one frame, cyclic row bands (config C4's shape) ----
    rows = R.global_rows(H, band, world, rank)
    rpr = SH.rows_per_rank(H, band, world)
    mine = oracle.render(scene, pod, 6, rows=rows, want=("rgba8", "object_id"))
    local = torch.zeros((rpr, W), dtype=torch.int32)
    local[: len(rows)] = torch.from_numpy(mine["rgba8"].view(np.int32))
    gathered = SH.all_gather_blocks(local, world)
    total_rays = torch.tensor([mine["total_rays"]], dtype=torch.int64)
    dist.all_reduce(total_rays)
    ok = True
    if rank == 0:
        full = oracle.render(scene, pod, 6, want=("rgba8", "ray_count"))
        frame = np.zeros((H, W), np.uint32)
        for r in range(world):
            g = R.global_rows(H, band, world, r)
            frame[g] = gathered[r][: len(g)].numpy().view(np.uint32)
        ok = ok and np.array_equal(frame, full["rgba8"]) and int(total_rays.item()) == full["total_rays"]

    # ---- extension: row-sharded tone map = local sums, ONE integer all-reduce, local apply (sharding.allreduce_tonemap_sums) ----
    p = oracle.default_params()
    p.tonemap, p.quantise_mode, p.tonemap_white = pkg.abi.RTX_TONEMAP_REINHARD, pkg.abi.RTX_QUANT_SATURATE, 3.0
    rad_mine = oracle.render(scene, pod, 6, rows=rows, want=("radiance",))["radiance"].reshape(1, -1, 3)
    sums = torch.zeros(1, dtype=torch.int64)
    oracle.tonemap_sums(rad_mine, sums.numpy())
    SH.allreduce_tonemap_sums(sums, world)
    tm_local = torch.zeros((rpr, W), dtype=torch.int32)
    tm_local[: len(rows)] = torch.from_numpy(oracle.tonemap_apply(rad_mine, sums.numpy(), H * W, p).reshape(len(rows), W).view(np.int32))
    tm_gathered = SH.all_gather_blocks(tm_local, world)
    if rank == 0:
        rad_full = oracle.render(scene, pod, 6, want=("radiance",))["radiance"].reshape(1, -1, 3)
        exp_tm, _ = oracle.tonemap(rad_full, p)
        tm = np.zeros((H, W), np.uint32)
        for r in range(world):
            g = R.global_rows(H, band, world, r)
            tm[g] = tm_gathered[r][: len(g)].numpy().view(np.uint32)
        ok = ok and np.array_equal(tm, exp_tm.reshape(H, W))          # bit for bit: the integer statistic does not depend on the split

    # ---- camera path, frames sharded (config C5's shape) ----
    cams = [c.pod() for c in S.flythrough_cameras(256, 32, 16.0 / 9.0)[::51]]     # 6 frames, ragged over 2 ranks? 6/2 = 3 each
    cams = cams[:5]                                                               # 5 frames: ragged
    owner = SH.frame_owner(len(cams), world)
    per_rank = (len(cams) + world - 1) // world
    h, w = cams[0].height, cams[0].width
    loc = torch.zeros((per_rank, h, w), dtype=torch.int32)
    dscene = S.default_scene()
    for k, f in enumerate(owner[rank]):
        loc[k] = torch.from_numpy(oracle.render(dscene, cams[f], 10, want=("rgba8",))["rgba8"].view(np.int32))
    g = SH.all_gather_blocks(loc, world)
    if rank == 0:
        frames = g.permute(1, 0, 2, 3).reshape(per_rank * world, h, w)[: len(cams)]
        for f, pod_f in enumerate(cams):
            exp = oracle.render(dscene, pod_f, 10, want=("rgba8",))["rgba8"]
            ok = ok and np.array_equal(frames[f].numpy().view(np.uint32), exp)
        with open(result_path, "w") as fh:
            fh.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("H,W,band", [(48, 64, 4), (50, 40, 3)])
def test_two_rank_band_gather_and_frame_sharding(tmp_path, H, W, band, port):
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), H, W, band, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"


def test_frame_owner_and_rows_per_rank(pkg):
    SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
    assert SH.frame_owner(5, 2) == [[0, 2, 4], [1, 3]]
    assert SH.frame_owner(256, 8)[3][:3] == [3, 11, 19] and all(len(x) == 32 for x in SH.frame_owner(256, 8))
    assert SH.rows_per_rank(4320, 4, 8) == 540 and SH.rows_per_rank(2160, 4, 8) == 272 and SH.rows_per_rank(50, 3, 4) == 14
