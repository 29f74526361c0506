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


class StandInRenderer:
    """Host-side stand-in for renderer.Renderer in the gloo tests of ShardedRenderer: same methods, same placement rules
    (cyclic bands, frame_offset / frame_stride, at most RTX_MAX_IN_FLIGHT calls in flight), pixels from the oracle instead of the
    kernels. It exists to exercise the PLUMBING on a machine without a GPU: the shared host frame every rank maps, the
    chunked render/copy pipeline, the barriers, the open/close order. It is test code; the product has no such path."""

    def __init__(self, oracle, scene, abi, R):
        import mmap
        self.device, self.oracle, self.scene, self.abi, self.R, self.mmap = 0, oracle, scene, abi, R, mmap
        self.maps, self.queue = {}, []
        self.calls = {"render_raw": 0, "render_async": 0, "max_in_flight": 0}

    def set_stream(self, handle):
        pass

    def host_shared_open(self, name, nbytes, create):
        import ctypes
        fd = os.open("/dev/shm" + name, (os.O_CREAT | os.O_RDWR) if create else os.O_RDWR, 0o600)
        if create:
            os.ftruncate(fd, nbytes)
        mm = self.mmap.mmap(fd, nbytes)
        os.close(fd)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
        self.maps[addr] = mm
        return addr

    def host_device_pointer(self, ptr):
        return ptr

    def host_shared_close(self, ptr, unlink_name=None):
        self.maps.pop(ptr)          # the mapping itself is released with the process (numpy views may still reference it)

    def _render(self, cams, p, o):
        import ctypes
        st = self.abi.Stats()
        H, W = cams[0].height, cams[0].width
        total_frames_words = ctypes.cast(o.frame_rgba8, ctypes.POINTER(ctypes.c_uint32))
        for k, cam in enumerate(cams):
            rows = self.R.global_rows(H, p.band_rows, p.n_ranks, p.rank) if p.n_ranks > 1 else np.arange(H, dtype=np.int32)
            out = self.oracle.render(self.scene, cam, p.max_depth, rows=rows, want=("rgba8", "ray_count"))
            frame = p.frame_offset + k * p.frame_stride
            view = np.ctypeslib.as_array(total_frames_words, shape=((frame + 1) * H, W))[frame * H:]
            view[rows] = out["rgba8"]
            st.total_rays += out["total_rays"]
        st.launches = 1
        return st

    def render_raw(self, cams, params, outputs):
        assert not self.queue, "a synchronous call with calls in flight"
        self.calls["render_raw"] += 1
        return self._render(list(cams), params, outputs)

    def render_async(self, cams, params, outputs):
        assert len(self.queue) < self.abi.RTX_MAX_IN_FLIGHT, "too many calls in flight"
        # the real call consumes params / outputs before it returns: the caller may change them afterwards
        self.queue.append((list(cams), self.abi.Params.from_buffer_copy(params), self.abi.Outputs.from_buffer_copy(outputs)))
        self.calls["render_async"] += 1
        self.calls["max_in_flight"] = max(self.calls["max_in_flight"], len(self.queue))

    def wait(self):
        return self._render(*self.queue.pop(0))


def _worker_sharded(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("ray-tracer-from-scratch_b200")
    R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
    SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
    from oracle import binding as ob
    S, abi, oracle = pkg.scene, pkg.abi, ob.load_port()
    ok = True
    # ---- one frame, cyclic row bands into ONE shared host frame (config C4's end-to-end shape), both frame modes ----
    scene = S.synthetic_scene(150, 6, seed=5)
    r = StandInRenderer(oracle, scene, abi, R)
    sh = SH.ShardedRenderer(r, rank, world, band_rows=3, n_chunks=3)
    pod = S.default_camera(40, 40.0 / 26).pod()                      # 26 rows: ragged over 3-row bands x 2 ranks
    for mode in (abi.RTX_FRAME_STORE, abi.RTX_FRAME_COPY):
        frame, st, launches = sh.render_frame(pod, max_depth=5, to_host=True, frame_mode=mode)
        dist.barrier()
        if rank == 0:
            exp = oracle.render(scene, pod, 5, want=("rgba8",))["rgba8"]
            ok = ok and frame.shape == exp.shape and np.array_equal(frame, exp)
        else:
            ok = ok and frame is None
        dist.barrier()
    # ---- a camera path, frames sharded, chunked asynchronous pipeline into the shared host frame set (config C5's shape) ----
    r.scene = S.default_scene()
    cams = [c.pod() for c in S.flythrough_cameras(256, 24, 16.0 / 9.0)[::37]]      # 7 frames: ragged over 2 ranks and 3 chunks
    frames, st, launches = sh.render_frames(cams, max_depth=10, to_host=True)
    dist.barrier()
    if rank == 0:
        for f, c in enumerate(cams):
            ok = ok and np.array_equal(frames[f], oracle.render(r.scene, c, 10, want=("rgba8",))["rgba8"])
    ok = ok and r.calls["render_async"] == len(SH.chunks_of(SH.frame_owner(len(cams), world)[rank], 3)) and r.calls["max_in_flight"] == min(abi.RTX_MAX_IN_FLIGHT, r.calls["render_async"])
    total = torch.tensor([st.total_rays if st else 0], dtype=torch.int64)
    dist.all_reduce(total)
    if rank == 0:
        ok = ok and int(total.item()) == sum(oracle.render(r.scene, c, 10, want=("ray_count",))["total_rays"] for c in cams)
    dist.barrier()
    sh.close()                                                       # importers first, barrier, then the owner (no hang, no leak)
    ok = ok and not r.maps and not [f for f in os.listdir("/dev/shm") if f.startswith("rtx_b200_%d_" % os.getpid())]
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(result_path, "w") as fh:
            fh.write("ok" if int(flag.item()) == 1 else "mismatch")
    dist.destroy_process_group()


def test_two_rank_shared_host_frame_and_async_pipeline(tmp_path):
    """ShardedRenderer's end-to-end plumbing over gloo with a stand-in renderer: every rank maps the same POSIX shared-memory
    frame, writes its own cyclic bands / its own frames of a camera path (chunked, several calls in flight), rank 0 reads the
    assembled result after the barrier."""
    result = tmp_path / "result.txt"
    mp.spawn(_worker_sharded, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"


def test_chunks_of():
    SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
    assert SH.chunks_of(list(range(32)), 4) == [list(range(k, k + 8)) for k in range(0, 32, 8)]
    assert SH.chunks_of([1, 3, 5], 4) == [[1], [3], [5]] and SH.chunks_of([], 4) == [] and SH.chunks_of(list(range(7)), 3) == [[0, 1, 2], [3, 4, 5], [6]]


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
