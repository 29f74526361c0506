"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

Pixels are independent (main.cpp:129-138 has no cross-iteration state), so the only exchange is the final gather:

  * one big frame (config C4): image rows are grouped into cyclic bands of `band_rows` rows, band b -> rank b % G
    (contiguous stripes are measurably imbalanced: SURVEY.md §7 hard part 4). Each rank renders its rows with
    rtx_render(n_ranks=G, rank=r) into a packed device buffer; ONE all-gather delivers the band-major frame; rank 0
    scatters it to row-major with the rtx_unpermute_bands kernel.
  * a camera path (config C5): frame f -> rank f % G; each rank renders its frames in chunks and bulk-copies every
    finished chunk into rank 0's frame set over NVLink while the next chunk renders (or, in gather mode, one batched
    launch + ONE all-gather + a reindexing on rank 0).

Two ways to bring the pixels to rank 0:

  * fused (default): rank 0 allocates the frame with rtx_buffer_alloc and exports a CUDA-IPC handle; every other rank
    maps it (rtx_buffer_import, peer access over NVLink) and passes it as rtx_outputs.frame_rgba8. The trace kernel
    then stores each finished pixel directly at its global position in rank 0's memory — the gather is part of the
    kernel, there is no all-gather and no unpermute pass, only one barrier.
  * gather: packed local buffers + all_gather_into_tensor + rtx_unpermute_bands (kept for comparison and for the
    optional object-id plane).

No data-path collective happens during tracing. Everything here is host logic + collectives; pixels are computed
only by the CUDA kernels behind the C ABI.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import abi
from .renderer import default_params, local_rows


def rows_per_rank(height, band_rows, world):
    """Equal all-gather block size: the largest local row count over ranks (ragged frames are padded)."""
    return max(local_rows(height, band_rows, world, r) for r in range(world))


def all_gather_blocks(local, world, group=None):
    """local: [rows_per_rank, ...] tensor (same shape on every rank) -> [world, rows_per_rank, ...] on every rank."""
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if world == 1:
        out[0].copy_(local)
    else:
        dist.all_gather_into_tensor(out.view(-1), local.reshape(-1), group=group)
    return out


def allreduce_tonemap_sums(sums, world, group=None):
    """Step 2 of the row-sharded tone map (extension, rtx_tonemap_sums / rtx_tonemap_apply): the per-frame log-luminance
    statistic is a 32.32 fixed-point INTEGER sum, so the all-reduce (SUM, int64; NCCL over NVLink on the GPUs) gives the
    same value whatever the number of ranks and the reduction order — the assembled frame equals the single-GPU one bit
    for bit. This is the one exchange the path has besides the final gather."""
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def tonemap_sharded(renderer, rad, pixels_per_frame_global, params, world, group=None):
    """rad: this rank's radiance rows, a contiguous CUDA tensor [n_frames][rows][W][3] (float32 or float64).
    Returns (int32 CUDA tensor [n_frames][rows][W] of RGBA8888 words for this rank's rows, the global int64 sums)."""
    n_frames = rad.shape[0]
    ppf = rad[0].numel() // 3
    sums = torch.zeros(n_frames, dtype=torch.int64, device=rad.device)
    out = torch.empty(tuple(rad.shape[:-1]), dtype=torch.int32, device=rad.device)
    torch.cuda.current_stream(rad.device).synchronize()          # the zero fill is done before the kernel adds to it
    renderer.tonemap_sums_device(rad.data_ptr(), rad.dtype == torch.float32, ppf, n_frames, sums.data_ptr())   # returns when complete
    allreduce_tonemap_sums(sums, world, group)
    torch.cuda.current_stream(rad.device).synchronize()          # ... and the all-reduce before the second pass reads it
    renderer.tonemap_apply_device(rad.data_ptr(), rad.dtype == torch.float32, ppf, n_frames, sums.data_ptr(), pixels_per_frame_global,
                                  params, out.data_ptr())
    return out, sums


def frame_owner(n_frames, world):
    """Frame f is rendered by rank f % world; returns the frames of each rank."""
    return [list(range(r, n_frames, world)) for r in range(world)]


class _DevicePtr:
    """Wraps a raw device pointer for torch.as_tensor through the CUDA array interface."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


class ShardedRenderer:
    """Row-band / frame sharding around one Renderer per rank."""

    def __init__(self, renderer, rank, world, band_rows=4, group=None, fused=True):
        self.r = renderer
        self.rank, self.world, self.band_rows, self.group = rank, world, band_rows, group
        self.device = torch.device("cuda", renderer.device)
        self.fused = fused
        self._frame = None      # (key, pointer valid in THIS process, tensor view on rank 0)
        self._token = None
        self._side = None       # side stream + double buffers of the pipelined camera-path gather
        self._bufs = {}

    # -- the shared frame on rank 0 -------------------------------------------------------------------
    def _shared_frame(self, n_frames, H, W):
        key = (n_frames, H, W)
        if self._frame is not None and self._frame[0] == key:
            return self._frame[1], self._frame[2]
        self.close()
        handle = [None]
        ptr = None
        if self.rank == 0:
            ptr = self.r.buffer_alloc(n_frames * H * W * 4)
            handle[0] = self.r.buffer_export(ptr)
        if self.world > 1:
            dist.broadcast_object_list(handle, src=0, group=self.group)
            if self.rank != 0:
                ptr = self.r.buffer_import(handle[0])
        view = torch.as_tensor(_DevicePtr(ptr, (n_frames, H, W)), device=self.device) if self.rank == 0 else None
        self._frame = (key, ptr, view)
        return ptr, view

    def close(self):
        if self._frame is not None:
            _, ptr, _ = self._frame
            if self.world > 1:
                dist.barrier(group=self.group)
            if self.rank == 0:
                self.r.buffer_free(ptr)
            else:
                self.r.buffer_release(ptr)
            self._frame = None

    def _fused_render(self, cam_pods, total_frames, H, W, params):
        ptr, view = self._shared_frame(total_frames, H, W)
        o = abi.Outputs()
        o.memory = abi.RTX_MEM_DEVICE
        o.frame_rgba8 = ptr
        st = self.r.render_raw(cam_pods, params, o) if cam_pods else None
        if self.world > 1:
            # Stream-ordered barrier: a 1-element all-reduce completes on rank 0 only after every rank's contribution,
            # which each rank enqueues behind its own trace kernel. Work queued after it on rank 0 sees the whole frame.
            if self._token is None:
                self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
            dist.all_reduce(self._token, group=self.group)
        return view, st, (st.launches if st else 0)

    def _pipelined_frames(self, cam_pods, mine, F, H, W, params, n_chunks=4):
        """Camera path, whole frames per rank: a frame is 8 MB, so instead of scattering 4-byte pixel stores over
        NVLink (fine for one frame spread over ranks, wasteful for gigabytes) each rank renders chunks of frames into
        local double buffers and copies every finished chunk into rank 0's IPC-mapped frame set with bulk peer copies
        on a side stream, overlapped with the rendering of the next chunk."""
        ptr, view = self._shared_frame(F, H, W)
        if self.rank == 0:
            # rank 0 owns the frame set: one launch, pixels stored straight at their place (local stores, no copies)
            params.frame_offset, params.frame_stride = 0, self.world
            o = abi.Outputs()
            o.memory, o.frame_rgba8 = abi.RTX_MEM_DEVICE, ptr
            st = self.r.render_raw([cam_pods[f] for f in mine], params, o) if mine else None
            if self.world > 1:
                if self._token is None:
                    self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
                dist.all_reduce(self._token, group=self.group)
            return view, st, (st.launches if st else 0)
        dest = torch.as_tensor(_DevicePtr(ptr, (F, H, W)), device=self.device)
        chunk = max(1, (len(mine) + n_chunks - 1) // n_chunks)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._bufs = {}
        key = (chunk, H, W)
        if key not in self._bufs:
            self._bufs = {key: ([torch.empty((chunk, H, W), dtype=torch.int32, device=self.device) for _ in range(2)],
                                [torch.cuda.Event(), torch.cuda.Event()])}
        bufs, events = self._bufs[key]
        o = abi.Outputs()
        o.memory = abi.RTX_MEM_DEVICE
        total = None
        launches = 0
        for ci, start in enumerate(range(0, len(mine), chunk)):
            frames = mine[start:start + chunk]
            buf = bufs[ci % 2]
            if ci >= 2:
                events[ci % 2].synchronize()           # the copies that read this buffer two chunks ago are done
            o.rgba8 = buf.data_ptr()
            st = self.r.render_raw([cam_pods[f] for f in frames], params, o)      # returns when the chunk is rendered
            launches += st.launches
            if total is None:
                total = st
            else:
                total.total_rays += st.total_rays
                total.raytracing_ms += st.raytracing_ms
                total.sphere_tests += st.sphere_tests
                total.wall_tests += st.wall_tests
            with torch.cuda.stream(self._side):
                for k, f in enumerate(frames):
                    dest[f].copy_(buf[k], non_blocking=True)
                events[ci % 2].record(self._side)
        torch.cuda.current_stream().wait_stream(self._side)
        if self.world > 1:
            if self._token is None:
                self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
            dist.all_reduce(self._token, group=self.group)     # stream-ordered barrier: all copies have landed
        return view, total, launches

    def render_frame(self, cam_pod, max_depth=10, want_ids=False, **param_overrides):
        """Renders one frame across all ranks. Returns (frame, stats): frame is an int32 CUDA tensor [H][W] of
        RGBA8888 words on rank 0 (None elsewhere); with want_ids also the object-id plane."""
        H, W = cam_pod.height, cam_pod.width
        rpr = rows_per_rank(H, self.band_rows, self.world)
        p = default_params(max_depth=max_depth, band_rows=self.band_rows, n_ranks=self.world, rank=self.rank,
                           **param_overrides)
        if self.fused and not want_ids:
            view, st, launches = self._fused_render([cam_pod], 1, H, W, p)
            return (view[0] if view is not None else None), st, launches
        local = torch.empty((rpr, W), dtype=torch.int32, device=self.device)
        ids = torch.empty((rpr, W), dtype=torch.int32, device=self.device) if want_ids else None
        o = abi.Outputs()
        o.memory = abi.RTX_MEM_DEVICE
        o.rgba8 = local.data_ptr()
        if want_ids:
            o.object_id = ids.data_ptr()
        st = self.r.render_raw([cam_pod], p, o)
        launches = st.launches
        gathered = all_gather_blocks(local, self.world, self.group)
        gathered_ids = all_gather_blocks(ids, self.world, self.group) if want_ids else None
        frame = frame_ids = None
        if self.rank == 0:
            frame = torch.empty((H, W), dtype=torch.int32, device=self.device)
            self.r.unpermute_bands(gathered.data_ptr(), frame.data_ptr(), H, W, 4, self.band_rows, self.world, rpr)
            launches += 1
            if want_ids:
                frame_ids = torch.empty((H, W), dtype=torch.int32, device=self.device)
                self.r.unpermute_bands(gathered_ids.data_ptr(), frame_ids.data_ptr(), H, W, 4, self.band_rows, self.world, rpr)
                launches += 1
        return (frame, frame_ids) if want_ids else frame, st, launches

    def render_frames(self, cam_pods, max_depth=10, **param_overrides):
        """Camera path: frame f on rank f % world, one batched launch per rank, one all-gather.
        Returns (frames, stats): int32 CUDA tensor [F][H][W] in frame order on rank 0 (None elsewhere)."""
        F = len(cam_pods)
        H, W = cam_pods[0].height, cam_pods[0].width
        mine = frame_owner(F, self.world)[self.rank]
        if self.fused:
            return self._pipelined_frames(cam_pods, mine, F, H, W, default_params(max_depth=max_depth, **param_overrides))
        per_rank = (F + self.world - 1) // self.world
        local = torch.zeros((per_rank, H, W), dtype=torch.int32, device=self.device)
        st = None
        launches = 0
        if mine:
            o = abi.Outputs()
            o.memory = abi.RTX_MEM_DEVICE
            o.rgba8 = local.data_ptr()
            st = self.r.render_raw([cam_pods[f] for f in mine], default_params(max_depth=max_depth, **param_overrides), o)
            launches = st.launches
        gathered = all_gather_blocks(local, self.world, self.group)        # [world][per_rank][H][W]
        frames = None
        if self.rank == 0:
            # frame f = r + world*k sits at gathered[r][k]: a pure reindexing
            frames = gathered.permute(1, 0, 2, 3).reshape(per_rank * self.world, H, W)[:F]
        return frames, st, launches
