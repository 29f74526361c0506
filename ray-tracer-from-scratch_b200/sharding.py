"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

Pixels are independent (main.cpp:129-138 has no cross-iteration state), so the only exchange is the final gather:

  * one big frame (config C4): image rows are grouped into cyclic bands of `band_rows` rows, band b -> rank b % G
    (contiguous stripes are measurably imbalanced: SURVEY.md §7 hard part 4).
  * a camera path (config C5): frame f -> rank f % G, rendered in chunks.

Where the pixels go (`to_host`):

  * rank 0's HBM (default; BASELINE.json: "gathered to rank 0"): rank 0 allocates the frame (set) with
    rtx_buffer_alloc and exports a CUDA-IPC handle; the other ranks map it (peer access over NVLink).
  * a pinned HOST frame shared by all ranks (`to_host=True`; the end-to-end path): a POSIX shared-memory object that
    every rank maps and pins (rtx_host_shared_open). Each rank delivers its own rows over its OWN PCIe link — rank 0's
    link no longer carries the whole frame.

How they get there (rtx_outputs.frame_mode):

  * RTX_FRAME_STORE — the trace kernel stores every finished pixel at its global position (peer memory or the device
    alias of the host frame): the gather is part of the kernel, only a barrier follows. Right for the 10k-object
    scene, where a pixel costs microseconds.
  * RTX_FRAME_COPY — each call renders into context staging and copy-engine transfers (2-D copies for bands, whole
    frames for a camera path) move it, overlapped with the next chunk's kernel through rtx_render_async. Right for
    the three-object scene, where pixels are produced at tens of gigabytes per second.
  * gather (`fused=False`): packed local buffers + ONE all_gather_into_tensor + rtx_unpermute_bands (kept as the
    comparison BASELINE.json names, and for the optional object-id plane).

No data-path collective happens during tracing. Everything here is host logic + collectives; pixels are computed
only by the CUDA kernels behind the C ABI.

Streams: the Renderer is bound to torch's current stream on construction, so that kernels, NCCL collectives and the
unpermute pass are all ordered on ONE stream (the collectives are ordered against torch's current stream only).
"""
import ctypes as C
import os
import uuid

import numpy as np
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


def chunks_of(frames, n_chunks):
    """A rank's frames in at most n_chunks consecutive pieces (the units of the render/copy pipeline)."""
    if not frames:
        return []
    size = max(1, (len(frames) + n_chunks - 1) // n_chunks)
    return [frames[k:k + size] for k in range(0, len(frames), size)]


class _DevicePtr:
    """Wraps a raw device pointer for torch.as_tensor through the CUDA array interface."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


def _add_stats(total, st):
    if total is None:
        return st
    total.total_rays += st.total_rays
    total.raytracing_ms += st.raytracing_ms
    total.sphere_tests += st.sphere_tests
    total.wall_tests += st.wall_tests
    total.d2h_ms += st.d2h_ms
    total.launches += st.launches
    return total


class ShardedRenderer:
    """Row-band / frame sharding around one Renderer per rank."""

    def __init__(self, renderer, rank, world, band_rows=4, group=None, fused=True, n_chunks=4):
        self.r = renderer
        self.rank, self.world, self.band_rows, self.group = rank, world, band_rows, group
        # The token tensor of the stream-ordered barrier lives where the collective backend wants it: on this rank's GPU
        # (NCCL). Without CUDA (the gloo tests of this host logic, with a stand-in renderer) it is a CPU tensor; pixels are
        # never computed here either way.
        self.on_gpu = torch.cuda.is_available()
        self.device = torch.device("cuda", renderer.device) if self.on_gpu else torch.device("cpu")
        self.fused = fused
        self.n_chunks = n_chunks
        self._frame = None      # device frame (set) on rank 0: (key, pointer valid in THIS process, tensor view on rank 0)
        self._host = None       # shared pinned host frame (set): (key, host pointer, device alias, numpy view)
        self._token = None
        # NCCL orders its collectives against torch's CURRENT stream only: bind the renderer's work to that stream, so
        # that the all-gather -> unpermute and kernel -> barrier orders hold without extra synchronisation.
        if self.on_gpu:
            renderer.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def _sync(self):
        if self.on_gpu:
            torch.cuda.current_stream(self.device).synchronize()

    # -- the shared frame on rank 0 -------------------------------------------------------------------
    def _shared_frame(self, n_frames, H, W):
        key = (n_frames, H, W)
        if self._frame is not None and self._frame[0] == key:
            return self._frame[1], self._frame[2]
        self._close_device_frame()
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

    def _shared_host_frame(self, n_frames, H, W):
        """A pinned host frame set every rank can write: POSIX shared memory, mapped + cudaHostRegister'ed in each process."""
        key = (n_frames, H, W)
        if self._host is not None and self._host[0] == key:
            return self._host[1:]
        self._close_host_frame()
        nbytes = n_frames * H * W * 4
        if self.world == 1:
            hptr = self.r.host_alloc(nbytes)
            name = None
        else:
            box = [None]
            if self.rank == 0:
                box[0] = "/rtx_b200_%d_%s" % (os.getpid(), uuid.uuid4().hex[:12])
                hptr = self.r.host_shared_open(box[0], nbytes, create=True)
            dist.broadcast_object_list(box, src=0, group=self.group)
            name = box[0]
            if self.rank != 0:
                hptr = self.r.host_shared_open(name, nbytes, create=False)
            dist.barrier(group=self.group)                 # everyone has it mapped ...
            if self.rank == 0:
                os.unlink("/dev/shm" + name)               # ... so the name can go; the mappings keep the memory alive
        dptr = self.r.host_device_pointer(hptr)
        view = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_uint32)), shape=(n_frames, H, W))
        self._host = (key, hptr, dptr, view, name)
        return hptr, dptr, view, name

    def _close_device_frame(self):
        if self._frame is not None:
            _, ptr, _ = self._frame
            self._frame = None
            self._sync()
            # importers unmap FIRST, then everybody meets, then the owner frees: freeing exported memory that a peer
            # still has open is undefined behaviour
            if self.rank != 0:
                self.r.buffer_release(ptr)
            if self.world > 1:
                dist.barrier(group=self.group)
            if self.rank == 0:
                self.r.buffer_free(ptr)

    def _close_host_frame(self):
        if self._host is not None:
            _, hptr, _, _, name = self._host
            self._host = None
            self._sync()
            if self.world > 1:
                dist.barrier(group=self.group)             # nobody still writes it
                self.r.host_shared_close(hptr)
            else:
                self.r.host_free(hptr)

    def close(self):
        self._close_device_frame()
        self._close_host_frame()

    def _barrier(self):
        """Stream-ordered barrier: a 1-element all-reduce completes on rank 0 only after every rank's contribution,
        which each rank enqueues behind its own work on the same stream. Work queued after it sees the whole frame."""
        if self.world > 1:
            if self._token is None:
                self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
            dist.all_reduce(self._token, group=self.group)

    def _destination(self, n_frames, H, W, to_host):
        """(pointer to hand to rtx_outputs.frame_rgba8 in STORE mode, same in COPY mode, rank 0's view of the result)."""
        if to_host:
            hptr, dptr, view, _ = self._shared_host_frame(n_frames, H, W)
            return dptr, hptr, (view if self.rank == 0 else None)
        ptr, view = self._shared_frame(n_frames, H, W)
        return ptr, ptr, view

    # -- one frame, rows sharded ------------------------------------------------------------------------------------
    def render_frame(self, cam_pod, max_depth=10, want_ids=False, to_host=False, frame_mode=abi.RTX_FRAME_STORE, **param_overrides):
        """Renders one frame across all ranks. Returns (frame, stats, launches): on rank 0 the frame is an int32 CUDA
        tensor [H][W] of RGBA8888 words (to_host: a uint32 numpy view of the shared pinned host frame), None elsewhere;
        with want_ids also the object-id plane (all-gather path)."""
        H, W = cam_pod.height, cam_pod.width
        rpr = rows_per_rank(H, self.band_rows, self.world)
        p = default_params(max_depth=max_depth, band_rows=self.band_rows, n_ranks=self.world, rank=self.rank,
                           **param_overrides)
        if self.fused and not want_ids:
            store_ptr, copy_ptr, view = self._destination(1, H, W, to_host)
            o = abi.Outputs()
            o.memory = abi.RTX_MEM_DEVICE
            o.frame_mode = frame_mode
            o.frame_rgba8 = copy_ptr if frame_mode == abi.RTX_FRAME_COPY else store_ptr
            st = self.r.render_raw([cam_pod], p, o)          # COPY mode: returns when this rank's bands have landed
            self._barrier()
            if to_host and self.rank == 0:
                self._sync()                  # the barrier has completed: every rank's rows are in host memory
            return (view[0] if view is not None else None), st, st.launches
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
        if to_host and self.rank == 0 and not want_ids:
            frame = self._read_back(frame)          # the round-1 path: rank 0's PCIe link carries the whole frame
        return (frame, frame_ids) if want_ids else frame, st, launches

    def _read_back(self, device_frames):
        """All-gather comparison path: copies the assembled frame (set) on rank 0 into a pinned host buffer; numpy view."""
        key = tuple(device_frames.shape)
        if getattr(self, "_pinned", None) is None or self._pinned[0] != key:
            self._pinned = (key, torch.empty(key, dtype=torch.int32, pin_memory=True))
        self._pinned[1].copy_(device_frames, non_blocking=True)
        self._sync()
        return self._pinned[1].numpy().view(np.uint32)

    # -- a camera path, frames sharded ------------------------------------------------------------------------------
    def _pipelined_frames(self, cam_pods, mine, F, H, W, params, to_host):
        """Camera path, whole frames per rank: a frame is 8 MB, so instead of scattering 4-byte pixel stores (fine for
        one costly frame spread over ranks, wasteful for gigabytes of cheap pixels) each rank renders chunks of frames with
        rtx_render_async in RTX_FRAME_COPY mode: a chunk is traced into one of the context's two staging slots and
        copy-engine transfers on the context's copy stream place its frames in the frame set (rank 0's HBM over NVLink,
        or the shared pinned host frame over this rank's own PCIe link) while the next chunk is traced."""
        store_ptr, copy_ptr, view = self._destination(F, H, W, to_host)
        o = abi.Outputs()
        o.memory = abi.RTX_MEM_DEVICE
        params.frame_stride = self.world
        total = None
        if self.rank == 0 and not to_host:
            # rank 0 owns the device frame set: one launch, pixels stored straight at their place (local stores, no copies)
            params.frame_offset = 0
            o.frame_mode, o.frame_rgba8 = abi.RTX_FRAME_STORE, store_ptr
            if mine:
                total = self.r.render_raw([cam_pods[f] for f in mine], params, o)
        else:
            o.frame_mode, o.frame_rgba8 = abi.RTX_FRAME_COPY, copy_ptr
            in_flight = 0
            for frames in chunks_of(mine, self.n_chunks):
                if in_flight == abi.RTX_MAX_IN_FLIGHT:
                    total = _add_stats(total, self.r.wait())
                    in_flight -= 1
                params.frame_offset = frames[0]              # frame k of this call is global frame frames[0] + k * world
                self.r.render_async([cam_pods[f] for f in frames], params, o)
                in_flight += 1
            while in_flight:
                total = _add_stats(total, self.r.wait())     # returns when that chunk's frames have landed
                in_flight -= 1
        self._barrier()
        if to_host and self.rank == 0:
            self._sync()
        return view, total, (total.launches if total else 0)

    def render_frames(self, cam_pods, max_depth=10, to_host=False, **param_overrides):
        """Camera path: frame f on rank f % world. Returns (frames, stats, launches): on rank 0 an int32 CUDA tensor
        [F][H][W] in frame order (to_host: a uint32 numpy view of the shared pinned host frame set), None elsewhere."""
        F = len(cam_pods)
        H, W = cam_pods[0].height, cam_pods[0].width
        mine = frame_owner(F, self.world)[self.rank]
        if self.fused:
            return self._pipelined_frames(cam_pods, mine, F, H, W, default_params(max_depth=max_depth, **param_overrides), to_host)
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
            if to_host:
                frames = self._read_back(frames.contiguous())
        return frames, st, launches
