"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

Pixels are independent (main.cpp:129-138 has no cross-iteration state), so the only exchange is the final gather:

  * one big frame (config C4): image rows are grouped into cyclic bands of `band_rows` rows, band b -> rank b % G
    (contiguous stripes are measurably imbalanced: SURVEY.md §7 hard part 4). Each rank renders its rows with
    rtx_render(n_ranks=G, rank=r) into a packed device buffer; ONE all-gather delivers the band-major frame; rank 0
    scatters it to row-major with the rtx_unpermute_bands kernel.
  * a camera path (config C5): frame f -> rank f % G, each rank renders its frames in one batched launch, ONE
    all-gather, rank 0 reorders frames.

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


def frame_owner(n_frames, world):
    """Frame f is rendered by rank f % world; returns the frames of each rank."""
    return [list(range(r, n_frames, world)) for r in range(world)]


class ShardedRenderer:
    """Row-band / frame sharding around one Renderer per rank."""

    def __init__(self, renderer, rank, world, band_rows=4, group=None):
        self.r = renderer
        self.rank, self.world, self.band_rows, self.group = rank, world, band_rows, group
        self.device = torch.device("cuda", renderer.device)

    def render_frame(self, cam_pod, max_depth=10, want_ids=False, **param_overrides):
        """Renders one frame across all ranks. Returns (frame, stats): frame is an int32 CUDA tensor [H][W] of
        RGBA8888 words on rank 0 (None elsewhere); with want_ids also the object-id plane."""
        H, W = cam_pod.height, cam_pod.width
        rpr = rows_per_rank(H, self.band_rows, self.world)
        p = default_params(max_depth=max_depth, band_rows=self.band_rows, n_ranks=self.world, rank=self.rank,
                           **param_overrides)
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
