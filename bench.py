#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the ray-tracing hot path on N B200s, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload auto|c2|c3|c4|c5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" renders one batch of the workload through the C ABI (include/rtx_b200.h):

    c2  1920x1080, default scene (main.cpp:160-163), depth 8                      BASELINE.json configs[1]
    c3  3840x2160, synthetic 10 000 spheres + 64 walls, depth 10                  configs[2]
    c4  7680x4320, same scene, cyclic 4-row bands over the N ranks                configs[3]  (auto, every N)
    c5  256 x 1080p camera orbit of the default scene, frames sharded over ranks  configs[4]

The default workload is c4 at every N: BASELINE.json quotes its metric "at 1080p/8K, 1/2/4/8 B200", the 8K frame fits one
GPU (0.2 s per frame), and one workload for the whole N = 1, 2, 4, 8 series makes it a literal strong-scaling run. The
default N = 1 run additionally measures c2 (1080p) and c3 (4K) and reports them under `also`.

Metric (BASELINE.json): Mrays/s = rays traced (primary + reflections, equal to the reference's count) / time.
`value`  : device-resident — scene and cameras already in HBM, outputs stay in HBM (CUDA events, max over ranks).
`e2e`    : the same through the host-facing call: scene + camera upload from host memory and the frame's RGBA8
           read back into pinned host memory inside the timed region.
`roofline`: the trace kernel against the FP32 FMA peak: algorithmic FLOPs = rays x (N_spheres*20 + N_walls*33)
            over the kernel's own CUDA-event time.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (oracle/_ref, built from /root/reference by
           oracle/build_ref.sh) row-parallel over all host threads on a bounded sample of the same frame.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_SPHERE, FLOP_WALL = 20, 33
L2_FLUSH_BYTES = 256 << 20               # > 126 MB L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--band-rows", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary configs[1] (c2) measurement in the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "allgather"],
                    help="N > 1: peer-memory stores from the trace kernel (fused) or all-gather + unpermute")
    ap.add_argument("--verify", action="store_true",
                    help="N > 1: rank 0 re-renders the frame(s) alone and compares them bit for bit with the gathered result")
    ap.add_argument("--scale", type=float, default=1.0, help="developer knob: shrink the frame (not a valid bench)")
    return ap.parse_args()


def workload_spec(name, n_gpus, scale=1.0):
    if name == "auto":
        name = "c4"
    spec = {
        "c2": dict(width=1920, depth=8, scene="default", frames=1, label="c2: 1920x1080 default scene (1 sphere + 2 walls), depth 8"),
        "c3": dict(width=3840, depth=10, scene="synthetic", frames=1, label="c3: 3840x2160 synthetic 10000 spheres + 64 walls, depth 10, brute force"),
        "c4": dict(width=7680, depth=10, scene="synthetic", frames=1, label="c4: 7680x4320 synthetic 10000 spheres + 64 walls, depth 10, brute force (cyclic row bands over the ranks when N > 1)"),
        "c5": dict(width=1920, depth=10, scene="default", frames=256, label="c5: 256-frame 1080p orbit of the default scene, frames sharded over ranks"),
    }[name]
    spec = dict(spec, name=name)
    if scale != 1.0:
        spec["width"] = max(16, int(spec["width"] * scale) // 16 * 16)
        spec["label"] += " [scaled x%g: NOT a valid bench]" % scale
    return spec


def build_scene(S, kind):
    return S.default_scene() if kind == "default" else S.synthetic_scene()


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 20 ms during the timed region (NVML)."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in list(self.BAD.items()) + list(self.NOTE.items()):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def visible_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def fp32_peak_tflops():
    """Denominator of the FP32 roofline. MEASURED_PEAKS.json (driver-written) carries no FP32 figure, so the peak
    is 148 SM x 128 lanes x 2 FLOP x its sm_max_mhz; the live FFMA2 microbenchmark is reported next to it."""
    mhz = 1965.0
    src = "148 SM x 128 lanes x 2 x 1965 MHz (fallback clock; MEASURED_PEAKS.json absent)"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            mhz = float(json.load(open(p)).get("sm_max_mhz", mhz))
            src = "148 SM x 128 lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json (it has no FP32 entry)"
        except Exception:
            pass
    return 148 * 128 * 2 * mhz * 1e6 / 1e12, src


def hbm_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ---- reference arm / cpu baseline -------------------------------------------------------------------------
def sample_rows(height, every, band=4):
    """Every `every`-th band of `band` rows: a cyclic subset that sees the whole frame."""
    return [r for r in range(height) if (r // band) % every == 0]


def reference_oracle():
    """oracle/_ref (the unmodified reference build) when present, else the C port. Test infrastructure used
    here ONLY as the thing being timed on the CPU side — never on the GPU arm's compute path."""
    import subprocess
    from oracle import binding as ob
    if not os.path.exists(ob.REF_PATH) and os.path.exists("/root/reference/main.cpp"):
        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    if os.path.exists(ob.REF_PATH):
        return ob.load_reference(), "reference"
    if not os.path.exists(ob.PORT_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return ob.load_port(), "port"


class CpuSample:
    """A bounded sample of the workload for the CPU arm: a cyclic subset of 4-row bands of one or two frames,
    sized from a short calibration so that one pass costs about `seconds` of wall time on all host threads."""

    def __init__(self, oracle, objs, pods, depth, seconds):
        import numpy as np
        self.oracle, self.objs, self.depth = oracle, objs, depth
        self.threads = oracle.max_threads()
        self.pods = pods[:: max(1, len(pods) // 2)][:2]
        H, self.W = self.pods[0].height, self.pods[0].width
        probe = np.array(sample_rows(H, max(1, H // 4 // 2)), dtype=np.int32)[:8]
        t = sum(self._render(p, probe, ("radiance",))["seconds"] for p in self.pods)
        sec_per_row = max(t / (len(probe) * len(self.pods)), 1e-9)
        rows_budget = max(4, min(H, int(seconds / (sec_per_row * len(self.pods))) // 4 * 4))
        self.every = max(1, (H // 4) // max(1, rows_budget // 4))
        self.rows = np.array(sample_rows(H, self.every), dtype=np.int32)
        self.height = H
        # exact ray count of the sample: an untimed chain walk over the same rows
        self.rays = sum(self._render(p, self.rows, ("ray_count",))["total_rays"] for p in self.pods)

    def _render(self, pod, rows, want):
        return self.oracle.render(self.objs, pod, self.depth, rows=rows, threads=self.threads, want=want)

    def time_once(self):
        """recursive_ray_tracing over the sample rows, timed inside the harness around the row loop only
        (std::chrono, as main.cpp:326-330)."""
        return sum(self._render(p, self.rows, ("radiance",))["seconds"] for p in self.pods)

    def describe(self):
        return "every %d-th 4-row band of %d frame(s): %d rows x %d px = %d rays" % (
            self.every, len(self.pods), len(self.rows), self.W, self.rays)


def run_reference_arm(args, spec, S):
    """`--impl reference`: the reference's own CPU implementation, all host threads, one bounded sample per step."""
    oracle, kind = reference_oracle()
    scene = build_scene(S, spec["scene"])
    if spec["name"] == "c5":
        pods = [c.pod() for c in S.flythrough_cameras(spec["frames"], spec["width"], 16.0 / 9.0)]
    else:
        pods = [S.default_camera(spec["width"], 16.0 / 9.0).pod()]
    budget = max(0.5, min(8.0, 120.0 / max(1, args.steps + args.warmup)))
    cs = CpuSample(oracle, S.flatten(scene), pods, spec["depth"], budget)
    for _ in range(args.warmup):
        cs.time_once()
    total = sum(cs.time_once() for _ in range(args.steps))
    ms = total / args.steps * 1e3
    value = cs.rays / (ms * 1e-3) / 1e6
    sample = cs.describe() + " per step"
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "n/a (CPU, rank 0 only)", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": spec["label"], "sample": sample,
                   "ms_per_frame_extrapolated": ms / (len(cs.rows) * len(cs.pods)) * cs.height},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cs.threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, spec, S, scene, pods):
    """The reference's CPU path on a bounded sample of the same workload (rank 0, N = 1): one pass of about
    args.cpu_seconds on all host threads (the calibration probe tends to undershoot, hence the 1.5x)."""
    oracle, kind = reference_oracle()
    cs = CpuSample(oracle, S.flatten(scene), pods, spec["depth"], args.cpu_seconds * 1.5)
    secs = cs.time_once()
    return {"value": cs.rays / secs / 1e6, "unit": "Mrays/s", "cores": cs.threads, "kind": kind,
            "sample": cs.describe() + ", %.2f s" % secs,
            "ms_per_frame_extrapolated": secs / (len(cs.rows) * len(cs.pods)) * cs.height * 1e3}


# ---- our arm -----------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pkg = importlib.import_module("ray-tracer-from-scratch_b200")
    S = pkg.scene
    spec = workload_spec(args.workload, max(args.gpus, world), args.scale)

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, spec, S)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
    SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
    abi = pkg.abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    scene = build_scene(S, spec["scene"])
    objs = S.flatten(scene)
    n_spheres = sum(1 for g in scene if g.kind == abi.RTX_SPHERE)
    n_walls = len(scene) - n_spheres
    if spec["name"] == "c5":
        pods = [c.pod() for c in S.flythrough_cameras(spec["frames"], spec["width"], 16.0 / 9.0)]
    else:
        pods = [S.default_camera(spec["width"], 16.0 / 9.0).pod()]
    H, W = pods[0].height, pods[0].width

    r = R.Renderer(local_rank)
    r.set_stream(torch.cuda.current_stream().cuda_stream)     # torch events then bracket our kernels
    r.set_scene(objs)
    sh = SH.ShardedRenderer(r, rank, world, band_rows=args.band_rows, fused=args.gather == "fused")
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.int32, device=dev)

    import ctypes
    scene_bytes = len(scene) * ctypes.sizeof(abi.ObjectPOD)
    frame_bytes = H * W * 4 * len(pods)
    host_frame = torch.empty((len(pods), H, W), dtype=torch.int32, pin_memory=True) if rank == 0 else None

    drain = []

    def step(e2e):
        """One step. Returns (rays on this rank, kernel ms, launches)."""
        if e2e:
            r.set_scene(objs)                                  # host -> device: the scene (main.cpp:156-163 equivalent)
        if spec["name"] == "c5":
            frames, st, launches = sh.render_frames(pods, max_depth=spec["depth"])
            if e2e and rank == 0:
                host_frame.copy_(frames, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        elif world > 1:
            frame, st, launches = sh.render_frame(pods[0], max_depth=spec["depth"])
            if e2e and rank == 0:
                host_frame[0].copy_(frame, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        else:
            st = r.render_raw(pods, single_params, single_out[1 if e2e else 0])
            launches = st.launches
        if st:
            drain.append(st.drain_ms)
        return (st.total_rays if st else 0), (st.raytracing_ms if st else 0.0), launches

    dev_frame = torch.empty((len(pods), H, W), dtype=torch.int32, device=dev) if world == 1 and spec["name"] != "c5" else None

    # single-GPU call arguments are built once, outside the timed region (as a caller rendering frame after frame would)
    single_params = R.default_params(max_depth=spec["depth"])
    single_out = [abi.Outputs(), abi.Outputs()]
    if dev_frame is not None:
        single_out[0].memory, single_out[0].rgba8 = abi.RTX_MEM_DEVICE, dev_frame.data_ptr()
        single_out[1].memory, single_out[1].rgba8 = abi.RTX_MEM_HOST, host_frame.data_ptr()

    def timed_region(e2e, steps, warmup, sampler=None):
        for _ in range(warmup):
            step(e2e)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        rays = kernel_ms = launches = 0
        for k in range(steps):
            flush.add_(1)                                      # L2 flush, outside the step's events
            ev[k][0].record()
            a, b, c = step(e2e)
            ev[k][1].record()
            rays += a
            kernel_ms += b
            launches += c
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        ms = sum(s.elapsed_time(e) for s, e in ev)
        t = torch.tensor([ms, float(rays), kernel_ms, float(launches)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, kernel_max = tmax[0].item(), tmax[2].item()
            rays, launches = t[1].item(), t[3].item()
        else:
            kernel_max = kernel_ms
        return ms, rays, kernel_max, int(launches), clocks

    verified = None
    if args.verify and world > 1:
        # every rank takes part in the sharded render; rank 0 then renders everything alone for comparison
        if spec["name"] == "c5":
            got, _, _ = sh.render_frames(pods, max_depth=spec["depth"])
        else:
            got, _, _ = sh.render_frame(pods[0], max_depth=spec["depth"])
        torch.cuda.synchronize()
        if rank == 0:
            got = got.clone().reshape(len(pods), H, W)
            ok = True
            for f0 in range(0, len(pods), 16):
                part = pods[f0:f0 + 16]
                alone = torch.empty((len(part), H, W), dtype=torch.int32, device=dev)
                o = abi.Outputs()
                o.memory, o.rgba8 = abi.RTX_MEM_DEVICE, alone.data_ptr()
                r.render_raw(part, R.default_params(max_depth=spec["depth"]), o)
                ok = ok and bool(torch.equal(alone, got[f0:f0 + len(part)]))
            verified = ok
        dist.barrier()

    sampler = ClockSampler(visible_index(local_rank)) if rank == 0 else None
    ms, rays, kernel_ms, launches, clocks = timed_region(False, args.steps, args.warmup, sampler)
    ms_e, rays_e, _, _, _ = timed_region(True, max(2, min(args.steps, 5)), 1)
    e2e_steps = max(2, min(args.steps, 5))

    if rank == 0:
        value = rays / (ms * 1e-3) / 1e6
        rays_per_step = rays / args.steps
        # roofline of the trace kernel: slowest rank's kernel time, that rank's share of the algorithmic work
        flops_per_step = rays_per_step * (n_spheres * FLOP_SPHERE + n_walls * FLOP_WALL)
        peak, peak_src = fp32_peak_tflops()
        achieved = flops_per_step / world / (kernel_ms / args.steps * 1e-3) / 1e12
        ffma2, _ = r.ffma_peak(1)
        ffma1, _ = r.ffma_peak(0)
        traffic = traffic_note = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            t = json.load(open(tp)).get(spec["name"])
            if t:
                traffic, traffic_note = t["dram_bytes_per_launch"], t.get("note")
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 screen + f64 decisions/shading", "data": "synthetic",
            "config": {"workload": spec["label"], "width": W, "height": H, "frames_per_step": len(pods), "depth": spec["depth"],
                       "n_spheres": n_spheres, "n_walls": n_walls, "band_rows": args.band_rows if world > 1 else None,
                       "parallelism": ("1 process per GPU; " + ("cyclic row bands" if spec["name"] != "c5" else "frames sharded over ranks") +
                                       (("; pixels stored into rank 0's frame over NVLink peer memory by the trace kernel + 1 barrier" if spec["name"] != "c5"
                                         else "; finished frame chunks bulk-copied into rank 0's frame set over NVLink, overlapped with rendering")
                                        if args.gather == "fused" else "; NCCL all-gather to rank 0 + unpermute")) if world > 1 else "single GPU",
                       "rays_per_step": rays_per_step, "ms_per_frame": ms / args.steps / len(pods),
                       "mpixel_per_s": H * W * len(pods) / (ms / args.steps * 1e-3) / 1e6,
                       "l2": "256 MiB written between steps (L2 flush), outside the per-step CUDA events"},
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note, "kernel": "rtx::trace_kernel", "kernel_ms_per_step": kernel_ms / args.steps,
                         "algorithmic_flop_per_step": flops_per_step, "peak_source": peak_src,
                         "drain_ms_per_step_rank0": sum(drain) / max(1, len(drain)),
                         "peak_measured_ffma2_tflops": ffma2, "peak_measured_ffma_scalar_tflops": ffma1,
                         "frac_of_measured_ffma2": achieved / ffma2 if ffma2 else None,
                         "hbm_write_gbs": frame_bytes / world / (kernel_ms / args.steps * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak_gbs()[0]},
            "e2e": {"value": rays_e / (ms_e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e / e2e_steps, "steps": e2e_steps,
                    "h2d_bytes_per_step": scene_bytes + ctypes.sizeof(abi.CameraPOD) * len(pods), "d2h_bytes_per_step": frame_bytes},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if verified is not None:
            line["verified_against_single_gpu"] = verified
        if world == 1:
            # HBM side: the standalone quantise kernel (main.cpp:338-347) on a device-resident radiance frame of the
            # same size; algorithmic bytes = 12 (f32) or 24 (f64) read + 4 written per pixel; L2 flushed before each launch.
            hbm_peak, hbm_src = hbm_peak_gbs()
            npx = H * W * len(pods)
            qout = torch.empty(npx, dtype=torch.int32, device=dev)
            line["roofline_quantise"] = {}
            for name, dt, bpp in (("f32", torch.float32, 16), ("f64", torch.float64, 28)):
                rad = torch.rand(npx * 3, dtype=dt, device=dev) * 1.3
                ts = []
                for _ in range(7):
                    flush.add_(1)
                    torch.cuda.synchronize()
                    ts.append(r.quantise_device(rad.data_ptr(), dt == torch.float32, npx, qout.data_ptr()).surface_update_ms)
                ms_q = sorted(ts[2:])[len(ts[2:]) // 2]
                line["roofline_quantise"][name] = {"bound": "hbm", "achieved": npx * bpp / (ms_q * 1e-3) / 1e9, "peak": hbm_peak,
                                                   "unit": "GB/s", "frac": npx * bpp / (ms_q * 1e-3) / 1e9 / hbm_peak, "ms": ms_q,
                                                   "bytes_per_pixel": bpp, "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs)"}
                del rad
            # extension: the Reinhard tone-map operator (two launches: luminance statistic, then map + pack) on the same
            # frame; algorithmic bytes = the radiance read TWICE + 4 B written per pixel.
            line["roofline_tonemap"] = {}
            ptm = R.default_params(tonemap=abi.RTX_TONEMAP_REINHARD, quantise_mode=abi.RTX_QUANT_SATURATE)
            for name, dt, bpp in (("f32", torch.float32, 28), ("f64", torch.float64, 52)):
                rad = torch.rand(npx * 3, dtype=dt, device=dev) * 1.3
                ts = []
                for _ in range(7):
                    flush.add_(1)
                    torch.cuda.synchronize()
                    ts.append(r.tonemap_device(rad.data_ptr(), dt == torch.float32, npx, 1, ptm, qout.data_ptr()).surface_update_ms)
                ms_q = sorted(ts[2:])[len(ts[2:]) // 2]
                line["roofline_tonemap"][name] = {"bound": "hbm", "achieved": npx * bpp / (ms_q * 1e-3) / 1e9, "peak": hbm_peak,
                                                  "unit": "GB/s", "frac": npx * bpp / (ms_q * 1e-3) / 1e9 / hbm_peak, "ms": ms_q,
                                                  "bytes_per_pixel": bpp, "launches": 2,
                                                  "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs)"}
                del rad
        if world == 1 and args.workload == "auto" and not args.no_also:
            # BASELINE.json configs[1] (1080p default scene, depth 8) in the same run: device-resident and end to end
            spec2 = workload_spec("c2", 1)
            scene2 = build_scene(S, spec2["scene"])
            pod2 = S.default_camera(spec2["width"], 16.0 / 9.0).pod()
            r.set_scene(S.flatten(scene2))
            p2 = R.default_params(max_depth=spec2["depth"])
            dev2 = torch.empty((pod2.height, pod2.width), dtype=torch.int32, device=dev)
            host2 = torch.empty((pod2.height, pod2.width), dtype=torch.int32, pin_memory=True)
            o_dev, o_host = abi.Outputs(), abi.Outputs()
            o_dev.memory, o_dev.rgba8 = abi.RTX_MEM_DEVICE, dev2.data_ptr()
            o_host.memory, o_host.rgba8 = abi.RTX_MEM_HOST, host2.data_ptr()
            res = {}
            for key, out_desc in (("device", o_dev), ("e2e", o_host)):
                for _ in range(3):
                    st2 = r.render_raw([pod2], p2, out_desc)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                tot_ms, kern = 0.0, 0.0
                for _ in range(20):
                    flush.add_(1)
                    e0.record()
                    if key == "e2e":
                        r.set_scene(S.flatten(scene2))
                    st2 = r.render_raw([pod2], p2, out_desc)
                    e1.record()
                    e1.synchronize()
                    tot_ms += e0.elapsed_time(e1)
                    kern += st2.raytracing_ms
                res[key] = {"ms_per_frame": tot_ms / 20, "mrays_s": st2.total_rays / (tot_ms / 20 * 1e-3) / 1e6, "kernel_ms": kern / 20}
            line["also"] = {"c2": {"workload": spec2["label"], "rays_per_frame": st2.total_rays, **res}}
            if not args.no_cpu_baseline:
                try:
                    oracle2, kind2 = reference_oracle()
                    cs2 = CpuSample(oracle2, S.flatten(scene2), [pod2], spec2["depth"], 2.0)
                    secs2 = min(cs2.time_once(), cs2.time_once())
                    line["also"]["c2"]["cpu_baseline"] = {"value": cs2.rays / secs2 / 1e6, "unit": "Mrays/s", "cores": cs2.threads, "kind": kind2,
                                                          "sample": cs2.describe() + ", %.3f s" % secs2,
                                                          "ms_per_frame_extrapolated": secs2 / len(cs2.rows) * cs2.height * 1e3}
                except Exception as e:
                    line["also"]["c2"]["cpu_baseline"] = {"unavailable": repr(e)}
            r.set_scene(objs)
            # BASELINE.json configs[2] (4K, the same 10 064-object scene): device-resident kernel time and roofline
            spec3 = workload_spec("c3", 1)
            pod3 = S.default_camera(spec3["width"], 16.0 / 9.0).pod()
            dev3 = torch.empty((pod3.height, pod3.width), dtype=torch.int32, device=dev)
            o3 = abi.Outputs()
            o3.memory, o3.rgba8 = abi.RTX_MEM_DEVICE, dev3.data_ptr()
            p3 = R.default_params(max_depth=spec3["depth"])
            ks = []
            for _ in range(8):
                flush.add_(1)
                torch.cuda.synchronize()
                st3 = r.render_raw([pod3], p3, o3)
                ks.append(st3.raytracing_ms)
            k3 = sorted(ks[3:])[len(ks[3:]) // 2]
            fl3 = st3.total_rays * (n_spheres * FLOP_SPHERE + n_walls * FLOP_WALL)
            line["also"]["c3"] = {"workload": spec3["label"], "rays_per_frame": st3.total_rays,
                                  "device": {"kernel_ms": k3, "mrays_s": st3.total_rays / (k3 * 1e-3) / 1e6,
                                             "tflops_algorithmic": fl3 / (k3 * 1e-3) / 1e12, "frac_of_fp32_peak": fl3 / (k3 * 1e-3) / 1e12 / peak}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, spec, S, scene, pods)
            except Exception as e:  # the oracle is optional equipment; say why it is missing
                line["cpu_baseline"] = {"unavailable": repr(e)}
        print(json.dumps(line), flush=True)
    sh.close()
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
