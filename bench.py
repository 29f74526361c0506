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
import gc
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
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "store", "copy"],
                    help="how pixels reach host memory in the e2e leg: store = the trace kernel writes the pinned mapped host frame itself "
                         "(zero copy), copy = device staging + copy-engine transfers; auto = store for the 10k-object scene, copy for the small one")
    ap.add_argument("--chunks", type=int, default=4, help="c5: chunks per rank of the render/copy pipeline")
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

    def reset(self):
        """Forget what was sampled so far (the warm-up steps): the timed region starts now."""
        self.samples, self.reasons = [], set()

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


def host_threads():
    """Hardware threads this process may run on. Passed to the harness EXPLICITLY (num_threads clause):
    torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, which must not throttle the CPU arm."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


CPU_CHUNK = 64          # pixels per OpenMP work item of the harness (oracle/ref_harness.cpp kChunk)


class CpuSample:
    """A bounded sample of the workload for the CPU arm: a cyclic subset of 4-row bands of one or two frames,
    sized from a short calibration so that one pass costs about `seconds` of wall time on all host threads.
    The harness hands out (row, 64-pixel chunk) work items dynamically; the sample always holds at least
    4 items per thread, so every core is busy whatever the core count."""

    def __init__(self, oracle, objs, pods, depth, seconds):
        import numpy as np
        self.oracle, self.objs, self.depth = oracle, objs, depth
        self.threads = host_threads()
        self.pods = pods[:: max(1, len(pods) // 2)][:2]
        H, self.W = self.pods[0].height, self.pods[0].width
        chunks_per_row = (self.W + CPU_CHUNK - 1) // CPU_CHUNK
        min_rows = -(-4 * self.threads // (chunks_per_row * len(self.pods)))          # >= 4 work items per thread
        min_rows = min(H, (min_rows + 3) // 4 * 4)
        probe = np.array(sample_rows(H, max(1, H // 4 // 2)), dtype=np.int32)[:max(8, min_rows)]
        t = sum(self._render(p, probe, ("radiance",))["seconds"] for p in self.pods)
        sec_per_row = max(t / (len(probe) * len(self.pods)), 1e-9)
        rows_budget = max(4, min_rows, min(H, int(seconds / (sec_per_row * len(self.pods))) // 4 * 4))
        self.every = max(1, (H // 4) // max(1, rows_budget // 4))
        self.rows = np.array(sample_rows(H, self.every), dtype=np.int32)
        self.height = H
        self.items = len(self.rows) * chunks_per_row * len(self.pods)
        # exact ray count of the sample: an untimed chain walk over the same rows
        self.rays = sum(self._render(p, self.rows, ("ray_count",))["total_rays"] for p in self.pods)

    def _render(self, pod, rows, want):
        return self.oracle.render(self.objs, pod, self.depth, rows=rows, threads=self.threads, want=want)

    def time_once(self):
        """recursive_ray_tracing over the sample rows, timed inside the harness around the pixel loop only
        (std::chrono, as main.cpp:326-330)."""
        return sum(self._render(p, self.rows, ("radiance",))["seconds"] for p in self.pods)

    def describe(self):
        return "every %d-th 4-row band of %d frame(s): %d rows x %d px = %d rays, %d work items of %d px over %d threads" % (
            self.every, len(self.pods), len(self.rows), self.W, self.rays, self.items, CPU_CHUNK, self.threads)


def structural_config(spec, S, pods, scene, world, band_rows):
    """The keys of `config` that name the workload — identical in both arms (ours and --impl reference)."""
    abi = importlib.import_module("ray-tracer-from-scratch_b200").abi
    n_spheres = sum(1 for g in scene if g.kind == abi.RTX_SPHERE)
    return {"workload": spec["label"], "width": pods[0].width, "height": pods[0].height, "frames_per_step": len(pods),
            "depth": spec["depth"], "n_spheres": n_spheres, "n_walls": len(scene) - n_spheres,
            "band_rows": band_rows if world > 1 and spec["name"] != "c5" else None,
            # one text for both arms, so that the two `config` objects are equal key for key
            "l2": "GPU arm: 256 MiB written between steps (L2 flush), outside the per-step CUDA events; inputs of the CPU arm exceed no cache rule (host)",
            "arms": "ours: one process per GPU, rows in cyclic bands (frames for the camera path); reference: the unmodified CPU implementation, "
                    "OpenMP over (row, 64-pixel chunk) work items on all host threads of rank 0, bounded sample of the same frame(s) — see `details`"}


def run_reference_arm(args, spec, S, world):
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
    config = structural_config(spec, S, pods, scene, world, args.band_rows)
    details = {"parallelism": "OpenMP, %d host threads, (row, 64-pixel chunk) work items, rank 0 only" % cs.threads,
               "sample": sample, "ms_per_frame_extrapolated": ms / (len(cs.rows) * len(cs.pods)) * cs.height}
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "n/a (CPU, rank 0 only)", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": config, "details": details,
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


# ---- parity of what was just timed ---------------------------------------------------------------------------
def parity_check(spec, planes):
    """Per-row CRC32 of the RGBA8 words of the frame(s) the timed region produced against tests/golden/fullsize_*.json —
    statistics the UNMODIFIED reference produced on the CPU (tests/golden/make_fullsize.py): every row of c2 / c3, every
    16th 4-row band of c4 (272 rows), frames 48 / 128 / 224 of c5. `planes`: {label: array [frames][H][W]} (device and
    host results). Returns the `parity` block of the JSON line."""
    import zlib
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "fullsize_%s.json" % spec["name"])
    if not os.path.exists(path):
        return {"rows_checked": 0, "rows_bad": None, "source": None, "note": "no golden file for this workload"}
    g = json.load(open(path))
    groups = [(int(k), v) for k, v in g["frames"].items()] if "frames" in g else [(0, g)]
    checked = bad = 0
    bad_list = []
    for label, arr in planes.items():
        a = np.asarray(arr)
        if a.ndim == 2:
            a = a[None]
        if a.shape[1:] != (g["height"], g["width"]):
            return {"rows_checked": 0, "rows_bad": None, "source": os.path.basename(path), "note": "frame size differs from the golden file (scaled run)"}
        for frame, gf in groups:
            rows = np.asarray(gf["rows"])
            sub = np.ascontiguousarray(a[frame][rows]).view(np.uint32)
            for r, row, want in zip(rows, sub, gf["rgba8_crc"]):
                checked += 1
                if zlib.crc32(row.tobytes()) != want:
                    bad += 1
                    if len(bad_list) < 8:
                        bad_list.append([label, frame, int(r)])
    out = {"rows_checked": checked, "rows_bad": bad, "source": os.path.basename(path) + " (unmodified reference, tests/golden/make_fullsize.py)",
           "checked": sorted(planes)}
    if bad_list:
        out["first_bad"] = bad_list
    return out


# ---- our arm -----------------------------------------------------------------------------------------------------
class Arm:
    """One process (= one GPU) of our arm: the renderer, the sharding layer and the timing loop for any workload."""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        self.pkg = importlib.import_module("ray-tracer-from-scratch_b200")
        self.S, self.abi = self.pkg.scene, self.pkg.abi
        self.R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
        self.SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
        self.dev = torch.device("cuda", local_rank)
        self.r = self.R.Renderer(local_rank)
        self.sh = self.SH.ShardedRenderer(self.r, rank, world, band_rows=args.band_rows, fused=args.gather == "fused", n_chunks=args.chunks)
        self.flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.int32, device=self.dev)
        self.host_single = None       # N = 1: pinned + mapped host frame (ptr, numpy view, key)

    def close(self):
        if self.host_single is not None:
            self.r.host_free(self.host_single[0])
            self.host_single = None
        self.sh.close()
        self.r.close()

    def _single_host(self, n_frames, H, W):
        import ctypes
        import numpy as np
        key = (n_frames, H, W)
        if self.host_single is None or self.host_single[2] != key:
            if self.host_single is not None:
                self.r.host_free(self.host_single[0])
            ptr = self.r.host_alloc(n_frames * H * W * 4)
            view = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), shape=key)
            self.host_single = (ptr, view, key)
        return self.host_single[0], self.host_single[1]

    def measure(self, spec, steps, warmup, e2e_steps, sampler=None):
        """Times `steps` device-resident steps and `e2e_steps` end-to-end steps of one workload; returns a dict on rank 0
        (None elsewhere). Every rank must call this with the same arguments."""
        import ctypes
        torch, dist, abi, S, R, r, sh, args = self.torch, self.dist, self.abi, self.S, self.R, self.r, self.sh, self.args
        world, rank, dev = self.world, self.rank, self.dev
        scene = build_scene(S, spec["scene"])
        objs = S.flatten(scene)
        if spec["name"] == "c5":
            pods = [c.pod() for c in S.flythrough_cameras(spec["frames"], spec["width"], 16.0 / 9.0)]
        else:
            pods = [S.default_camera(spec["width"], 16.0 / 9.0).pod()]
        H, W, F = pods[0].height, pods[0].width, len(pods)
        r.set_scene(objs)
        single = world == 1 and spec["name"] != "c5"
        e2e_store = args.e2e_mode == "store" or (args.e2e_mode == "auto" and spec["scene"] == "synthetic")
        params = R.default_params(max_depth=spec["depth"])
        outs = {}
        if single:
            dev_frame = torch.empty((F, H, W), dtype=torch.int32, device=dev)
            hptr, hview = self._single_host(F, H, W)
            o_dev, o_host = abi.Outputs(), abi.Outputs()
            o_dev.memory, o_dev.rgba8 = abi.RTX_MEM_DEVICE, dev_frame.data_ptr()
            o_host.memory, o_host.rgba8 = (abi.RTX_MEM_HOST_MAPPED if e2e_store else abi.RTX_MEM_HOST), hptr
            outs = {False: o_dev, True: o_host}
        result = {}
        drain = []
        frame_mode = abi.RTX_FRAME_STORE if e2e_store else abi.RTX_FRAME_COPY

        def step(e2e):
            """One step. Returns (rays on this rank, kernel ms, launches)."""
            if e2e:
                r.set_scene(objs)                                  # host -> device: the scene (main.cpp:156-163 equivalent)
            if spec["name"] == "c5":
                frames, st, launches = sh.render_frames(pods, max_depth=spec["depth"], to_host=e2e)
                result["host" if e2e else "device"] = frames
            elif world > 1:
                frame, st, launches = sh.render_frame(pods[0], max_depth=spec["depth"], to_host=e2e,
                                                      frame_mode=frame_mode if e2e else abi.RTX_FRAME_STORE)
                result["host" if e2e else "device"] = frame
            else:
                st = r.render_raw(pods, params, outs[e2e])
                launches = st.launches
                result["host" if e2e else "device"] = hview if e2e else dev_frame
            if st:
                drain.append(st.drain_ms)
            return (st.total_rays if st else 0), (st.raytracing_ms if st else 0.0), launches

        def timed_region(e2e, n_steps, n_warm, smp=None):
            # Warm-up steps are the timed steps without the clock: L2 flush included. (Until round 2 the flush kernel was
            # first launched in the first TIMED step, where CUDA's lazy module loading made that one launch block the host
            # for ~10 ms on rank 0; at N > 1 every other rank spent those 10 ms waiting in the first step's barrier — 1-2 ms
            # on the mean of a 5-10 step run: the whole "multi-GPU step overhead" of round 1. Found with RTX_BENCH_TRACE=1
            # and the per-rank step lists the line carries.)
            if smp:
                smp.start()
            for _ in range(n_warm):
                self.flush.add_(1)
                step(e2e)
            gc.collect()
            gc.disable()                                           # no cyclic GC pass over ~100 000 scene objects mid-measurement
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if smp:
                smp.reset()
            if not e2e:
                drain.clear()                                      # drain_ms_per_step_rank0 is the mean over the timed steps
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
            rays = kernel_ms = launches = 0
            trace = os.environ.get("RTX_BENCH_TRACE") == "1"
            t_sync = time.time()
            for k in range(n_steps):
                h0 = time.time()
                self.flush.add_(1)                                 # L2 flush, outside the step's events
                ev[k][0].record()
                h1 = time.time()
                a, b, c = step(e2e)
                h2 = time.time()
                ev[k][1].record()
                if trace and k < 2:
                    print("[trace] rank %d e2e=%s step %d: starts +%.2f ms after the barrier, flush+record %.2f ms, step() %.2f ms (kernel %.2f)" % (
                        rank, e2e, k, (h0 - t_sync) * 1e3, (h1 - h0) * 1e3, (h2 - h1) * 1e3, b), file=sys.stderr, flush=True)
                rays += a
                kernel_ms += b
                launches += c
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            gc.enable()
            clocks = smp.stop() if smp else None
            per_step = [s.elapsed_time(e) for s, e in ev]
            ms = sum(per_step)
            tag = "_e2e" if e2e else ""
            result["step_ms_rank0" + tag] = per_step
            if world > 1:      # every rank's per-step times, for the record (the headline uses the slowest rank's total)
                gaps = [ev[k][1].elapsed_time(ev[k + 1][0]) for k in range(n_steps - 1)] + [0.0]    # untimed: the L2 flush between steps
                mine = torch.tensor(per_step + gaps, dtype=torch.float64, device=dev)
                everyone = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(everyone, mine)
                result["step_ms_per_rank" + tag] = [[round(x, 3) for x in t.tolist()[:n_steps]] for t in everyone]
                result["gap_ms_per_rank" + tag] = [[round(x, 3) for x in t.tolist()[n_steps:-1]] for t in everyone]
            t = torch.tensor([ms, float(rays), kernel_ms, float(launches)], dtype=torch.float64, device=dev)
            per_rank_kernel = [kernel_ms / n_steps]
            if world > 1:
                tmax = t.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                allk = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
                dist.all_gather(allk, t[2:3].clone())
                per_rank_kernel = [x.item() / n_steps for x in allk]
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                ms, kernel_max = tmax[0].item(), tmax[2].item()
                rays, launches = t[1].item(), t[3].item()
            else:
                kernel_max = kernel_ms
            return ms, rays, kernel_max, int(launches), clocks, per_rank_kernel

        ms, rays, kernel_ms, launches, clocks, per_rank_kernel = timed_region(False, steps, warmup, sampler)
        drain_timed = list(drain)
        ms_e, rays_e, _, launches_e, _, _ = timed_region(True, e2e_steps, max(1, min(warmup, 2)))
        verified = None
        if args.verify and world > 1:
            torch.cuda.synchronize()
            if rank == 0:
                got = result["device"].clone().reshape(F, H, W)
                ok = True
                for f0 in range(0, F, 16):
                    part = pods[f0:f0 + 16]
                    alone = torch.empty((len(part), H, W), dtype=torch.int32, device=dev)
                    o = abi.Outputs()
                    o.memory, o.rgba8 = abi.RTX_MEM_DEVICE, alone.data_ptr()
                    r.render_raw(part, R.default_params(max_depth=spec["depth"]), o)
                    ok = ok and bool(torch.equal(alone, got[f0:f0 + len(part)]))
                verified = ok
            dist.barrier()
        import numpy as np
        torch.cuda.synchronize()
        planes = None
        if rank == 0:     # the frames of the timed regions, before anything else is rendered into the same buffers
            planes = {"device frame (value)": result["device"].reshape(F, H, W).cpu().numpy().view(np.uint32),
                      "host frame (e2e)": np.asarray(result["host"]).reshape(F, H, W).view(np.uint32)}
        # The same frame in plain scan order (rtx_params.pixel_order = RTX_ORDER_SCAN), untimed: what the scheduling hint of the
        # timed steps (tiles in the order of the previous frame's ray counts) is worth on this box.
        scan_drain, scan_kernel = [], []
        if spec["name"] != "c5":
            for _ in range(2):
                if world > 1:
                    _, st, _ = sh.render_frame(pods[0], max_depth=spec["depth"], pixel_order=abi.RTX_ORDER_SCAN)
                else:
                    st = r.render_raw(pods, R.default_params(max_depth=spec["depth"], pixel_order=abi.RTX_ORDER_SCAN), outs[False])
                if st:
                    scan_drain.append(st.drain_ms)
                    scan_kernel.append(st.raytracing_ms)
        if rank != 0:
            return None
        parity = parity_check(spec, planes)
        if planes["device frame (value)"].shape == planes["host frame (e2e)"].shape:
            parity["device_equals_host_frame"] = bool(np.array_equal(planes["device frame (value)"], planes["host frame (e2e)"]))
        n_spheres = sum(1 for g in scene if g.kind == abi.RTX_SPHERE)
        return dict(spec=spec, scene=scene, pods=pods, H=H, W=W, F=F, n_spheres=n_spheres, n_walls=len(scene) - n_spheres,
                    ms=ms, rays=rays, kernel_ms=kernel_ms, launches=launches, clocks=clocks, per_rank_kernel=per_rank_kernel,
                    ms_e=ms_e, rays_e=rays_e, launches_e=launches_e, steps=steps, e2e_steps=e2e_steps, drain=drain_timed, parity=parity,
                    scan_drain=scan_drain, scan_kernel=scan_kernel,
                    step_ms=result.get("step_ms_rank0"), step_ms_e2e=result.get("step_ms_rank0_e2e"),
                    step_ms_per_rank=result.get("step_ms_per_rank"), step_ms_per_rank_e2e=result.get("step_ms_per_rank_e2e"),
                    gap_ms_per_rank=result.get("gap_ms_per_rank"),
                    verified=verified, scene_bytes=len(scene) * ctypes.sizeof(abi.ObjectPOD), frame_bytes=H * W * 4 * F,
                    camera_bytes=ctypes.sizeof(abi.CameraPOD) * F,
                    e2e_path=("the trace kernel stores pixels straight into the pinned, mapped host frame (zero copy, RTX_FRAME_STORE / RTX_MEM_HOST_MAPPED)"
                              if e2e_store and spec["name"] != "c5" else
                              "device staging + copy-engine read-back on the copy stream, overlapped with tracing (RTX_FRAME_COPY / RTX_MEM_HOST)") +
                             ("; every rank writes its own rows of ONE shared pinned host frame over its own PCIe link" if world > 1 else ""))


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
            run_reference_arm(args, spec, S, max(args.gpus, world))
        return

    import torch
    import torch.distributed as dist
    R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
    abi = pkg.abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    arm = Arm(args, rank, local_rank, world)
    r = arm.r
    flush = arm.flush

    sampler = ClockSampler(visible_index(local_rank)) if rank == 0 and os.environ.get("RTX_BENCH_SAMPLER", "1") != "0" else None
    e2e_steps = max(2, min(args.steps, 5))
    m = arm.measure(spec, args.steps, args.warmup, e2e_steps, sampler)
    also = {}
    if args.workload == "auto" and not args.no_also:
        # BASELINE.json configs[4] (the 256-frame 1080p orbit, frames sharded over the ranks) in the same run, at every N
        m5 = arm.measure(workload_spec("c5", world), 5, 3, 3)
        if rank == 0:
            also["c5"] = {
                "workload": m5["spec"]["label"], "frames_per_step": m5["F"], "rays_per_step": m5["rays"] / m5["steps"],
                "device": {"ms_per_step": m5["ms"] / m5["steps"], "frames_per_s": m5["F"] / (m5["ms"] / m5["steps"] * 1e-3),
                           "mrays_s": m5["rays"] / (m5["ms"] * 1e-3) / 1e6, "kernel_ms_per_rank": m5["per_rank_kernel"],
                           "bytes_into_rank0_over_nvlink": m5["frame_bytes"] * (world - 1) // world},
                "e2e": {"ms_per_step": m5["ms_e"] / m5["e2e_steps"], "frames_per_s": m5["F"] / (m5["ms_e"] / m5["e2e_steps"] * 1e-3),
                        "mrays_s": m5["rays_e"] / (m5["ms_e"] * 1e-3) / 1e6, "d2h_bytes_per_step": m5["frame_bytes"],
                        "h2d_bytes_per_step": world * (m5["scene_bytes"] + m5["camera_bytes"] // world), "path": m5["e2e_path"]},
                "parity": m5["parity"], "gpu_launches": m5["launches"]}

    rc = 0
    if rank == 0:
        H, W, pods, scene = m["H"], m["W"], m["pods"], m["scene"]
        n_spheres, n_walls = m["n_spheres"], m["n_walls"]
        ms, rays, kernel_ms, launches = m["ms"], m["rays"], m["kernel_ms"], m["launches"]
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
        config = structural_config(spec, S, pods, scene, world, args.band_rows)
        details = ({
            "parallelism": ("1 process per GPU; " + ("cyclic row bands" if spec["name"] != "c5" else "frames sharded over ranks") +
                            (("; pixels stored into rank 0's frame over NVLink peer memory by the trace kernel + 1 barrier" if spec["name"] != "c5"
                              else "; finished frame chunks bulk-copied into rank 0's frame set over NVLink, overlapped with rendering")
                             if args.gather == "fused" else "; NCCL all-gather to rank 0 + unpermute")) if world > 1 else "single GPU",
            "rays_per_step": rays_per_step, "ms_per_frame": ms / args.steps / len(pods),
            "mpixel_per_s": H * W * len(pods) / (ms / args.steps * 1e-3) / 1e6})
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 screen + f64 decisions/shading", "data": "synthetic",
            "config": config, "details": details,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note, "kernel": "rtx::trace_kernel", "kernel_ms_per_step": kernel_ms / args.steps,
                         "kernel_ms_per_rank": m["per_rank_kernel"],
                         "algorithmic_flop_per_step": flops_per_step, "peak_source": peak_src,
                         "executed_flop_per_test": 14, "frac_executed": achieved / peak * 14.0 / 20.0 if n_walls * 33 < n_spheres else None,
                         "drain_ms_per_step_rank0": sum(m["drain"]) / max(1, len(m["drain"])),
                         "pixel_order": "RTX_ORDER_AUTO: tiles of 256 pixels handed out in descending order of the previous frame's ray counts "
                                        "(scheduling only, identical pixels; every step renders the same camera, so the hint is exact here)",
                         "scan_order_rank0": {"drain_ms": m["scan_drain"], "kernel_ms": m["scan_kernel"],
                                              "what": "the same frame with rtx_params.pixel_order = RTX_ORDER_SCAN, after the timed regions"},
                         "peak_measured_ffma2_tflops": ffma2, "peak_measured_ffma_scalar_tflops": ffma1,
                         "frac_of_measured_ffma2": achieved / ffma2 if ffma2 else None,
                         "hbm_write_gbs": m["frame_bytes"] / world / (kernel_ms / args.steps * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak_gbs()[0]},
            "e2e": {"value": m["rays_e"] / (m["ms_e"] * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": m["ms_e"] / e2e_steps, "steps": e2e_steps,
                    "h2d_bytes_per_step": world * m["scene_bytes"] + m["camera_bytes"] * (world if spec["name"] != "c5" else 1),
                    "d2h_bytes_per_step": m["frame_bytes"], "path": m["e2e_path"]},
            "parity": m["parity"],
            "step_ms_rank0": m["step_ms"], "e2e_step_ms_rank0": m["step_ms_e2e"],
            "step_ms_per_rank": m["step_ms_per_rank"], "e2e_step_ms_per_rank": m["step_ms_per_rank_e2e"],
            "untimed_gap_ms_per_rank": m["gap_ms_per_rank"],
            "gpu_launches": launches,
            "clocks": m["clocks"],
        }
        if m["verified"] is not None:
            line["verified_against_single_gpu"] = m["verified"]
        if also:
            line["also"] = also
        bad = [p for p in ([m["parity"]] + [v["parity"] for v in also.values() if "parity" in v]) if p.get("rows_bad")]
        if bad:
            rc = 1
        objs = S.flatten(scene)
        r.set_scene(objs)
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
            objs2 = S.flatten(scene2)
            for key, out_desc in (("device", o_dev), ("e2e", o_host)):
                for _ in range(3):
                    st2 = r.render_raw([pod2], p2, out_desc)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                tot_ms, kern = 0.0, 0.0
                for _ in range(20):
                    flush.add_(1)
                    e0.record()
                    if key == "e2e":
                        r.set_scene(objs2)
                    st2 = r.render_raw([pod2], p2, out_desc)
                    e1.record()
                    e1.synchronize()
                    tot_ms += e0.elapsed_time(e1)
                    kern += st2.raytracing_ms
                res[key] = {"ms_per_frame": tot_ms / 20, "mrays_s": st2.total_rays / (tot_ms / 20 * 1e-3) / 1e6, "kernel_ms": kern / 20}
            # the same end-to-end work as a stream of frames: rtx_render_async keeps three frames in flight, so the read-back
            # of frame k (copy stream) overlaps the kernel of frame k+1 and the host's queueing of frame k+2
            depth_q = abi.RTX_MAX_IN_FLIGHT
            hosts = [host2] + [torch.empty((pod2.height, pod2.width), dtype=torch.int32, pin_memory=True) for _ in range(depth_q - 1)]
            o_hosts = []
            for h in hosts:
                oh = abi.Outputs()
                oh.memory, oh.rgba8 = abi.RTX_MEM_HOST, h.data_ptr()
                o_hosts.append(oh)
            host2b = hosts[1]
            n_stream = 64
            for rep in range(2):                                  # first pass warms up
                flush.add_(1)
                e0.record()
                stage = [0.0, 0.0, 0.0, 0.0]
                for k in range(n_stream):
                    r.set_scene(objs2)
                    r.render_async([pod2], p2, o_hosts[k % depth_q])
                    if k >= depth_q - 1:
                        st2 = r.wait()
                        stage = [a + b for a, b in zip(stage, (st2.h2d_ms, st2.raytracing_ms, st2.d2h_ms, st2.total_ms))]
                for _ in range(depth_q - 1):
                    st2 = r.wait()
                e1.record()
                e1.synchronize()
            ms_stream = e0.elapsed_time(e1) / n_stream
            stage = [x / (n_stream - depth_q + 1) for x in stage]
            # ... and the device-resident frame as a stream (no host wait between frames: launch latency overlaps the kernel)
            for rep in range(2):
                flush.add_(1)
                e0.record()
                for k in range(n_stream):
                    r.render_async([pod2], p2, o_dev)
                    if k >= depth_q - 1:
                        st2 = r.wait()
                for _ in range(depth_q - 1):
                    st2 = r.wait()
                e1.record()
                e1.synchronize()
            ms_dstream = e0.elapsed_time(e1) / n_stream
            res["device_stream"] = {"ms_per_frame": ms_dstream, "mrays_s": st2.total_rays / (ms_dstream * 1e-3) / 1e6, "frames": n_stream,
                                    "frames_in_flight": depth_q, "api": "rtx_render_async + rtx_wait",
                                 "per_frame_device_stages_ms": {"h2d": stage[0], "kernel": stage[1], "kernel_end_to_readback_end": stage[2], "first_to_last_event": stage[3]}}
            # the PCIe read-back of one frame alone (pinned memory): the floor of any host-facing 1080p frame on this box
            ts = []
            for _ in range(8):
                e0.record()
                host2.copy_(dev2, non_blocking=True)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            d2h_ms = sorted(ts)[len(ts) // 2]
            res["d2h_copy_alone"] = {"ms": d2h_ms, "gbs": dev2.numel() * 4 / (d2h_ms * 1e-3) / 1e9}
            res["e2e_stream"] = {"ms_per_frame": ms_stream, "mrays_s": st2.total_rays / (ms_stream * 1e-3) / 1e6, "frames": n_stream,
                                 "frames_in_flight": depth_q, "api": "rtx_render_async + rtx_wait",
                                 "per_frame_device_stages_ms": {"h2d": stage[0], "kernel": stage[1], "kernel_end_to_readback_end": stage[2], "first_to_last_event": stage[3]}}
            import numpy as np
            par2 = parity_check(spec2, {"device frame": dev2.cpu().numpy().view(np.uint32), "host frame (e2e)": host2.numpy().view(np.uint32),
                                        "host frame (e2e_stream)": host2b.numpy().view(np.uint32)})
            line.setdefault("also", {})["c2"] = {"workload": spec2["label"], "rays_per_frame": st2.total_rays, **res, "parity": par2}
            if par2.get("rows_bad"):
                rc = 1
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
            par3 = parity_check(spec3, {"device frame": dev3.cpu().numpy().view(np.uint32)})
            line["also"]["c3"] = {"workload": spec3["label"], "rays_per_frame": st3.total_rays,
                                  "device": {"kernel_ms": k3, "mrays_s": st3.total_rays / (k3 * 1e-3) / 1e6,
                                             "tflops_algorithmic": fl3 / (k3 * 1e-3) / 1e12, "frac_of_fp32_peak": fl3 / (k3 * 1e-3) / 1e12 / peak},
                                  "parity": par3}
            if par3.get("rows_bad"):
                rc = 1
            # EXTENSION (rtx_params.accel = RTX_ACCEL_GRID): the same 4K frame with the uniform grid in front of
            # the same exact tests. Never the roofline line: it removes work instead of doing it faster.
            p3g = R.default_params(max_depth=spec3["depth"], accel=abi.RTX_ACCEL_GRID)
            ks = []
            for _ in range(8):
                flush.add_(1)
                torch.cuda.synchronize()
                st3g = r.render_raw([pod3], p3g, o3)
                ks.append(st3g.raytracing_ms)
            k3g = sorted(ks[3:])[len(ks[3:]) // 2]
            par3g = parity_check(spec3, {"device frame": dev3.cpu().numpy().view(np.uint32)})
            line["also"]["c3_accel"] = {"workload": spec3["label"].replace("brute force", "uniform grid (extension, default off)"),
                                        "rays_per_frame": st3g.total_rays, "rays_equal_brute_force": st3g.total_rays == st3.total_rays,
                                        "device": {"kernel_ms": k3g, "mrays_s": st3g.total_rays / (k3g * 1e-3) / 1e6, "speedup_over_brute_force": k3 / k3g},
                                        "parity": par3g}
            if par3g.get("rows_bad"):
                rc = 1
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, spec, S, scene, pods)
            except Exception as e:  # the oracle is optional equipment; say why it is missing
                line["cpu_baseline"] = {"unavailable": repr(e)}
        print(json.dumps(line), flush=True)
        if rc:
            print("bench.py: PARITY FAILURE — rows of the timed frame differ from the reference's golden CRCs", file=sys.stderr, flush=True)
    arm.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
