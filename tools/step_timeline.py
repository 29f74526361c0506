#!/usr/bin/env python
"""Developer tool (torchrun, N >= 2): host wall-clock timeline of one sharded c4 step on every rank — where does the
time between the end of the trace kernel and the end of the step go?
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_timeline.py [width]
"""
import importlib
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ray-tracer-from-scratch_b200")
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")
SH = importlib.import_module("ray-tracer-from-scratch_b200.sharding")
S, abi = pkg.scene, pkg.abi

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
width = int(sys.argv[1]) if len(sys.argv) > 1 else 7680
r = R.Renderer(lr)
sh = SH.ShardedRenderer(r, rank, world)
r.set_scene(S.synthetic_scene())
pod = S.default_camera(width, 16.0 / 9.0).pod()
H, W = pod.height, pod.width
flush = torch.empty(64 << 20, dtype=torch.int32, device="cuda")
for to_host in (False, True):
    store_ptr, copy_ptr, view = sh._destination(1, H, W, to_host)
    p = R.default_params(max_depth=10, band_rows=4, n_ranks=world, rank=rank)
    o = abi.Outputs()
    o.memory, o.frame_mode, o.frame_rgba8 = abi.RTX_MEM_DEVICE, abi.RTX_FRAME_STORE, store_ptr
    rows = []
    for it in range(6):
        flush.add_(1)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.time()
        e0.record()
        st = r.render_raw([pod], p, o)
        t1 = time.time()
        e1.record()
        sh._barrier()
        t2 = time.time()
        e2.record()
        torch.cuda.current_stream().synchronize()
        t3 = time.time()
        rows.append((t0, t1 - t0, t2 - t1, t3 - t2, st.raytracing_ms, e0.elapsed_time(e1), e1.elapsed_time(e2), st.drain_ms, st.total_ms))
    for k in range(world):
        dist.barrier()
        if k == rank:
            for t0, a, b, c, km, ev01, ev12, dr, tot in rows[2:]:
                print("to_host=%d rank %d start %.6f  render_raw %.3f ms (kernel %.3f, rtx total %.3f, drain %.3f)  enqueue barrier %.3f ms  wait %.3f ms | events: render %.3f barrier %.3f"
                      % (to_host, rank, t0 % 100, a * 1e3, km, tot, dr, b * 1e3, c * 1e3, ev01, ev12), flush=True)
# ---- free-running, exactly like bench.py's timed region (no barrier / synchronise between steps) -----------------------
import threading


def free_run(label, n=6):
    store_ptr, copy_ptr, view = sh._destination(1, H, W, False)
    p = R.default_params(max_depth=10, band_rows=4, n_ranks=world, rank=rank)
    o = abi.Outputs()
    o.memory, o.frame_mode, o.frame_rgba8 = abi.RTX_MEM_DEVICE, abi.RTX_FRAME_STORE, store_ptr
    for _ in range(2):
        r.render_raw([pod], p, o)
        sh._barrier()
    dist.barrier()
    torch.cuda.synchronize()
    t_sync = time.time()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    ks, host = [], []
    for k in range(n):
        h0 = time.time()
        flush.add_(1)
        ev[k][0].record()
        h1 = time.time()
        st = r.render_raw([pod], p, o)
        h2 = time.time()
        ev[k][1].record()
        sh._barrier()
        ev[k][2].record()
        h3 = time.time()
        ks.append(st.raytracing_ms)
        host.append((h0 - t_sync, h1 - h0, h2 - h1, h3 - h2))
    torch.cuda.synchronize()
    for q in range(world):
        dist.barrier()
        if q == rank:
            print("%s rank %d: " % (label, rank) + "  ".join("[render %.2f (kernel %.2f) barrier %.2f | host: start +%.2f ms, flush+record %.2f, render_raw %.2f, barrier call %.2f]" % (
                ev[k][0].elapsed_time(ev[k][1]), ks[k], ev[k][1].elapsed_time(ev[k][2]),
                host[k][0] * 1e3, host[k][1] * 1e3, host[k][2] * 1e3, host[k][3] * 1e3) for k in range(0, 3)), flush=True)


free_run("free-running")
stop = threading.Event()


def poll():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(0.02)


if rank == 0:
    t = threading.Thread(target=poll, daemon=True)
    t.start()
free_run("free-running + NVML polling on rank 0")
stop.set()
sh.close()
r.close()
dist.destroy_process_group()
