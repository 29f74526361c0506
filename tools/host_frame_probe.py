#!/usr/bin/env python
"""Developer probe (GPU box): device -> host copy rate into the three kinds of pinned host memory the library hands out —
rtx_host_alloc (cudaHostAlloc), rtx_host_shared_open (POSIX shared memory + cudaHostRegister: the multi-process host frame)
and rtx_host_register of plain malloc'ed memory — to see whether the shared frame itself limits the camera-path read-back."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
R = importlib.import_module("ray-tracer-from-scratch_b200.renderer")

n = 512 << 20
r = R.Renderer(0)
src = torch.zeros(n // 4, dtype=torch.int32, device="cuda")
plain = np.zeros(n, np.uint8)
kinds = {"rtx_host_alloc (cudaHostAlloc)": r.host_alloc(n), "rtx_host_shared_open (shm + cudaHostRegister)": r.host_shared_open("/rtx_probe_%d" % os.getpid(), n, True)}
r.host_register(plain.ctypes.data, n)
kinds["rtx_host_register (malloc + cudaHostRegister)"] = plain.ctypes.data
for name, ptr in kinds.items():
    dst = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,))
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r.buffer_read(src.data_ptr(), dst)
        best = min(best, time.perf_counter() - t0)
    print("%-50s %.1f GB/s" % (name, n / best / 1e9), flush=True)
os.unlink("/dev/shm/rtx_probe_%d" % os.getpid())
for p in ("/sys/kernel/mm/transparent_hugepage/shmem_enabled", "/sys/kernel/mm/transparent_hugepage/enabled", "/proc/sys/vm/nr_hugepages"):
    try:
        print(p, open(p).read().strip())
    except Exception as e:
        print(p, "unreadable", e)
