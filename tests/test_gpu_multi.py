"""GPU suite: more than one context at a time — the C++ multi-GPU host and two contexts driven from two host threads.

The C++ host (host/rtx_scene.hpp, rtx::ShardedRenderer) runs one rtx_ctx and one host thread per GPU in ONE process
and lets every trace kernel store its cyclic row bands straight into one pinned host surface; north_star: "the host side
stays C++". A device may be listed twice (two contexts on one GPU), so the whole path runs on a one-GPU box too; with two
or more GPUs visible the same test also runs across real devices.
"""
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def headless(renderer_mod):
    exe = os.path.join(os.path.dirname(renderer_mod.LIB_PATH), "rtx_headless")
    assert os.path.exists(exe), "build() must produce rtx_headless"
    return exe


def n_devices(renderer_mod):
    return renderer_mod.load_library().rtx_device_count()


@pytest.mark.parametrize("devices,band", [("0,0", 4), ("0,0,0", 3), ("0,1", 4), ("0,1,2,3", 1)])
def test_cpp_sharded_host_assembles_the_reference_frame(tmp_path, renderer_mod, port, S, devices, band):
    if max(int(d) for d in devices.split(",")) >= n_devices(renderer_mod):
        pytest.skip("needs %s" % devices)
    exe = headless(renderer_mod)
    raw = tmp_path / "multi.rgba"
    # default scene, square frame whose height (250) is not a multiple of band * ranks: ragged last bands
    out = subprocess.run([exe, "--width", "250", "--frames", "2", "--keys", "xw", "--devices", devices, "--band-rows", str(band),
                          "--raw", str(raw), "--out", ""], check=True, capture_output=True, text=True).stdout
    assert "microseconds for average raytracing" in out and "GPUs" in out
    pod = S.default_camera(250, 1.0).pod()
    pod.position = type(pod.position)(0.1, 0.0, 0.0)                    # one 'w' key: Camera::forward, init() not re-run
    exp = port.render(S.default_scene(), pod, 10, want=("rgba8",))["rgba8"]
    assert np.array_equal(np.fromfile(raw, dtype=np.uint32).reshape(250, 250), exp)
    # the 10 064-object scene (big kernel, 186 KB of shared memory per CTA, peer / mapped stores): same frame as one context
    one, many = tmp_path / "one.rgba", tmp_path / "many.rgba"
    common = [exe, "--scene", "synthetic", "--width", "192", "--aspect", "1.7777777777777777", "--depth", "6", "--frames", "1", "--out", ""]
    subprocess.run(common + ["--devices", "0", "--raw", str(one)], check=True, capture_output=True)
    subprocess.run(common + ["--devices", devices, "--band-rows", str(band), "--raw", str(many)], check=True, capture_output=True)
    a, b = np.fromfile(one, dtype=np.uint32), np.fromfile(many, dtype=np.uint32)
    assert a.size == 192 * 108 and np.array_equal(a, b)
    # the same frame assembled in the first GPU's memory (peer stores, rtx_enable_peer_access) and read back with rtx_buffer_read
    dev = tmp_path / "dev.rgba"
    subprocess.run(common + ["--devices", devices, "--band-rows", str(band), "--to-device", "1", "--raw", str(dev)], check=True, capture_output=True)
    assert np.array_equal(np.fromfile(dev, dtype=np.uint32), a)
    exp = port.render(S.synthetic_scene(), S.default_camera(192, 16.0 / 9.0).pod(), 6, want=("rgba8",))["rgba8"]
    assert np.array_equal(a.reshape(108, 192), exp)                     # and the C++ scene generator draws the survey's scene


def test_two_contexts_from_two_threads_need_large_shared_memory_each(renderer_mod, port, S):
    """One host thread per context, both rendering the 10 064-object scene (161 KB of dynamic shared memory): the
    per-device function attribute is raised per CONTEXT, not per thread or process. With two GPUs the contexts sit on
    different devices; with one they share it."""
    nd = n_devices(renderer_mod)
    devs = [0, 1 if nd > 1 else 0]
    scene = S.synthetic_scene()
    pod = S.default_camera(96, 16.0 / 9.0).pod()
    exp = port.render(scene, pod, 4, want=("rgba8",))["rgba8"]
    results, errors = {}, []

    def work(k):
        try:
            with renderer_mod.Renderer(devs[k]) as r:
                r.set_scene(scene)
                for _ in range(2):
                    planes, _ = r.render([pod], renderer_mod.default_params(max_depth=4), want=("rgba8",))
                results[k] = planes["rgba8"][0]
        except Exception as e:          # surfaces in the main thread
            errors.append(e)

    # also: ONE thread driving both contexts in turn (what the advisor flagged: a thread-local cache of the attribute)
    with renderer_mod.Renderer(devs[0]) as r0, renderer_mod.Renderer(devs[1]) as r1:
        for r in (r0, r1):
            r.set_scene(scene)
            planes, _ = r.render([pod], renderer_mod.default_params(max_depth=4), want=("rgba8",))
            assert np.array_equal(planes["rgba8"][0], exp)
    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert np.array_equal(results[0], exp) and np.array_equal(results[1], exp)
