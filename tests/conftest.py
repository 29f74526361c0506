import ctypes
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("ray-tracer-from-scratch_b200")


@pytest.fixture(scope="session")
def S(pkg):
    return pkg.scene


@pytest.fixture(scope="session")
def ob():
    from oracle import binding
    return binding


class PinnedOracle:
    """The plain-C oracle, pinned AT RUN TIME: every frame it renders with the reference's own parameters is rendered a
    second time by oracle/_ref — the unmodified reference sources — and must agree bit for bit (ids, masks, ray counts,
    RGBA8 words, radiance) before a GPU test is allowed to compare the CUDA path with it. Frames the reference cannot
    express (this repo's extensions: boxes, sun, non-default parameters) go through the port alone."""

    def __init__(self, port, ref):
        self._port, self._ref = port, ref
        self.cross_checked = 0

    def __getattr__(self, name):
        return getattr(self._port, name)

    def render(self, scene, cam_pod, max_depth=10, rows=None, threads=0,
               want=("radiance", "rgba8", "object_id", "hit_mask", "ray_count"), params=None):
        out = self._port.render(scene, cam_pod, max_depth, rows=rows, threads=threads, want=want, params=params)
        if self._ref is not None and params is None and not isinstance(scene, ctypes.Array) and all(g.kind in (0, 1) for g in scene):
            again = self._ref.render(scene, cam_pod, max_depth, rows=rows, threads=threads, want=want)
            for k in want:
                a, b = out[k], again[k]
                same = np.array_equal(a.view(np.uint64), b.view(np.uint64)) if a.dtype == np.float64 else np.array_equal(a, b)
                if not same and a.dtype == np.float64:
                    same = bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())
                assert same, "oracle.c and the unmodified reference disagree on plane %r" % k
            if any(k in want for k in ("object_id", "hit_mask", "ray_count")):
                assert out["total_rays"] == again["total_rays"]
            self.cross_checked += 1
        return out


@pytest.fixture(scope="session")
def port(ob):
    """The plain-C oracle (oracle/oracle.c; built on demand, gcc only), cross-checked against oracle/_ref on every frame
    when the compiled reference is present (it travels to the GPU box with the snapshot)."""
    if not os.path.exists(ob.PORT_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    ref = None
    if not os.path.exists(ob.REF_PATH) and os.path.exists("/root/reference/main.cpp"):
        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    if os.path.exists(ob.REF_PATH):
        ref = ob.load_reference()
    return PinnedOracle(ob.load_port(), ref)


@pytest.fixture(scope="session")
def ref(ob):
    """The unmodified reference build (oracle/_ref). Present wherever oracle/build_ref.sh has run
    (it needs /root/reference); the prebuilt .so travels to the GPU box."""
    if not os.path.exists(ob.REF_PATH):
        if os.path.exists("/root/reference/main.cpp"):
            subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return ob.load_reference()


@pytest.fixture(scope="session")
def renderer_mod(pkg):
    return importlib.import_module("ray-tracer-from-scratch_b200.renderer")


@pytest.fixture(scope="session")
def gpu(renderer_mod):
    """One Renderer on cuda:0 for the whole session. Fails loudly (no CPU fallback) if unavailable."""
    r = renderer_mod.Renderer(0)
    yield r
    r.close()


def load_golden_frame(name):
    z = np.load(os.path.join(GOLDEN, name))
    d = {k: z[k] for k in z.files}
    W, H, depth, rays = [int(x) for x in d.pop("meta")]
    d.update(width=W, height=H, depth=depth, total_rays=rays)
    return d


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def fh(x):
    return float.fromhex(x)


def fh3(t):
    return tuple(float.fromhex(x) for x in t)


def same_float(a, b):
    """Bit-level equality that treats NaN == NaN."""
    a, b = float(a), float(b)
    return (a != a and b != b) or (a == b and np.signbit(a) == np.signbit(b)) or (a == b == 0.0)
