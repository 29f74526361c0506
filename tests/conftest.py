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


@pytest.fixture(scope="session")
def port(ob):
    """The plain-C oracle (oracle/oracle.c); built on demand (gcc only)."""
    if not os.path.exists(ob.PORT_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return ob.load_port()


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
