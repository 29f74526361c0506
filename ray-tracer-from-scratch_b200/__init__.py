"""B200-native ray-tracing hot path. Import with importlib:

    rtx = importlib.import_module("ray-tracer-from-scratch_b200")

Submodules: `abi` (ctypes mirror of include/rtx_b200.h), `scene` (host-side mirror of the reference's
Material/Sphere/Wall/Camera), `renderer` (the C-ABI binding; needs the built CUDA library).
"""
from . import abi, scene  # noqa: F401
