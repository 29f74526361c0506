"""ctypes binding of librtx_b200.so (include/rtx_b200.h) plus the `rt_scene`-shaped convenience call.

This module is host plumbing only: every pixel is computed by the CUDA kernels behind the C ABI.
There is no CPU fallback — if the library is missing or no B200 is visible, calls raise.

    r = Renderer(device=0)
    r.set_scene(scene.default_scene())
    out = r.render([scene.default_camera().pod()], want=("rgba8", "object_id"))     # numpy planes [F][rows][W]

`rt_scene(u, scene, cam, frame_buffer)` mirrors the reference's entry point (main.cpp:124-125).
"""
import ctypes as C
import os

import numpy as np

from . import abi, scene as scene_mod

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTX_B200_LIB lets a developer time an alternative build of the same library (tools/variants.sh)
LIB_PATH = os.environ.get("RTX_B200_LIB") or os.path.join(_HERE, "librtx_b200.so")

_lib = None


class RtxError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("rtx status %d (%s): %s" % (status, _status_name(status), message))
        self.status = status


def _status_name(status):
    return {0: "ok", 1: "invalid", 2: "cuda", 3: "no_scene", 4: "nomem"}.get(status, "?")


def load_library(path=LIB_PATH):
    """Loads the C-ABI library and declares every entry point of include/rtx_b200.h."""
    global _lib
    if _lib is not None and path == LIB_PATH:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the hot path)" % path)
    lib = C.CDLL(path)
    ctx = C.c_void_p
    lib.rtx_abi_version.restype = C.c_int
    lib.rtx_abi_version.argtypes = []
    lib.rtx_status_string.restype = C.c_char_p
    lib.rtx_status_string.argtypes = [C.c_int]
    lib.rtx_create.restype = C.c_int
    lib.rtx_create.argtypes = [C.POINTER(ctx), C.c_int]
    lib.rtx_destroy.restype = None
    lib.rtx_destroy.argtypes = [ctx]
    lib.rtx_last_error.restype = C.c_char_p
    lib.rtx_last_error.argtypes = [ctx]
    lib.rtx_set_stream.restype = C.c_int
    lib.rtx_set_stream.argtypes = [ctx, C.c_void_p]
    lib.rtx_set_scene.restype = C.c_int
    lib.rtx_set_scene.argtypes = [ctx, C.POINTER(abi.ObjectPOD), C.c_int32]
    lib.rtx_camera_init.restype = C.c_int
    lib.rtx_camera_init.argtypes = [C.POINTER(abi.CameraDesc), C.POINTER(abi.CameraPOD)]
    lib.rtx_default_params.restype = None
    lib.rtx_default_params.argtypes = [C.POINTER(abi.Params)]
    lib.rtx_local_rows.restype = C.c_int32
    lib.rtx_local_rows.argtypes = [C.c_int32] * 4
    lib.rtx_global_row.restype = C.c_int32
    lib.rtx_global_row.argtypes = [C.c_int32] * 5
    lib.rtx_render.restype = C.c_int
    lib.rtx_render.argtypes = [ctx, C.POINTER(abi.CameraPOD), C.c_int32, C.POINTER(abi.Params),
                               C.POINTER(abi.Outputs), C.POINTER(abi.Stats)]
    lib.rtx_render_async.restype = C.c_int
    lib.rtx_render_async.argtypes = [ctx, C.POINTER(abi.CameraPOD), C.c_int32, C.POINTER(abi.Params), C.POINTER(abi.Outputs)]
    lib.rtx_wait.restype = C.c_int
    lib.rtx_wait.argtypes = [ctx, C.POINTER(abi.Stats)]
    lib.rtx_trace_rays.restype = C.c_int
    lib.rtx_trace_rays.argtypes = [ctx, C.POINTER(abi.RayPOD), C.c_int64, C.POINTER(abi.Params), C.POINTER(abi.Outputs),
                                   C.POINTER(abi.Stats)]
    lib.rtx_enable_peer_access.restype = C.c_int
    lib.rtx_enable_peer_access.argtypes = [ctx, C.c_int]
    lib.rtx_device_count.restype = C.c_int
    lib.rtx_device_count.argtypes = []
    lib.rtx_host_alloc.restype = C.c_int
    lib.rtx_host_alloc.argtypes = [ctx, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.rtx_host_free.restype = C.c_int
    lib.rtx_host_free.argtypes = [ctx, C.c_void_p]
    lib.rtx_host_register.restype = C.c_int
    lib.rtx_host_register.argtypes = [ctx, C.c_void_p, C.c_uint64]
    lib.rtx_host_unregister.restype = C.c_int
    lib.rtx_host_unregister.argtypes = [ctx, C.c_void_p]
    lib.rtx_host_device_pointer.restype = C.c_int
    lib.rtx_host_device_pointer.argtypes = [ctx, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.rtx_host_shared_open.restype = C.c_int
    lib.rtx_host_shared_open.argtypes = [ctx, C.c_char_p, C.c_uint64, C.c_int32, C.POINTER(C.c_void_p)]
    lib.rtx_host_shared_close.restype = C.c_int
    lib.rtx_host_shared_close.argtypes = [ctx, C.c_void_p, C.c_char_p]
    lib.rtx_quantise.restype = C.c_int
    lib.rtx_quantise.argtypes = [ctx, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                 C.POINTER(abi.Stats)]
    lib.rtx_tonemap.restype = C.c_int
    lib.rtx_tonemap.argtypes = [ctx, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(abi.Params), C.c_void_p, C.c_int32,
                                C.POINTER(C.c_double), C.POINTER(abi.Stats)]
    lib.rtx_tonemap_sums.restype = C.c_int
    lib.rtx_tonemap_sums.argtypes = [ctx, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.rtx_tonemap_apply.restype = C.c_int
    lib.rtx_tonemap_apply.argtypes = [ctx, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(abi.Params),
                                      C.c_void_p, C.POINTER(abi.Stats)]
    lib.rtx_unpermute_bands.restype = C.c_int
    lib.rtx_unpermute_bands.argtypes = [ctx, C.c_void_p, C.c_void_p] + [C.c_int32] * 6
    lib.rtx_ffma_peak.restype = C.c_int
    lib.rtx_ffma_peak.argtypes = [ctx, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in ("rtx_buffer_free", "rtx_buffer_release"):
        getattr(lib, name).restype = C.c_int
        getattr(lib, name).argtypes = [ctx, C.c_void_p]
    lib.rtx_buffer_alloc.restype = C.c_int
    lib.rtx_buffer_alloc.argtypes = [ctx, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.rtx_buffer_export.restype = C.c_int
    lib.rtx_buffer_export.argtypes = [ctx, C.c_void_p, C.c_char_p]
    lib.rtx_buffer_read.restype = C.c_int
    lib.rtx_buffer_read.argtypes = [ctx, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.rtx_buffer_import.restype = C.c_int
    lib.rtx_buffer_import.argtypes = [ctx, C.c_char_p, C.POINTER(C.c_void_p)]
    if lib.rtx_abi_version() != abi.ABI_VERSION:
        raise RuntimeError("librtx_b200.so ABI %d != binding ABI %d" % (lib.rtx_abi_version(), abi.ABI_VERSION))
    if path == LIB_PATH:
        _lib = lib
    return lib


def default_params(**overrides):
    """rtx_params with the reference's literals (main.cpp:14-17,34,89,111), then keyword overrides."""
    p = abi.Params()
    load_library().rtx_default_params(C.byref(p))
    for k, v in overrides.items():
        if k in ("light_pos", "ground_color", "sky_low", "sky_high", "sun_color", "sun_direction"):
            v = abi.Vec3(*v)
        if not hasattr(p, k):
            raise AttributeError("rtx_params has no field %r" % k)
        setattr(p, k, v)
    return p


def camera_init(cam):
    """Camera::init through the library (host code of the C ABI). cam: scene.Camera."""
    out = abi.CameraPOD()
    d = cam.desc()
    rc = load_library().rtx_camera_init(C.byref(d), C.byref(out))
    if rc != abi.RTX_OK:
        raise RtxError(rc, "rtx_camera_init")
    return out


def local_rows(height, band_rows, n_ranks, rank):
    return load_library().rtx_local_rows(height, band_rows, n_ranks, rank)


def global_rows(height, band_rows, n_ranks, rank):
    """Global row index of every packed local row of `rank` (numpy int32)."""
    lib = load_library()
    n = lib.rtx_local_rows(height, band_rows, n_ranks, rank)
    return np.array([lib.rtx_global_row(k, height, band_rows, n_ranks, rank) for k in range(n)], dtype=np.int32)


_PLANES = {  # name -> (Outputs field, numpy dtype, trailing shape)
    "rgba8": ("rgba8", np.uint32, ()),
    "radiance_f32": ("radiance_f32", np.float32, (3,)),
    "radiance_f64": ("radiance_f64", np.float64, (3,)),
    "object_id": ("object_id", np.int32, ()),
    "hit_mask": ("hit_mask", np.uint8, ()),
    "ray_count": ("ray_count", np.uint8, ()),
    "hit_distance": ("hit_distance", np.float64, ()),
    "hit_normal": ("hit_normal", np.float64, (3,)),
}


class Renderer:
    """One rtx_ctx (one GPU)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        rc = self.lib.rtx_create(C.byref(self._ctx), int(device))
        if rc != abi.RTX_OK:
            msg = self.lib.rtx_last_error(None)
            raise RtxError(rc, (msg or b"").decode() or "rtx_create failed")
        self.device = device
        self.n_objects = 0
        self.last_stats = None

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.rtx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != abi.RTX_OK:
            raise RtxError(rc, (self.lib.rtx_last_error(self._ctx) or b"").decode())

    def set_stream(self, cuda_stream_handle):
        """cuda_stream_handle: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream; 0 is CUDA's
        default stream) or None for the context's private stream."""
        h = C.c_void_p(-1) if cuda_stream_handle is None else C.c_void_p(int(cuda_stream_handle))
        self._check(self.lib.rtx_set_stream(self._ctx, h))

    def set_scene(self, scene):
        """scene: list of scene.Sphere / scene.Wall in scene order, or a ctypes array of rtx_object."""
        if isinstance(scene, C.Array):
            objs, n = scene, len(scene)
        else:
            objs, n = scene_mod.flatten(scene), len(scene)
        self._check(self.lib.rtx_set_scene(self._ctx, objs, n))
        self.n_objects = n

    def render_raw(self, cams, params, outputs):
        """Direct rtx_render with caller-built ctypes structures (device or host pointers). Returns Stats."""
        arr = (abi.CameraPOD * len(cams))(*cams)
        st = abi.Stats()
        self._check(self.lib.rtx_render(self._ctx, arr, len(cams), C.byref(params), C.byref(outputs), C.byref(st)))
        self.last_stats = st
        return st

    def render_async(self, cams, params, outputs):
        """rtx_render_async: queues the call and returns; wait() completes the oldest call in flight (at most
        abi.RTX_MAX_IN_FLIGHT)."""
        arr = (abi.CameraPOD * len(cams))(*cams)
        self._check(self.lib.rtx_render_async(self._ctx, arr, len(cams), C.byref(params), C.byref(outputs)))

    def wait(self):
        st = abi.Stats()
        self._check(self.lib.rtx_wait(self._ctx, C.byref(st)))
        self.last_stats = st
        return st

    def trace_rays(self, rays, params=None, want=("radiance_f64", "object_id", "hit_distance", "hit_normal", "ray_count")):
        """rtx_trace_rays: recursive_ray_tracing / find_closest_hit (main.cpp:67-119) for a batch of rays.
        rays: sequence of (origin, direction) triples. Returns (numpy planes [n_rays](+[3]), Stats)."""
        params = params if params is not None else default_params()
        n = len(rays)
        arr = (abi.RayPOD * n)()
        for k, (o, d) in enumerate(rays):
            arr[k].origin, arr[k].direction = abi.Vec3(*o), abi.Vec3(*d)
        o = abi.Outputs()
        o.memory = abi.RTX_MEM_HOST
        planes = {}
        for name in want:
            field, dtype, tail = _PLANES[name]
            planes[name] = np.empty((n,) + tail, dtype)
            setattr(o, field, planes[name].ctypes.data)
        st = abi.Stats()
        self._check(self.lib.rtx_trace_rays(self._ctx, arr, n, C.byref(params), C.byref(o), C.byref(st)))
        self.last_stats = st
        return planes, st

    # -- pinned / mapped / shared host memory (zero-copy surfaces, multi-process host frames) ------------------
    def host_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.rtx_host_alloc(self._ctx, int(nbytes), C.byref(p)))
        return p.value

    def host_free(self, ptr):
        self._check(self.lib.rtx_host_free(self._ctx, C.c_void_p(ptr)))

    def host_register(self, ptr, nbytes):
        self._check(self.lib.rtx_host_register(self._ctx, C.c_void_p(ptr), int(nbytes)))

    def host_unregister(self, ptr):
        self._check(self.lib.rtx_host_unregister(self._ctx, C.c_void_p(ptr)))

    def host_device_pointer(self, ptr):
        p = C.c_void_p()
        self._check(self.lib.rtx_host_device_pointer(self._ctx, C.c_void_p(ptr), C.byref(p)))
        return p.value

    def host_shared_open(self, name, nbytes, create):
        p = C.c_void_p()
        self._check(self.lib.rtx_host_shared_open(self._ctx, name.encode(), int(nbytes), 1 if create else 0, C.byref(p)))
        return p.value

    def host_shared_close(self, ptr, unlink_name=None):
        self._check(self.lib.rtx_host_shared_close(self._ctx, C.c_void_p(ptr), unlink_name.encode() if unlink_name else None))

    def render(self, cams, params=None, want=("rgba8",), out=None):
        """Renders len(cams) frames into host numpy planes shaped [F][rows][W](+[3]).
        `out` may carry preallocated (e.g. pinned) arrays by plane name. Returns (planes dict, Stats)."""
        if isinstance(cams, abi.CameraPOD):
            cams = [cams]
        params = params if params is not None else default_params()
        W, H = cams[0].width, cams[0].height
        rows = H if params.n_ranks <= 1 else self.lib.rtx_local_rows(H, params.band_rows, params.n_ranks, params.rank)
        o = abi.Outputs()
        o.memory = abi.RTX_MEM_HOST
        planes = {}
        for name in want:
            field, dtype, tail = _PLANES[name]
            shape = (len(cams), rows, W) + tail
            a = out[name] if out is not None and name in out else np.empty(shape, dtype)
            if a.dtype != dtype or a.size != int(np.prod(shape)) or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("output plane %s must be C-contiguous %s of shape %s" % (name, dtype, shape))
            planes[name] = a.reshape(shape)
            setattr(o, field, a.ctypes.data)
        st = self.render_raw(cams, params, o)
        return planes, st

    def quantise(self, radiance, mode=abi.RTX_QUANT_WRAP):
        """Standalone quantise pass (main.cpp:338-347) on a host float32/float64 array of RGB triples."""
        rad = np.ascontiguousarray(radiance)
        if rad.dtype not in (np.float32, np.float64):
            raise ValueError("radiance must be float32 or float64")
        n = rad.size // 3
        out = np.empty(n, np.uint32)
        st = abi.Stats()
        p32 = rad.ctypes.data if rad.dtype == np.float32 else None
        p64 = rad.ctypes.data if rad.dtype == np.float64 else None
        self._check(self.lib.rtx_quantise(self._ctx, p32, p64, n, int(mode), out.ctypes.data, abi.RTX_MEM_HOST,
                                          C.byref(st)))
        self.last_stats = st
        return out.reshape(rad.shape[:-1]) if rad.ndim > 1 else out

    def tonemap(self, radiance, params):
        """EXTENSION (rtx_tonemap): Reinhard's global operator + 8-bit pack on a host float32/float64 array shaped
        [n_frames][pixels][3] (or [n_frames][H][W][3]). Returns (rgba8 words shaped like the input minus the last axis,
        log-average luminance per frame)."""
        rad = np.ascontiguousarray(radiance)
        if rad.dtype not in (np.float32, np.float64) or rad.ndim < 3 or rad.shape[-1] != 3:
            raise ValueError("radiance must be float32/float64 shaped [n_frames][...][3]")
        n_frames = rad.shape[0]
        pixels = rad.size // 3 // n_frames
        out = np.empty(rad.shape[:-1], np.uint32)
        lavg = np.zeros(n_frames, np.float64)
        st = abi.Stats()
        p32 = rad.ctypes.data if rad.dtype == np.float32 else None
        p64 = rad.ctypes.data if rad.dtype == np.float64 else None
        self._check(self.lib.rtx_tonemap(self._ctx, p32, p64, pixels, n_frames, C.byref(params), out.ctypes.data, abi.RTX_MEM_HOST,
                                         lavg.ctypes.data_as(C.POINTER(C.c_double)), C.byref(st)))
        self.last_stats = st
        return out, lavg

    def tonemap_device(self, rad_ptr, is_f32, pixels_per_frame, n_frames, params, out_ptr):
        st = abi.Stats()
        self._check(self.lib.rtx_tonemap(self._ctx, rad_ptr if is_f32 else None, None if is_f32 else rad_ptr, int(pixels_per_frame),
                                         int(n_frames), C.byref(params), out_ptr, abi.RTX_MEM_DEVICE, None, C.byref(st)))
        self.last_stats = st
        return st

    def tonemap_sums_device(self, rad_ptr, is_f32, pixels_per_frame, n_frames, sums_ptr):
        """Step 1 of the sharded tone map: adds this buffer's per-frame fixed-point sums to the int64 device array."""
        self._check(self.lib.rtx_tonemap_sums(self._ctx, rad_ptr if is_f32 else None, None if is_f32 else rad_ptr, int(pixels_per_frame),
                                              int(n_frames), sums_ptr))

    def tonemap_apply_device(self, rad_ptr, is_f32, pixels_per_frame, n_frames, sums_ptr, pixels_per_frame_global, params, out_ptr):
        """Step 3 of the sharded tone map: maps and packs with the (all-reduced) sums over pixels_per_frame_global pixels."""
        st = abi.Stats()
        self._check(self.lib.rtx_tonemap_apply(self._ctx, rad_ptr if is_f32 else None, None if is_f32 else rad_ptr, int(pixels_per_frame),
                                               int(n_frames), sums_ptr, int(pixels_per_frame_global), C.byref(params), out_ptr, C.byref(st)))
        self.last_stats = st
        return st

    def quantise_device(self, rad_ptr, is_f32, n_pixels, out_ptr, mode=abi.RTX_QUANT_WRAP):
        st = abi.Stats()
        self._check(self.lib.rtx_quantise(self._ctx, rad_ptr if is_f32 else None, None if is_f32 else rad_ptr,
                                          int(n_pixels), int(mode), out_ptr, abi.RTX_MEM_DEVICE, C.byref(st)))
        self.last_stats = st
        return st

    def unpermute_bands(self, band_major_ptr, row_major_ptr, height, width, elem_bytes, band_rows, n_ranks,
                        rows_per_rank):
        self._check(self.lib.rtx_unpermute_bands(self._ctx, band_major_ptr, row_major_ptr, height, width, elem_bytes,
                                                 band_rows, n_ranks, rows_per_rank))

    # -- shareable device buffers (fused multi-GPU gather) --------------------------------------------
    def buffer_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.rtx_buffer_alloc(self._ctx, int(nbytes), C.byref(p)))
        return p.value

    def buffer_free(self, ptr):
        self._check(self.lib.rtx_buffer_free(self._ctx, C.c_void_p(ptr)))

    def buffer_export(self, ptr):
        h = C.create_string_buffer(64)
        self._check(self.lib.rtx_buffer_export(self._ctx, C.c_void_p(ptr), h))
        return h.raw

    def buffer_import(self, handle):
        p = C.c_void_p()
        self._check(self.lib.rtx_buffer_import(self._ctx, C.create_string_buffer(bytes(handle), 64), C.byref(p)))
        return p.value

    def buffer_read(self, ptr, host_array):
        """Blocking copy of host_array.nbytes bytes from a device pointer into a C-contiguous numpy array."""
        self._check(self.lib.rtx_buffer_read(self._ctx, C.c_void_p(ptr), host_array.ctypes.data, host_array.nbytes))
        return host_array

    def buffer_release(self, ptr):
        self._check(self.lib.rtx_buffer_release(self._ctx, C.c_void_p(ptr)))

    def ffma_peak(self, variant=0):
        t, m = C.c_double(), C.c_double()
        self._check(self.lib.rtx_ffma_peak(self._ctx, variant, C.byref(t), C.byref(m)))
        return t.value, m.value


_default_renderer = None


def rt_scene(u, scene, cam, frame_buffer, max_depth=10):
    """Drop-in shape of the reference's `rt_scene(u, scene, cam, frame_buffer)` (main.cpp:124-139):
    u = [pixel_delta_x, pixel_delta_y] as returned by cam.init(); scene = list of Sphere/Wall; cam = scene.Camera;
    frame_buffer = float64 array [image_height][image_width][3], overwritten in place (row i, column j)."""
    global _default_renderer
    if _default_renderer is None:
        _default_renderer = Renderer(0)
    pod = cam.pod()
    pod.delta_x, pod.delta_y = abi.Vec3(*u[0]), abi.Vec3(*u[1])
    fb = np.asarray(frame_buffer)
    if fb.dtype != np.float64 or fb.shape != (pod.height, pod.width, 3) or not fb.flags["C_CONTIGUOUS"]:
        raise ValueError("frame_buffer must be a C-contiguous float64 array [image_height][image_width][3]")
    _default_renderer.set_scene(scene)
    _default_renderer.render([pod], default_params(max_depth=max_depth), want=("radiance_f64",),
                             out={"radiance_f64": fb})
    return fb


def write_png(path, rgba8):
    """Headless output: 8-bit RGB PNG from RGBA8888 words (alpha is always 0xFF in the reference's surface)."""
    import struct
    import zlib
    a = np.asarray(rgba8, dtype=np.uint32)
    h, w = a.shape
    rgb = np.stack([(a >> 24) & 0xFF, (a >> 16) & 0xFF, (a >> 8) & 0xFF], axis=-1).astype(np.uint8)
    raw = b"".join(b"\x00" + rgb[i].tobytes() for i in range(h))      # filter type 0 per scanline

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def write_ppm(path, rgba8):
    """Headless output (replaces the SDL present, main.cpp:351-358): binary PPM from RGBA8888 words."""
    a = np.asarray(rgba8, dtype=np.uint32)
    h, w = a.shape
    rgb = np.stack([(a >> 24) & 0xFF, (a >> 16) & 0xFF, (a >> 8) & 0xFF], axis=-1).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(rgb.tobytes())
