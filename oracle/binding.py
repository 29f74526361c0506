"""ctypes doors onto the two CPU oracles — TEST INFRASTRUCTURE ONLY.

  load_port()       -> oracle/liboracle.so            (oracle.c, the plain-C restatement, prefix orc_)
  load_reference()  -> oracle/_ref/libref_oracle.so   (the unmodified reference sources + harness, prefix ref_)

Both export the same doors (render_rows, camera_init, quantise, intersect, find_closest_hit, trace_ray,
out_color, reflect, diffuse, specular), so one wrapper class serves both. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package never does.
"""
import ctypes as C
import importlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "liboracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libref_oracle.so")

_pkg = importlib.import_module("ray-tracer-from-scratch_b200")
abi = _pkg.abi
scene_mod = _pkg.scene


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


class Oracle:
    def __init__(self, path, prefix, kind):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = kind  # "port" | "reference"
        V = C.POINTER(abi.Vec3)
        f = self._f
        f("abi_version", C.c_int)
        f("max_threads", C.c_int)
        f("camera_init", None, C.POINTER(abi.CameraDesc), C.POINTER(abi.CameraPOD))
        f("camera_walk", None, C.POINTER(abi.CameraDesc), C.c_char_p, C.POINTER(C.c_double), C.c_int32,
          C.POINTER(C.c_double), C.POINTER(abi.CameraPOD))
        f("render_rows", C.c_double, C.POINTER(abi.ObjectPOD), C.c_int32, C.POINTER(abi.CameraPOD), C.c_int32,
          C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint32),
          C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64))
        f("quantise", None, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_uint32))
        f("intersect", None, C.POINTER(abi.ObjectPOD), V, V, C.POINTER(C.c_double), V, C.POINTER(C.c_int32))
        f("find_closest_hit", None, C.POINTER(abi.ObjectPOD), C.c_int32, V, V, C.POINTER(C.c_double), V,
          C.POINTER(C.c_int32))
        f("trace_ray", None, C.POINTER(abi.ObjectPOD), C.c_int32, V, V, C.c_int32, V)
        f("out_color", None, V, V)
        f("reflect", None, V, V, V)
        f("diffuse", C.c_double, V, V, V)
        f("specular", C.c_double, V, V, V, V)
        if kind == "reference":
            f("rt_scene", C.c_double, C.POINTER(abi.ObjectPOD), C.c_int32, C.POINTER(abi.CameraDesc),
              C.POINTER(C.c_double))
            f("run_main", C.c_int, C.c_int32, C.POINTER(C.c_uint32), C.c_int32, C.POINTER(C.c_int32),
              C.POINTER(C.c_int32))
        else:
            f("default_params", None, C.POINTER(abi.Params))
            f("render_rows_params", C.c_double, C.POINTER(abi.ObjectPOD), C.c_int32, C.POINTER(abi.CameraPOD),
              C.POINTER(abi.Params), C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_double),
              C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8),
              C.POINTER(C.c_uint64))
            f("quantise_mode", None, C.POINTER(C.c_double), C.c_int64, C.c_int32, C.POINTER(C.c_uint32))
            f("quantise_f32", None, C.POINTER(C.c_float), C.c_int64, C.c_int32, C.POINTER(C.c_uint32))
            f("tonemap", None, C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_int64, C.c_int32, C.POINTER(abi.Params),
              C.POINTER(C.c_uint32), C.POINTER(C.c_double))
            f("tonemap_sums", None, C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_int64, C.c_int32, C.POINTER(C.c_int64))
            f("tonemap_apply", None, C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.c_int64,
              C.POINTER(abi.Params), C.POINTER(C.c_uint32), C.POINTER(C.c_double))

    def _f(self, name, restype, *argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = list(argtypes)
        setattr(self, "_" + name, fn)

    # -- doors -------------------------------------------------------------------------------------
    def max_threads(self):
        return self._max_threads()

    def camera_init(self, cam):
        """cam: scene.Camera (inputs only are read). Returns abi.CameraPOD computed by the oracle."""
        out = abi.CameraPOD()
        d = cam.desc()
        self._camera_init(C.byref(d), C.byref(out))
        return out

    def camera_walk(self, cam, steps):
        """Camera::init once, then one Camera method per step (scene.cpp:108-165); steps = [(op, arg), ...] with op in
        'wsad' (moves), 'y' (rotate_left_right(arg)), 'p' (rotate_up_down(arg)). Returns (states [n][3][3] =
        position, direction, vup after every step; abi.CameraPOD after the last step)."""
        ops = "".join(op for op, _ in steps).encode()
        args = np.ascontiguousarray([float(a) for _, a in steps], dtype=np.float64)
        states = np.zeros((len(steps), 3, 3), np.float64)
        out = abi.CameraPOD()
        d = cam.desc()
        self._camera_walk(C.byref(d), ops, _ptr(args, C.c_double), len(steps), _ptr(states, C.c_double), C.byref(out))
        return states, out

    def render(self, scene, cam_pod, max_depth=10, rows=None, threads=0, want=("radiance", "rgba8", "object_id",
                                                                                 "hit_mask", "ray_count"),
               params=None):
        """Renders the global rows `rows` (default: all) of one frame. Returns a dict of numpy planes packed
        [n_rows][W], plus 'seconds' and 'total_rays'."""
        objs = scene if isinstance(scene, C.Array) else scene_mod.flatten(scene)
        n = len(scene)
        W, H = cam_pod.width, cam_pod.height
        rows = np.arange(H, dtype=np.int32) if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        nr = len(rows)
        out = {
            "radiance": np.zeros((nr, W, 3), np.float64) if "radiance" in want else None,
            "rgba8": np.zeros((nr, W), np.uint32) if "rgba8" in want else None,
            "object_id": np.zeros((nr, W), np.int32) if "object_id" in want else None,
            "hit_mask": np.zeros((nr, W), np.uint8) if "hit_mask" in want else None,
            "ray_count": np.zeros((nr, W), np.uint8) if "ray_count" in want else None,
        }
        total = C.c_uint64(0)
        want_total = any(k in want for k in ("object_id", "hit_mask", "ray_count", "total_rays"))
        args_tail = (_ptr(rows, C.c_int32), nr, int(threads), _ptr(out["radiance"], C.c_double),
                     _ptr(out["rgba8"], C.c_uint32), _ptr(out["object_id"], C.c_int32), _ptr(out["hit_mask"], C.c_uint8),
                     _ptr(out["ray_count"], C.c_uint8), C.byref(total) if want_total or self.kind == "port" else None)
        if params is not None:
            if self.kind != "port":
                raise ValueError("the reference oracle has compile-time parameters only")
            secs = self._render_rows_params(objs, n, C.byref(cam_pod), C.byref(params), *args_tail)
        else:
            secs = self._render_rows(objs, n, C.byref(cam_pod), int(max_depth), *args_tail)
        if secs < 0:
            raise RuntimeError("oracle render failed")
        out = {k: v for k, v in out.items() if v is not None}
        out["seconds"] = secs
        out["total_rays"] = int(total.value)
        return out

    def default_params(self):
        p = abi.Params()
        self._default_params(C.byref(p))
        return p

    def quantise(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(rgb), np.uint32)
        self._quantise(_ptr(rgb, C.c_double), len(rgb), _ptr(out, C.c_uint32))
        return out

    def quantise_mode(self, rgb, mode):
        rgb = np.ascontiguousarray(rgb).reshape(-1, 3)
        out = np.zeros(len(rgb), np.uint32)
        if rgb.dtype == np.float32:
            self._quantise_f32(_ptr(rgb, C.c_float), len(rgb), int(mode), _ptr(out, C.c_uint32))
        else:
            rgb = rgb.astype(np.float64)
            self._quantise_mode(_ptr(rgb, C.c_double), len(rgb), int(mode), _ptr(out, C.c_uint32))
        return out

    def tonemap(self, rgb, params):
        """EXTENSION (port only): Reinhard's global operator + 8-bit pack on [n_frames][pixels][3] radiance (f32 or f64).
        Returns (rgba8 [n_frames][pixels], log-average luminance per frame)."""
        rgb = np.ascontiguousarray(rgb)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        n_frames, pixels = rgb.shape[0], rgb.shape[1]
        out = np.zeros((n_frames, pixels), np.uint32)
        lavg = np.zeros(n_frames, np.float64)
        if rgb.dtype == np.float32:
            self._tonemap(None, _ptr(rgb, C.c_float), pixels, n_frames, C.byref(params), _ptr(out, C.c_uint32), _ptr(lavg, C.c_double))
        else:
            rgb = np.ascontiguousarray(rgb, dtype=np.float64)
            self._tonemap(_ptr(rgb, C.c_double), None, pixels, n_frames, C.byref(params), _ptr(out, C.c_uint32), _ptr(lavg, C.c_double))
        return out, lavg

    def tonemap_sums(self, rgb, sums):
        """Step 1 of the sharded tone map: ADDS the per-frame fixed-point sums of rgb [n_frames][pixels][3] (float64) to
        the int64 array `sums`."""
        rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        self._tonemap_sums(_ptr(rgb, C.c_double), None, rgb.shape[1], rgb.shape[0], _ptr(sums, C.c_int64))
        return sums

    def tonemap_apply(self, rgb, sums, pixels_global, params):
        """Step 2: maps and packs rgb [n_frames][pixels][3] with sums taken over pixels_global pixels per frame."""
        rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        out = np.zeros(rgb.shape[:2], np.uint32)
        self._tonemap_apply(_ptr(rgb, C.c_double), None, rgb.shape[1], rgb.shape[0], _ptr(np.ascontiguousarray(sums, dtype=np.int64), C.c_int64),
                            int(pixels_global), C.byref(params), _ptr(out, C.c_uint32), None)
        return out

    def intersect(self, geom, origin, direction):
        obj = scene_mod.flatten([geom])
        o, d, nrm = abi.Vec3(*origin), abi.Vec3(*direction), abi.Vec3()
        dist, hit = C.c_double(), C.c_int32()
        self._intersect(obj, C.byref(o), C.byref(d), C.byref(dist), C.byref(nrm), C.byref(hit))
        return dist.value, nrm.tuple(), bool(hit.value)

    def find_closest_hit(self, scene, origin, direction):
        objs = scene if isinstance(scene, C.Array) else scene_mod.flatten(scene)
        o, d, nrm = abi.Vec3(*origin), abi.Vec3(*direction), abi.Vec3()
        dist, idx = C.c_double(), C.c_int32()
        self._find_closest_hit(objs, len(scene), C.byref(o), C.byref(d), C.byref(dist), C.byref(nrm), C.byref(idx))
        return dist.value, nrm.tuple(), idx.value

    def trace_ray(self, scene, origin, direction, depth=10):
        objs = scene if isinstance(scene, C.Array) else scene_mod.flatten(scene)
        o, d, rgb = abi.Vec3(*origin), abi.Vec3(*direction), abi.Vec3()
        self._trace_ray(objs, len(scene), C.byref(o), C.byref(d), int(depth), C.byref(rgb))
        return rgb.tuple()

    def out_color(self, v):
        a, rgb = abi.Vec3(*v), abi.Vec3()
        self._out_color(C.byref(a), C.byref(rgb))
        return rgb.tuple()

    def reflect(self, v, n):
        a, b, out = abi.Vec3(*v), abi.Vec3(*n), abi.Vec3()
        self._reflect(C.byref(a), C.byref(b), C.byref(out))
        return out.tuple()

    def diffuse(self, pos, n, light=(0, 0, 0)):
        a, b, c = abi.Vec3(*pos), abi.Vec3(*n), abi.Vec3(*light)
        return self._diffuse(C.byref(a), C.byref(b), C.byref(c))

    def specular(self, pos, n, view, light=(0, 0, 0)):
        a, b, c, d = abi.Vec3(*pos), abi.Vec3(*n), abi.Vec3(*light), abi.Vec3(*view)
        return self._specular(C.byref(a), C.byref(b), C.byref(c), C.byref(d))

    # -- reference-only doors ----------------------------------------------------------------------
    def rt_scene(self, scene, cam):
        """The reference's own rt_scene (square frames only). Returns (radiance[H][W][3], seconds)."""
        objs = scene_mod.flatten(scene)
        d = cam.desc()
        W = int(cam.image_width)
        H = int(W / cam.aspect_ratio)
        rad = np.zeros((H, W, 3), np.float64)
        secs = self._rt_scene(objs, len(scene), C.byref(d), _ptr(rad, C.c_double))
        if secs < 0:
            raise RuntimeError("rt_scene only works for square frames (main.cpp:243)")
        return rad, secs

    def run_main(self, frames=1, capacity=640 * 640):
        """The unmodified main() headless. Returns the SDL surface words [h][w]."""
        buf = np.zeros(capacity, np.uint32)
        w, h = C.c_int32(), C.c_int32()
        rc = self._run_main(int(frames), _ptr(buf, C.c_uint32), capacity, C.byref(w), C.byref(h))
        if rc != 0:
            raise RuntimeError("ref_main returned %d" % rc)
        return buf[: w.value * h.value].reshape(h.value, w.value)


def load_port():
    return Oracle(PORT_PATH, "orc_", "port")


def load_reference():
    return Oracle(REF_PATH, "ref_", "reference")


def rgb8_bytes(rgba8):
    """RGBA8888 words -> H*W*3 bytes R,G,B (the payload SURVEY.md §8(c) hashes)."""
    a = np.asarray(rgba8, dtype=np.uint32)
    return np.stack([(a >> 24) & 0xFF, (a >> 16) & 0xFF, (a >> 8) & 0xFF], axis=-1).astype(np.uint8)
