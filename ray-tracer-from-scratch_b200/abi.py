"""ctypes mirror of include/rtx_b200.h (the C ABI). Pure declarations: no compute, no library loading.

Each structure matches the POD of the same name in the header field for field; the header cites
the reference code (file:line) every field stands for.
"""
import ctypes as C

ABI_VERSION = 5

RTX_OK, RTX_ERR_INVALID, RTX_ERR_CUDA, RTX_ERR_NO_SCENE, RTX_ERR_NOMEM = 0, 1, 2, 3, 4
RTX_SPHERE, RTX_WALL, RTX_BOX = 0, 1, 2        # RTX_BOX: extension, see the header
RTX_QUANT_WRAP, RTX_QUANT_SATURATE = 0, 1
RTX_ACCEL_NONE, RTX_ACCEL_GRID = 0, 1               # extension, see the header
RTX_ORDER_AUTO, RTX_ORDER_SCAN, RTX_ORDER_COST = 0, 1, 2   # scheduling hint of the big kernel, see the header
RTX_TONEMAP_NONE, RTX_TONEMAP_REINHARD = 0, 1    # extension, see the header
RTX_MEM_HOST, RTX_MEM_DEVICE, RTX_MEM_HOST_MAPPED = 0, 1, 2
RTX_FRAME_STORE, RTX_FRAME_COPY = 0, 1
RTX_MAX_DEPTH = 254
RTX_MAX_IN_FLIGHT = 3


class Vec3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]

    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(float(x), float(y), float(z))

    def tuple(self):
        return (self.x, self.y, self.z)


class MaterialPOD(C.Structure):
    _fields_ = [("color", Vec3), ("ambient", C.c_double), ("metallic", C.c_double), ("diffuse", C.c_double),
                ("specular", C.c_double), ("specular_exponent", C.c_double)]


class ObjectPOD(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("mat", MaterialPOD), ("p", Vec3), ("n", Vec3),
                ("a", C.c_double), ("b", C.c_double)]


class CameraDesc(C.Structure):
    _fields_ = [("position", Vec3), ("lookat", Vec3), ("vup", Vec3), ("vfov", C.c_double),
                ("aspect_ratio", C.c_double), ("image_width", C.c_double)]


class CameraPOD(C.Structure):
    _fields_ = [("position", Vec3), ("image_top_left", Vec3), ("delta_x", Vec3), ("delta_y", Vec3),
                ("width", C.c_int32), ("height", C.c_int32)]


class RayPOD(C.Structure):
    _fields_ = [("origin", Vec3), ("direction", Vec3)]


class Params(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("quantise_mode", C.c_int32), ("fuse_quantise", C.c_int32),
                ("accel", C.c_int32),
                ("light_pos", Vec3), ("ground_color", Vec3), ("sky_low", Vec3), ("sky_high", Vec3),
                ("reflect_offset", C.c_double), ("sky_exponent", C.c_double),
                ("band_rows", C.c_int32), ("n_ranks", C.c_int32), ("rank", C.c_int32), ("pixel_order", C.c_int32),
                ("frame_offset", C.c_int32), ("frame_stride", C.c_int32),
                # extensions (off by default): sun and tone-map operator
                ("sun_enabled", C.c_int32), ("tonemap", C.c_int32), ("sun_color", Vec3), ("sun_direction", Vec3),
                ("tonemap_key", C.c_double), ("tonemap_white", C.c_double)]


class Outputs(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("radiance_f32", C.c_void_p), ("radiance_f64", C.c_void_p),
                ("object_id", C.c_void_p), ("hit_mask", C.c_void_p), ("ray_count", C.c_void_p),
                ("memory", C.c_int32), ("frame_mode", C.c_int32), ("frame_rgba8", C.c_void_p),
                ("hit_distance", C.c_void_p), ("hit_normal", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("raytracing_ms", C.c_double), ("surface_update_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("total_rays", C.c_uint64), ("sphere_tests", C.c_uint64), ("wall_tests", C.c_uint64),
                ("over_range_pixels", C.c_uint64), ("max_luminance", C.c_double),
                ("launches", C.c_int32), ("reserved", C.c_int32),
                ("drain_ms", C.c_double), ("exit_spread_ms", C.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_ if not name.startswith("reserved")}


# Every symbol include/rtx_b200.h declares (checked by the CPU test-suite against the built .so).
EXPORTS = (
    "rtx_abi_version", "rtx_status_string", "rtx_create", "rtx_destroy", "rtx_last_error", "rtx_set_stream",
    "rtx_set_scene", "rtx_camera_init", "rtx_default_params", "rtx_local_rows", "rtx_global_row",
    "rtx_render", "rtx_render_async", "rtx_wait", "rtx_trace_rays", "rtx_quantise", "rtx_tonemap", "rtx_tonemap_sums", "rtx_tonemap_apply", "rtx_unpermute_bands", "rtx_ffma_peak",
    "rtx_enable_peer_access", "rtx_device_count", "rtx_host_alloc", "rtx_host_free", "rtx_host_register", "rtx_host_unregister", "rtx_host_device_pointer",
    "rtx_host_shared_open", "rtx_host_shared_close",
    "rtx_buffer_alloc", "rtx_buffer_free", "rtx_buffer_export", "rtx_buffer_import", "rtx_buffer_release", "rtx_buffer_read",
)
