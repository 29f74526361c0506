/*
 * rtx_b200.h — C ABI of the B200-native ray-tracing hot path.
 *
 * This is the drop-in boundary for the reference's only seam on this path:
 *
 *   void rt_scene(std::vector<vec3> u, const std::vector<std::unique_ptr<SceneGeometry>>& scene,
 *                 const Camera& cam, std::vector<std::vector<RGB>>& frame_buffer)   (main.cpp:124-139)
 *   followed by the 8-bit quantise loop                                              (main.cpp:338-347)
 *
 * The reference has no FFI; every entry point below names the reference code it
 * replaces. Plain pointers and sizes only — no C++ or torch types cross this line.
 * All reference arithmetic is IEEE double (vec.h:15), so every POD here is double.
 *
 * Conventions (mirroring main.cpp:329): calls are synchronous and blocking; the
 * caller owns and pre-allocates every output, the callee overwrites every pixel;
 * the scene is copied by rtx_set_scene (caller keeps ownership). Errors are int
 * status codes (0 = ok) plus rtx_last_error(); nothing is thrown across the ABI.
 * Calls on one context must be serialised by the caller (the reference is single
 * threaded); different contexts (GPUs) may be driven from different host threads.
 *
 * There is NO CPU fallback: every compute entry point fails with RTX_ERR_CUDA when
 * no sm_100-class device is usable.
 */
#ifndef RTX_B200_H
#define RTX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTX_ABI_VERSION 5

/* ---- status codes ------------------------------------------------------- */
#define RTX_OK              0
#define RTX_ERR_INVALID     1   /* bad argument (null pointer, negative size, H/W <= 0 ...) */
#define RTX_ERR_CUDA        2   /* CUDA runtime error or no usable device */
#define RTX_ERR_NO_SCENE    3   /* rtx_render before rtx_set_scene */
#define RTX_ERR_NOMEM       4   /* host or device allocation failed */

/* ---- PODs ---------------------------------------------------------------- */

/* vec3 / point3 / RGB (vec.h:12-40): three doubles. */
typedef struct rtx_vec3 { double x, y, z; } rtx_vec3;

/* Material (scene.h:35-49); same field order as the reference struct.
 * Reference ctor order is (color, metallic=.5, ambient=.1, diffuse=.9, specular=.4,
 * specular_exponent=50) (scene.h:48). */
typedef struct rtx_material {
    rtx_vec3 color;
    double ambient;
    double metallic;
    double diffuse;
    double specular;
    double specular_exponent;
} rtx_material;

#define RTX_SPHERE 0   /* Sphere(Material, point3 center, double radius)              scene.h:75-84  */
#define RTX_WALL   1   /* Wall(Material, point3 position, vec3 normal, length, width) scene.h:62-73  */
#define RTX_BOX    2   /* EXTENSION (no reference code: README.md:21 names a sprint-2 `Box`, the snapshot has none; parity
                          unpinned, the specification is oracle/oracle.c::box_intersect). Axis-aligned box = six Wall-like
                          faces -x,+x,-y,+y,-z,+z, each with Wall::intersect's arithmetic (scene.cpp:7-29), its minimum
                          corner, its outward normal (never flipped) and the positive unit axes as in-plane basis; nearest
                          face with t > 0, lowest face on ties; distance = t like Wall. ONE object id for the whole box. */

/* One entry of the reference's std::vector<std::unique_ptr<SceneGeometry>> (main.cpp:156-163).
 * The array index IS the object id (hit_object_index, main.cpp:80) and the tie-break order
 * (strict '<' in main.cpp:77 => lowest index wins equal distances). */
typedef struct rtx_object {
    int32_t      kind;     /* RTX_SPHERE | RTX_WALL | RTX_BOX */
    int32_t      reserved; /* must be 0 */
    rtx_material mat;
    rtx_vec3     p;        /* sphere: center   | wall: corner 'position' (scene.h:64) | box: minimum corner */
    rtx_vec3     n;        /* sphere: ignored  | wall: normal AS PASSED to the ctor; the library normalises it exactly
                              as scene.h:71 does (n / sqrt(n.n))            | box: extents (sx, sy, sz) */
    double       a;        /* sphere: radius   | wall: length                        | box: ignored */
    double       b;        /* sphere: ignored  | wall: width                         | box: ignored */
} rtx_object;

/* Inputs of Camera::init (scene.h:94-99, main.cpp:146-153). image_width/aspect_ratio are doubles
 * in the reference (scene.h:95). */
typedef struct rtx_camera_desc {
    rtx_vec3 position, lookat, vup;
    double   vfov;          /* degrees; converted with 3.14/180 (scene.cpp:84), NOT pi */
    double   aspect_ratio;
    double   image_width;
} rtx_camera_desc;

/* What Camera::init yields and rt_scene consumes (scene.cpp:80-106, main.cpp:132-134):
 * pixel (row i, col j) has centre image_top_left + delta_x*j + delta_y*i; the primary ray is
 * origin = position, direction = position - centre (unnormalised). */
typedef struct rtx_camera {
    rtx_vec3 position;
    rtx_vec3 image_top_left;
    rtx_vec3 delta_x;        /* u[0], scene.cpp:99  */
    rtx_vec3 delta_y;        /* u[1], scene.cpp:100 */
    int32_t  width;          /* columns j, image_width  */
    int32_t  height;         /* rows i,    image_height = int(image_width / aspect_ratio), scene.cpp:82 */
} rtx_camera;

/* ray (scene.h:5-25): the reference's ctor takes (direction, origin) in that order (scene.h:14); the direction is
 * NOT normalised by it and the hot path depends on its length. */
typedef struct rtx_ray {
    rtx_vec3 origin;
    rtx_vec3 direction;
} rtx_ray;

#define RTX_QUANT_WRAP     0  /* reference: implicit double->Uint8, truncation, wraps mod 256 (main.cpp:345) */
#define RTX_QUANT_SATURATE 1  /* labelled non-parity option: clamp to [0,255] */

/* Compile-time constants of the reference, exposed as run-time parameters.
 * rtx_default_params() fills in the reference literals. */
typedef struct rtx_params {
    int32_t  max_depth;        /* remaining_iterations, default 10 (main.cpp:89); <= RTX_MAX_DEPTH */
    int32_t  quantise_mode;    /* RTX_QUANT_* */
    int32_t  fuse_quantise;    /* 1: trace kernel writes rgba8 itself; 0: radiance buffer + quantise kernel */
    int32_t  accel;            /* RTX_ACCEL_* (extension, default RTX_ACCEL_NONE = the reference's brute-force loop) */
    rtx_vec3 light_pos;        /* LIGHT_POS      (0,0,0)            main.cpp:14 */
    rtx_vec3 ground_color;     /* GROUND_COLOR   (.025,.05,.075)    main.cpp:15 */
    rtx_vec3 sky_low;          /* SKYCOLOR_LOW   (.36,.45,.57)      main.cpp:16 */
    rtx_vec3 sky_high;         /* SKYCOLOR_HIGH  (.14,.21,.49)      main.cpp:17 */
    double   reflect_offset;   /* .0001                             main.cpp:111 */
    double   sky_exponent;     /* 1./4. (as float -> 0.25 exactly)  main.cpp:34 */
    /* Row sharding for one big frame split over several GPUs (one context per GPU).
     * Rows are grouped in bands of band_rows; band b belongs to rank (b mod n_ranks).
     * A rank renders only its own rows, packed in increasing row order.
     * n_ranks = 1 renders the whole frame. */
    int32_t  band_rows;
    int32_t  n_ranks;
    int32_t  rank;
    int32_t  pixel_order;      /* RTX_ORDER_*: WHEN a pixel is traced, never what it looks like; default RTX_ORDER_AUTO */
    /* Camera paths sharded by frame: camera k of this call is frame (frame_offset + k * frame_stride) of the whole
     * path. Only used to place pixels in rtx_outputs.frame_rgba8. Defaults 0, 1. */
    int32_t  frame_offset;
    int32_t  frame_stride;
    /* EXTENSIONS — off by default, i.e. the reference's semantics. README.md:13-14 speaks of a sun and of tone
     * mapping and main.cpp:18-19 defines SUN_COLOR / SUN_DIRECTION, but the snapshot contains no code that uses
     * them: parity is unpinned, the specification is the CPU restatement in oracle/oracle.c.
     * sun: every local colour (main.cpp:104) gains  color * sun_color * (max(0, s.n) * diffuse + pow(max(0, h.n),
     *   exponent) * specular)  with s = normalize(sun_direction), h = normalize(normalize(-d) + s), n the unit normal;
     *   no shadow ray (the reference's point light has none either), sky unchanged.
     * tonemap: RTX_TONEMAP_REINHARD applies Reinhard's global photographic operator per frame before the 8-bit
     *   pack: L = .2126 R + .7152 G + .0722 B, Lavg = exp(mean(log(1e-4 + L))), Ls = key / Lavg * L,
     *   Ld = Ls * (1 + Ls / white^2) / (1 + Ls) (white <= 0: Ld = Ls / (1 + Ls)), rgb *= Ld / L (0 where L <= 0).
     *   Inside rtx_render it needs the whole frame on one GPU (n_ranks must be 1); row-sharded frames use the two-step
     *   rtx_tonemap_sums / rtx_tonemap_apply with one integer all-reduce in between. */
    int32_t  sun_enabled;
    int32_t  tonemap;          /* RTX_TONEMAP_* */
    rtx_vec3 sun_color;        /* SUN_COLOR     (1.64,1.27,.99)     main.cpp:18 */
    rtx_vec3 sun_direction;    /* SUN_DIRECTION (.7,.4,.7)          main.cpp:19; towards the sun, any length > 0 */
    double   tonemap_key;      /* Reinhard's a, default .18 */
    double   tonemap_white;    /* luminance that maps to pure white; <= 0 (default): no burn-out term */
} rtx_params;

/* EXTENSION — README.md:17 names an acceleration structure as the obvious next step; the snapshot has none
 * (find_closest_hit is a plain loop over the scene vector, main.cpp:73-82). RTX_ACCEL_GRID puts a uniform grid over the
 * spheres in front of the SAME exact tests and acceptance rule: it only narrows which objects are tested, conservatively,
 * so object ids, distances, ray counts and pixels are bit for bit those of RTX_ACCEL_NONE (asserted on the full 4K frame
 * of the 10 064-object scene). The roofline / headline numbers are always quoted on RTX_ACCEL_NONE. */
#define RTX_ACCEL_NONE 0
#define RTX_ACCEL_GRID 1

/* The order in which the trace kernel's persistent lanes pick up pixels. A frame ends with the ray chains that are still
 * in flight when the last pixel has been handed out; if those last pixels are cheap ones (sky: one ray) that tail is short.
 * RTX_ORDER_COST hands out tiles of 256 pixels most expensive first, judged by the ray counts the PREVIOUS call on this
 * context collected for the same frame geometry (an interactive loop or a camera path: consecutive frames look alike; the
 * first frame of a geometry runs in scan order). Every pixel is traced exactly as before — object ids, ray counts and
 * pixels do not depend on the order. Only the brute-force kernel of scenes with more than 16 objects uses it.
 * RTX_ORDER_AUTO = RTX_ORDER_COST for frames of at least 2^18 pixels, RTX_ORDER_SCAN below. */
#define RTX_ORDER_AUTO 0
#define RTX_ORDER_SCAN 1
#define RTX_ORDER_COST 2

#define RTX_TONEMAP_NONE     0  /* reference: radiance goes straight to the 8-bit pack (main.cpp:338-347) */
#define RTX_TONEMAP_REINHARD 1  /* extension, see rtx_params */

#define RTX_MAX_DEPTH 254     /* ray_count is a uint8: depth+1 rays per pixel at most */

#define RTX_MAX_IN_FLIGHT 3

#define RTX_FRAME_STORE 0
#define RTX_FRAME_COPY  1

#define RTX_MEM_HOST        0   /* plain host pointers: device staging + copies (pinned or pageable memory) */
#define RTX_MEM_DEVICE      1   /* device pointers: outputs stay in HBM */
#define RTX_MEM_HOST_MAPPED 2   /* host pointers into PINNED, MAPPED memory (rtx_host_alloc / rtx_host_register /
                                   rtx_host_shared_open): zero copy — the kernel stores every finished pixel straight into
                                   the caller's surface over PCIe, the way the reference writes into SDL's surface->pixels
                                   (main.cpp:193,344); no staging buffer, no copy after the kernel */

/* Output planes, each optional (NULL = not wanted), caller-allocated, all of them either host or
 * device pointers (memory). Shapes are [n_frames][rows][width] where rows = height when n_ranks = 1,
 * else rtx_local_rows(height, band_rows, n_ranks, rank). */
typedef struct rtx_outputs {
    uint32_t* rgba8;         /* R<<24 | G<<16 | B<<8 | 0xFF : SDL RGBA8888 surface word (main.cpp:193,345) */
    float*    radiance_f32;  /* [..][3] frame_buffer value (main.cpp:136) rounded to float */
    double*   radiance_f64;  /* [..][3] frame_buffer value in double */
    int32_t*  object_id;     /* primary-ray hit_object_index (main.cpp:80), -1 = miss */
    uint8_t*  hit_mask;      /* 1 iff the primary ray hit anything */
    uint8_t*  ray_count;     /* rays traced for the pixel (primary + reflections), 1..max_depth+1 */
    int32_t   memory;        /* RTX_MEM_* : where the planes above live */
    int32_t   frame_mode;    /* RTX_FRAME_* : how pixels reach frame_rgba8 */
    /* Fused multi-GPU gather: a DEVICE pointer (regardless of `memory`) to a whole row-major frame set
     * [total_frames][height][width] of RGBA8888 words, typically rank 0's buffer mapped into this process with
     * rtx_buffer_import (peer memory over NVLink). The trace kernel stores every finished pixel at its GLOBAL
     * position (global row from the band map, global frame from frame_offset/frame_stride), so no all-gather and
     * no unpermute pass is needed: a barrier after the call completes the frame. NULL = not wanted.
     * The pointer may also be the device alias of a pinned HOST frame (rtx_host_device_pointer): the ranks of one box then
     * assemble the frame in shared host memory, each over its own PCIe link.
     * frame_mode = RTX_FRAME_STORE (default): the trace kernel stores each finished pixel itself (4-byte stores; right
     *   when pixels are expensive, i.e. large scenes: the stores hide entirely under the kernel).
     * frame_mode = RTX_FRAME_COPY: the call renders into context staging and then moves whole bands / whole frames to
     *   their place with copy-engine transfers (2-D copies for cyclic bands) on the context's copy stream — right when
     *   pixels are cheap (small scenes, camera paths: gigabytes per second of pixels); frame_rgba8 is then any address
     *   cudaMemcpy accepts (pinned host memory itself, local or peer device memory) and rgba8 must be NULL. With
     *   rtx_render_async the copies of one call overlap the kernel of the next. */
    uint32_t* frame_rgba8;
    /* What find_closest_hit (main.cpp:67-84) returns for the PRIMARY ray, beyond its index (object_id): */
    double*   hit_distance;  /* Collision.distance: sphere = projection * |d| (scene.cpp:77), wall = t (scene.cpp:30);
                                DBL_MAX on a miss (main.cpp:70) */
    double*   hit_normal;    /* [..][3] Collision.normal as returned (sphere: P - c, unnormalised; wall: n as stored);
                                (0,0,0) on a miss */
} rtx_outputs;

/* Device-side timing and diagnostics of the last rtx_render / rtx_quantise on a context.
 * Stage names follow the reference's log (main.cpp:386-391): "raytracing" = trace, "surface update" = quantise. */
typedef struct rtx_stats {
    double   raytracing_ms;      /* trace kernel(s), CUDA events */
    double   surface_update_ms;  /* quantise kernel (0 when fused) */
    double   h2d_ms;             /* camera upload */
    double   d2h_ms;             /* output download (RTX_MEM_HOST only) */
    double   total_ms;           /* first event to last event */
    uint64_t total_rays;         /* sum of ray_count over all pixels rendered by this call */
    uint64_t sphere_tests;       /* total_rays * n_spheres */
    uint64_t wall_tests;         /* total_rays * (n_walls + 6 per RTX_BOX: a box is six wall-like faces) */
    uint64_t over_range_pixels;  /* pixels with a channel outside [0, 256/255): quantise wraps there */
    double   max_luminance;      /* max over pixels of (R+G+B)/3; diagnostic only, never alters pixels */
    int32_t  launches;           /* kernels launched by the call */
    int32_t  reserved;
    double   drain_ms;           /* trace kernel: from the pixel pool running empty to the last warp's exit */
    double   exit_spread_ms;     /* trace kernel: first warp exit to last warp exit */
} rtx_stats;

typedef struct rtx_ctx rtx_ctx;

/* ---- entry points ---------------------------------------------------------- */

/* ABI version of the loaded library (== RTX_ABI_VERSION). */
int rtx_abi_version(void);

/* Human-readable text for a status code. */
const char* rtx_status_string(int status);

/* One context per GPU: owns the stream, events, the device copy of the scene and scratch buffers.
 * Replaces nothing in the reference (it has no device); plays the role of the objects that live
 * for the duration of main() (main.cpp:146-173). */
int  rtx_create(rtx_ctx** out, int device);
void rtx_destroy(rtx_ctx* ctx);

/* Text of the last error on this context ("" if none). Valid until the next call on ctx. */
const char* rtx_last_error(const rtx_ctx* ctx);

/* Use an externally owned cudaStream_t (passed as void*) for all work of this context, e.g. the
 * caller's current stream so that its own events bracket the kernels. NULL is CUDA's default stream (a valid
 * stream handle); RTX_STREAM_PRIVATE restores the context's own non-blocking stream. */
#define RTX_STREAM_PRIVATE ((void*)(intptr_t)-1)
int rtx_set_stream(rtx_ctx* ctx, void* cuda_stream);

/* Replaces the scene.push_back(...) sequence (main.cpp:156-163): copies n objects in scene order,
 * normalises wall normals (scene.h:71), precomputes each wall's in-plane basis (scene.cpp:18-19, which
 * the reference recomputes per intersection although it is ray independent) and uploads SoA arrays. */
int rtx_set_scene(rtx_ctx* ctx, const rtx_object* objects, int32_t n_objects);

/* Camera::init (scene.cpp:80-106) on the host, in double, quirks included (3.14, int() truncation). */
int rtx_camera_init(const rtx_camera_desc* desc, rtx_camera* out);

/* Fills params with the reference's literals (see rtx_params). */
void rtx_default_params(rtx_params* params);

/* Number of rows rank 'rank' renders of a frame of 'height' rows with cyclic bands. */
int32_t rtx_local_rows(int32_t height, int32_t band_rows, int32_t n_ranks, int32_t rank);

/* Global row index of this rank's packed local row (inverse of the band map); -1 if out of range. */
int32_t rtx_global_row(int32_t local_row, int32_t height, int32_t band_rows, int32_t n_ranks, int32_t rank);

/* Replaces rt_scene (main.cpp:124-139) + the quantise loop (main.cpp:338-347) for n_frames cameras
 * that share width/height: ray generation, nearest hit over all objects (main.cpp:67-84), the
 * reflection chain with the depth cap (main.cpp:89-119), Blinn-Phong (main.cpp:42-62,102-104),
 * sky (main.cpp:28-37) and the truncating 8-bit pack. stats may be NULL. */
int rtx_render(rtx_ctx* ctx, const rtx_camera* cameras, int32_t n_frames,
               const rtx_params* params, const rtx_outputs* outputs, rtx_stats* stats);

/* The same call without the wait: everything (camera upload, kernels, read-back) is queued and the call returns.
 * rtx_wait blocks until the OLDEST call in flight on this context has completed and yields its stats; until then
 * the caller must not touch that call's outputs. At most RTX_MAX_IN_FLIGHT calls may be in flight per context (each
 * owns its own staging buffers): with host outputs, frame k crosses PCIe on the context's copy stream while the kernel
 * of frame k+1 runs and the host queues frame k+2 — the per-call latency of small frames (host launch path, kernel,
 * read-back) overlaps instead of adding up; a further call before rtx_wait is refused with RTX_ERR_INVALID.
 * A synchronous rtx_render first completes the calls in flight (their stats are dropped). */
int rtx_render_async(rtx_ctx* ctx, const rtx_camera* cameras, int32_t n_frames,
                     const rtx_params* params, const rtx_outputs* outputs);
int rtx_wait(rtx_ctx* ctx, rtx_stats* stats);

/* Replaces recursive_ray_tracing(scene, ray, remaining_iterations) (main.cpp:89-119, called per pixel at main.cpp:136)
 * and find_closest_hit(scene, ray) (main.cpp:67-84) for a caller-supplied batch of rays: the same kernels as
 * rtx_render, with ray k taking the place of pixel k's primary ray. Output planes are shaped [n_rays] (radiance
 * [n_rays][3]); object_id / hit_distance / hit_normal describe the nearest hit of the ray itself, radiance is the
 * value recursive_ray_tracing returns for it with params->max_depth. rays is a HOST pointer; band sharding,
 * frame_rgba8 and the tone-map extension do not apply. */
int rtx_trace_rays(rtx_ctx* ctx, const rtx_ray* rays, int64_t n_rays,
                   const rtx_params* params, const rtx_outputs* outputs, rtx_stats* stats);

/* The quantise loop alone (main.cpp:338-347) on a caller-supplied radiance buffer of n_pixels
 * RGB triples (exactly one of radiance_f32 / radiance_f64 non-NULL), memory = RTX_MEM_*. */
int rtx_quantise(rtx_ctx* ctx, const float* radiance_f32, const double* radiance_f64, int64_t n_pixels,
                 int32_t quantise_mode, uint32_t* rgba8, int32_t memory, rtx_stats* stats);

/* EXTENSION (rtx_params.tonemap; no reference code, specification = oracle/oracle.c::orc_tonemap): Reinhard's global
 * operator per frame followed by the 8-bit pack (params->quantise_mode), on a caller-supplied radiance buffer of n_frames
 * frames of pixels_per_frame RGB triples (exactly one of radiance_f32 / radiance_f64 non-NULL), memory = RTX_MEM_*.
 * params->tonemap must be RTX_TONEMAP_REINHARD. log_avg_luminance (HOST pointer, may be NULL) receives the n_frames
 * log-average luminances the operator used. The statistic is accumulated in fixed point with integer atomics, so the
 * output is identical from run to run. A float buffer is processed in single precision (luminance, logf, the map; the
 * per-frame constants and the 8-bit pack stay double) — the specification's rgb32 branch — a double buffer in double. */
int rtx_tonemap(rtx_ctx* ctx, const float* radiance_f32, const double* radiance_f64, int64_t pixels_per_frame, int32_t n_frames,
                const rtx_params* params, uint32_t* rgba8, int32_t memory, double* log_avg_luminance, rtx_stats* stats);

/* The same operator in two steps, for frames whose rows are spread over several GPUs (DEVICE pointers only):
 *   1. every rank: rtx_tonemap_sums ADDS the fixed-point log-luminance sums of its own pixels to sums[n_frames]
 *      (int64, zeroed by the caller beforehand);
 *   2. the caller all-reduces sums over the ranks with an integer SUM (NCCL over NVLink; sharding.py::tonemap_sharded) —
 *      integer addition is associative, so the result does not depend on the number of ranks or the reduction order;
 *   3. every rank: rtx_tonemap_apply maps and packs its own pixels with the global sums and the GLOBAL number of
 *      pixels per frame. The assembled frame is bit for bit the one rtx_tonemap produces on a single GPU. */
int rtx_tonemap_sums(rtx_ctx* ctx, const float* radiance_f32, const double* radiance_f64, int64_t pixels_per_frame, int32_t n_frames,
                     int64_t* sums);
int rtx_tonemap_apply(rtx_ctx* ctx, const float* radiance_f32, const double* radiance_f64, int64_t pixels_per_frame, int32_t n_frames,
                      const int64_t* sums, int64_t pixels_per_frame_global, const rtx_params* params, uint32_t* rgba8, rtx_stats* stats);

/* Multi-GPU gather epilogue: scatters a band-major buffer (n_ranks blocks of rows_per_rank rows, block r =
 * rank r's packed rows, as an all-gather delivers them) into a row-major frame. Device pointers,
 * elem_bytes in {1,4}. rows_per_rank must be >= every rank's rtx_local_rows(). */
int rtx_unpermute_bands(rtx_ctx* ctx, const void* band_major, void* row_major, int32_t height, int32_t width,
                        int32_t elem_bytes, int32_t band_rows, int32_t n_ranks, int32_t rows_per_rank);

/* Device buffers that can be shared between the per-GPU processes of one box (CUDA IPC): rank 0 allocates the
 * frame and exports a 64-byte handle, the other ranks import it and pass the mapped pointer as
 * rtx_outputs.frame_rgba8. Import enables peer access (NVLink) lazily. */
#define RTX_IPC_HANDLE_BYTES 64
int rtx_buffer_alloc(rtx_ctx* ctx, uint64_t bytes, void** device_ptr);
int rtx_buffer_free(rtx_ctx* ctx, void* device_ptr);
int rtx_buffer_export(rtx_ctx* ctx, void* device_ptr, uint8_t handle[RTX_IPC_HANDLE_BYTES]);
int rtx_buffer_import(rtx_ctx* ctx, const uint8_t handle[RTX_IPC_HANDLE_BYTES], void** device_ptr);
int rtx_buffer_release(rtx_ctx* ctx, void* imported_ptr);
/* Copies `bytes` from device memory (rtx_buffer_alloc / _import, or any device pointer of this context's GPU or a peer)
 * into host memory, ordered after the context's work; blocking. The host-side end of a frame assembled in rank 0's HBM
 * (rtx::ShardedRenderer::render_to_device, the `value` path of bench.py), e.g. to hand it to a presenter or a file. */
int rtx_buffer_read(rtx_ctx* ctx, const void* device_ptr, void* host_ptr, uint64_t bytes);

/* ONE process driving several GPUs (one context + one host thread per GPU, the C++ host's rtx::ShardedRenderer): the
 * kernels of `ctx`'s device may then store into memory of `peer_device` (rank 0's frame from rtx_buffer_alloc, passed to
 * the other contexts as rtx_outputs.frame_rgba8) over NVLink — the in-process counterpart of rtx_buffer_export / _import.
 * Idempotent. rtx_device_count: number of visible CUDA devices (0 if none / no driver). */
int rtx_enable_peer_access(rtx_ctx* ctx, int peer_device);
int rtx_device_count(void);

/* Pinned, mapped host memory for RTX_MEM_HOST_MAPPED outputs and for fast RTX_MEM_HOST read-back — the role of the
 * SDL surface the reference quantises into (SDL_CreateRGBSurface, main.cpp:193; written main.cpp:338-347).
 *   rtx_host_alloc / rtx_host_free         a new pinned + mapped buffer;
 *   rtx_host_register / rtx_host_unregister  pin + map memory the caller already owns (e.g. a presenter's surface);
 *   rtx_host_device_pointer                the device-side alias of such memory (what a kernel dereferences), e.g.
 *                                          to pass a HOST frame as rtx_outputs.frame_rgba8;
 *   rtx_host_shared_open / _close          a POSIX shared-memory object ("/name") mapped and pinned in THIS process:
 *                                          the per-GPU processes of one box all map the same host frame and each
 *                                          writes its own rows over its own PCIe link (create != 0 in exactly one of
 *                                          them, before the others open it; unlink_name != NULL removes the object). */
int rtx_host_alloc(rtx_ctx* ctx, uint64_t bytes, void** host_ptr);
int rtx_host_free(rtx_ctx* ctx, void* host_ptr);
int rtx_host_register(rtx_ctx* ctx, void* host_ptr, uint64_t bytes);
int rtx_host_unregister(rtx_ctx* ctx, void* host_ptr);
int rtx_host_device_pointer(rtx_ctx* ctx, void* host_ptr, void** device_ptr);
int rtx_host_shared_open(rtx_ctx* ctx, const char* name, uint64_t bytes, int32_t create, void** host_ptr);
int rtx_host_shared_close(rtx_ctx* ctx, void* host_ptr, const char* unlink_name);

/* FP32 FFMA throughput microbenchmark (the roofline denominator has no entry in MEASURED_PEAKS.json):
 * returns achieved TFLOP/s of a dependent-chain-free FFMA loop over the whole chip. variant 0 = scalar
 * FFMA, 1 = packed fma.rn.f32x2 with two operands shared by all instructions, 2..4 = packed with three /
 * two-plus-one-shared / two distinct register pairs per instruction (register-file bandwidth probes), 5..6 =
 * packed and scalar FMAs interleaved (is there a second FP32 pipe? no: 66-68 vs 74 TFLOP/s), 7 = packed HALF FMAs
 * (fma.rn.f16x2; reported in half-precision TFLOP/s; not used by any kernel, a data point only). */
int rtx_ffma_peak(rtx_ctx* ctx, int32_t variant, double* tflops, double* sm_clock_mhz_estimate);

#ifdef __cplusplus
}
#endif
#endif /* RTX_B200_H */
