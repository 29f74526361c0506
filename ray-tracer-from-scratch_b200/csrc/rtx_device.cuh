// rtx_device.cuh — device-side data layout and the exact-double vocabulary of the trace kernel.
//
// The reference computes everything in IEEE double without FMA contraction (x86-64 baseline build,
// CMakeLists.txt:23). Every helper in namespace ex:: therefore uses the explicitly rounded intrinsics
// (__dadd_rn, __dmul_rn, __ddiv_rn, __dsqrt_rn), which nvcc never fuses into DFMA, and keeps the
// reference's operation order (cited per function). That is what makes object ids and distances
// bit-identical to the reference, not merely "close".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rtx_b200.h"

namespace rtx {

struct d3 {
    double x, y, z;
};

namespace ex {
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }

__device__ __forceinline__ d3 mk(double x, double y, double z) { return d3{x, y, z}; }
__device__ __forceinline__ d3 mk(const rtx_vec3& v) { return d3{v.x, v.y, v.z}; }
// vec3::dot, vec.cpp:11-14: (x*x' + y*y') + z*z'
__device__ __forceinline__ double dot(d3 a, d3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
// vec3::length_squared / length, vec.cpp:3-9
__device__ __forceinline__ double len2(d3 a) { return add(add(mul(a.x, a.x), mul(a.y, a.y)), mul(a.z, a.z)); }
__device__ __forceinline__ double len(d3 a) { return sqrt(len2(a)); }
__device__ __forceinline__ d3 add(d3 a, d3 b) { return d3{add(a.x, b.x), add(a.y, b.y), add(a.z, b.z)}; }   // vec.cpp:26-28
__device__ __forceinline__ d3 sub(d3 a, d3 b) { return d3{sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)}; }   // vec.cpp:32-34
__device__ __forceinline__ d3 neg(d3 a) { return d3{-a.x, -a.y, -a.z}; }                                    // vec.cpp:29-31
__device__ __forceinline__ d3 scale(d3 a, double s) { return d3{mul(a.x, s), mul(a.y, s), mul(a.z, s)}; }   // vec.cpp:38-40
__device__ __forceinline__ d3 divs(d3 a, double s) { return d3{div(a.x, s), div(a.y, s), div(a.z, s)}; }    // vec.cpp:41-43
// vec3::normalize, vec.cpp:21-24: three divides by length(), not a multiply by the reciprocal
__device__ __forceinline__ d3 unit(d3 a) { return divs(a, len(a)); }
// vec3::linear_interp, vec.cpp:45-49: a + d*(b - a)
__device__ __forceinline__ d3 lerp(d3 a, d3 b, double t)
{
    return d3{add(a.x, mul(t, sub(b.x, a.x))), add(a.y, mul(t, sub(b.y, a.y))), add(a.z, mul(t, sub(b.z, a.z)))};
}
}  // namespace ex

// ---- device copy of the scene (built by rtx_set_scene) --------------------------------------------
//
// The virtual SceneGeometry::intersect (scene.h:51-60) is replaced by one flat array of screen entries
// (warp-uniform scan, no dispatch) plus type-switched exact data for the few survivors. Every exact record carries
// a KEY = scene id * 8 + face: the scene id is the index in the reference's scene vector, which is both the object
// id and the tie-break order (main.cpp:77-80); the low three bits order the six faces of an RTX_BOX (extension; 0 for
// spheres and walls), so comparing keys reproduces the reference's order and key >> 3 is the object id.

struct SphereExact {   // 32 B, read only for filter survivors
    double cx, cy, cz, r;
};

struct WallDev {       // everything Wall::intersect needs, with the ray-independent basis precomputed
    d3 p;              // corner (scene.h:64)
    d3 n;              // normal after the ctor's normalize (scene.h:71)
    d3 right;          // normalize(cross(n, (0,0,1)))        scene.cpp:18
    d3 up;             // normalize(cross(right, n))          scene.cpp:19
    double length, width;
    int32_t key;       // tie-break key = scene id * 8 + face (face = 0 for a Wall, 0..5 for the faces of a box)
    int32_t pad;
};

struct MaterialDev {   // Material (scene.h:35-49), indexed by scene id
    d3 color;
    double ambient, metallic, diffuse, specular, exponent;
};

struct SceneDev {
    int32_t n_objects, n_spheres, n_walls;
    int32_t n_entries;                 // n_spheres + n_walls: what the FP32 screen scans
    int32_t n_entries_padded;          // multiple of the hot loop's entries per iteration
    // Screen entries (cx, cy, cz, r) rounded to float: spheres first (entry e = sphere slot e), then one
    // BOUNDING sphere per wall (entry n_spheres + w = wall slot w); padding has r = -1.
    const float4* ent32;               // [n_entries_padded]
    const SphereExact* sph64;          // [n_spheres]
    const int32_t* sph_key;            // [n_spheres] scene id * 8
    const WallDev* walls;              // [n_walls] walls and box faces (six consecutive entries per box)
    const MaterialDev* mats;           // [n_objects]
    const int32_t* kind;               // [n_objects] RTX_SPHERE | RTX_WALL | RTX_BOX
    const int32_t* slot;               // [n_objects] index into sph64 / walls (box: its first face)
};

// EXTENSION (rtx_params.accel = RTX_ACCEL_GRID; SURVEY §8(f)4, README.md:17 "acceleration structure"): a uniform grid over
// the finite spheres. Cell (ix, iy, iz) lists the sphere slots whose inflated bounding box overlaps it; walls and
// spheres the grid cannot hold (non-finite, or so large that they would fill it) stay on the "always" list and are
// screened for every ray. The grid only PROPOSES candidates: every hit decision is still taken by the exact double
// tests, with the reference's acceptance rule, so ids / distances / pixels equal the brute-force ones bit for bit.
struct GridDev {
    int32_t nx, ny, nz;                // 0, 0, 0: no grid (no finite sphere)
    int32_t n_items;
    double x0, y0, z0;                 // minimum corner
    double x1, y1, z1;                 // maximum corner
    double cell, inv_cell;             // cubic cells
    double slack;                      // distance margin of the early exit (>> DDA rounding, << cell)
    const uint32_t* cell_start;        // [nx*ny*nz + 1]
    const void* items;                 // sphere slots, cell after cell: uint16_t if items16 else uint32_t
    int32_t items16;
    int32_t n_always;
    const int32_t* always;             // [n_always] sphere slots screened for every ray
};

// Per-launch arguments of the trace kernel.
struct TraceArgs {
    SceneDev scene;
    const rtx_camera* cameras;         // [n_frames] device copy
    int32_t use_grid;                  // 1: trace_grid_kernel with `grid` (extension); 0: brute force
    GridDev grid;
    const rtx_ray* rays;               // rtx_trace_rays: ray p replaces pixel p's primary ray (n_frames = local_rows = 1, width = n_rays)
    int32_t n_frames, width, height;   // global frame size
    int32_t local_rows;                // rows this rank renders per frame
    int32_t band_rows, n_ranks, rank;
    int32_t max_depth, quantise_mode;
    d3 light, ground, sky_low, sky_high;
    double reflect_offset, sky_exponent;
    int32_t sun_enabled;               // extension (rtx_params.sun_enabled): directional Blinn-Phong term
    d3 sun_dir, sun_color;             // unit vector towards the sun; its colour
    float filter_eps;                  // E: bound on the FP32 filter's distance error (see trace.cu)
    float origin_bound;                // rays whose origin exceeds this fall back to exact tests
    // outputs (device pointers, any may be null)
    uint32_t* rgba8;
    float* rad32;
    double* rad64;
    int32_t* object_id;
    uint8_t* hit_mask;
    uint8_t* ray_count;
    uint32_t* frame_rgba8;             // whole row-major frame set, possibly peer or mapped host memory (fused gather); may be null
    double* hit_distance;              // primary ray: Collision.distance (DBL_MAX on a miss, main.cpp:70)
    double* hit_normal;                // primary ray: Collision.normal as returned, (0,0,0) on a miss
    int32_t frame_offset, frame_stride;
    // One 16-byte slot, two readings (keeping sizeof(TraceArgs) where it was: the kernel copies the argument block to its
    // stack, and 16 more bytes there reshuffled ptxas' choices in trace_kernel's hot loop, +0.3 % frame time):
    union {
        // trace_small_kernel — pixel range of this launch in the packed pixel space [n_frames][local_rows][width];
        // pixel_end = 0 means all. rtx_render launches a small scene as a few consecutive ranges so that the read-back of one
        // range overlaps the tracing of the next (counters[0] is preset to pixel_begin by the host).
        struct {
            unsigned long long pixel_begin, pixel_end;
        };
        // trace_kernel — scheduling hint; WHICH pixel a pool slot stands for, no result depends on it:
        // slot s is pixel tile_order[s >> 8] * 256 + (s & 255) (tiles of 256 consecutive packed pixels, most expensive
        // first), and tile_cost[t] collects the rays traced for tile t's pixels — the next frame's order is built from it.
        struct {
            const uint32_t* tile_order;    // may be null: slot s = pixel s
            uint32_t* tile_cost;           // may be null
        };
    };
    // counters: [0] next pixel, [1] total rays, [2] over-range pixels, [3] max luminance (double bits),
    // [4] kernel start, [5] pixel pool empty, [6] first warp exit, [7] last warp exit (globaltimer ns; [4..6] start at ~0)
    unsigned long long* counters;
    void* tail_scratch;                // trace_kernel: n_sms * 1024 chain records of 104 B for the tail rebalance (may be null: off)
};

constexpr int kOrderTileShift = 8;     // 256 pixels per tile
constexpr int kOrderClasses = 128;     // cost classes of a tile: cost >> 5
constexpr int kOrderKeys = 4 * kOrderClasses;   // counting sort of the tiles by (cost class, how expensive the surroundings are)

constexpr size_t kTailScratchBytesPerCta = 1024 * 104;

// Reference quantise, main.cpp:345: implicit double -> Uint8. g++ emits cvttsd2si (32-bit) and keeps the
// low byte: truncation toward zero, wrap mod 256; NaN or |v| >= 2^31 give 0x80000000 -> byte 0.
// CUDA's own conversion saturates instead, so the out-of-range cases are spelled out.
__device__ __forceinline__ uint32_t to_u8_wrap(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return 0u;
    return static_cast<uint32_t>(__double2int_rz(v)) & 0xFFu;
}
__device__ __forceinline__ uint32_t to_u8_sat(double v)
{
    if (!(v > 0.0)) return 0u;
    if (v >= 255.0) return 255u;
    return static_cast<uint32_t>(__double2int_rz(v));
}
// SDL_MapRGB for the RGBA8888 masks of main.cpp:193.
__device__ __forceinline__ uint32_t pack_rgba(double r, double g, double b, int mode)
{
    const double R = ex::mul(r, 255.0), G = ex::mul(g, 255.0), B = ex::mul(b, 255.0);
    uint32_t ur, ug, ub;
    if (mode == RTX_QUANT_SATURATE) {
        ur = to_u8_sat(R); ug = to_u8_sat(G); ub = to_u8_sat(B);
    } else {
        ur = to_u8_wrap(R); ug = to_u8_wrap(G); ub = to_u8_wrap(B);
    }
    return (ur << 24) | (ug << 16) | (ub << 8) | 0xFFu;
}
// Pack + "over range" flag from ONE set of products (the streaming quantise kernel is bound by its double
// arithmetic for float input, so the products are not formed twice).
__device__ __forceinline__ uint32_t pack_rgba_flag(double r, double g, double b, int mode, bool& over)
{
    const double R = ex::mul(r, 255.0), G = ex::mul(g, 255.0), B = ex::mul(b, 255.0);
    over = !(R >= 0.0 && R < 256.0 && G >= 0.0 && G < 256.0 && B >= 0.0 && B < 256.0);
    uint32_t ur, ug, ub;
    if (mode == RTX_QUANT_SATURATE) {
        ur = to_u8_sat(R); ug = to_u8_sat(G); ub = to_u8_sat(B);
    } else if (!over) {   // in range: plain truncation, no wrap or NaN handling needed
        ur = static_cast<uint32_t>(__double2int_rz(R)); ug = static_cast<uint32_t>(__double2int_rz(G)); ub = static_cast<uint32_t>(__double2int_rz(B));
    } else {
        ur = to_u8_wrap(R); ug = to_u8_wrap(G); ub = to_u8_wrap(B);
    }
    return (ur << 24) | (ug << 16) | (ub << 8) | 0xFFu;
}
// A channel is "over range" when the reference's wrap would alter it: v*255 outside [0,256).
__device__ __forceinline__ bool over_range(double r, double g, double b)
{
    const double R = ex::mul(r, 255.0), G = ex::mul(g, 255.0), B = ex::mul(b, 255.0);
    return !(R >= 0.0 && R < 256.0 && G >= 0.0 && G < 256.0 && B >= 0.0 && B < 256.0);
}

constexpr int kSmallSceneEntries = 16;   // scenes of at most this many screen entries run trace_small_kernel

// Per-context (= per-device) launch state of the trace kernels: cudaFuncAttributeMaxDynamicSharedMemorySize is a
// per-device attribute, so what has been raised is remembered per context, not per thread or per process.
struct TraceLaunchState {
    bool smem_opt_in = false;          // trace_kernel<*>: the opt-in to > 48 KB of dynamic shared memory has been made by this context
    int small_per_sm = 0;              // resident CTAs per SM of trace_small_kernel (occupancy query, once)
    bool smem_opt_in_grid = false;     // trace_grid_kernel (extension)
    int grid_per_sm = 0;
};

// Launch wrappers implemented in the .cu files, called by api.cu.
cudaError_t launch_trace(const TraceArgs& args, int n_sms, cudaStream_t stream, int* launches, TraceLaunchState* state);
cudaError_t launch_trace_grid(const TraceArgs& args, int n_sms, cudaStream_t stream, int* launches, TraceLaunchState* state);
cudaError_t launch_quantise_f64(const double* rad, int64_t n_pixels, int mode, uint32_t* rgba8,
                                unsigned long long* counters, int n_sms, cudaStream_t stream);
cudaError_t launch_quantise_f32(const float* rad, int64_t n_pixels, int mode, uint32_t* rgba8,
                                unsigned long long* counters, int n_sms, cudaStream_t stream);
cudaError_t launch_tonemap_sums(const float* rad32, const double* rad64, int64_t pixels_per_frame, int n_frames, long long* sums,
                                int n_sms, cudaStream_t stream);
cudaError_t launch_tonemap_apply(const float* rad32, const double* rad64, int64_t pixels_per_frame, int n_frames, const long long* sums,
                                 int64_t pixels_global, double key, double white, int mode, uint32_t* rgba8, unsigned long long* counters,
                                 int n_sms, cudaStream_t stream);
cudaError_t launch_small_upload(void* dst, const void* src_host_mapped, size_t bytes, cudaStream_t stream);
size_t tile_order_cells(int width, long long packed_rows);
cudaError_t launch_tile_reset(uint32_t* tile_order, uint32_t* tile_cost, int n_tiles_all, uint32_t* hist_fill_cells, int n_aux, cudaStream_t stream);
cudaError_t launch_tile_order(uint32_t* tile_cost, int n_tiles, int width, long long packed_rows, uint32_t* cells, uint16_t* tile_key,
                              uint32_t* hist_now, uint32_t* hist_next, uint32_t* fill, uint32_t* tile_order, cudaStream_t stream);
cudaError_t launch_reset_counters(unsigned long long* counters, unsigned long long first_pixel, bool all, cudaStream_t stream);
cudaError_t launch_unpermute(const void* band_major, void* row_major, int height, int width, int elem_bytes,
                             int band_rows, int n_ranks, int rows_per_rank, int n_sms, cudaStream_t stream);
cudaError_t run_ffma_peak(int variant, int n_sms, cudaStream_t stream, double* tflops, double* mhz);

}  // namespace rtx
