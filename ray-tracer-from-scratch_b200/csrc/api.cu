// api.cu — the extern "C" layer of include/rtx_b200.h: context, scene upload, render, quantise, gather epilogue.
// Host code only (kernels are in trace.cu / aux_kernels.cu). Built with -ffp-contract=off: the few doubles
// computed here (wall normal + basis, Camera::init) must round exactly like the reference's x86-64 build.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "rtx_device.cuh"

using namespace rtx;

constexpr int kRanges = 4;                        // pixel ranges of a small-scene frame rendered into host memory
constexpr size_t kOrderMinPixels = 1u << 18;      // RTX_ORDER_AUTO: frames below this keep the scan order (the sort is two more launches)
constexpr size_t kRangedMinPixels = 1u << 17;     // below this a frame is one launch and one copy
constexpr int kPlanes = 8;                        // rgba8, radiance f32, radiance f64, object id, hit mask, ray count, hit distance, hit normal
constexpr size_t kSmallUploadBytes = 64 << 10;    // uploads up to this size go through a kernel, not a copy engine (aux_kernels.cu)
constexpr int kSlots = RTX_MAX_IN_FLIGHT;         // calls in flight per context (rtx_render_async)

// Everything ONE call in flight owns: pinned + device copies of its cameras / rays, its counters, its device staging
// for host outputs and its events. Two slots let frame k's read-back overlap frame k+1's kernel (rtx_render_async).
struct Slot {
    bool pending = false;
    rtx_camera* d_cameras = nullptr;
    rtx_camera* h_cameras = nullptr;   // pinned
    int cameras_cap = 0;
    rtx_ray* d_rays = nullptr;
    rtx_ray* h_rays = nullptr;         // pinned
    size_t rays_cap = 0;
    unsigned long long* d_counters = nullptr;
    unsigned long long* h_counters = nullptr;   // pinned: [0..7] read-back, [8..15] reset template, [16..23] first pixel of each range
    void* d_out[kPlanes] = {};
    size_t d_out_cap[kPlanes] = {};
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_range[kRanges + 1] = {};   // range k traced (0..kRanges-1); all copies done (kRanges)
    cudaEvent_t ev_done = nullptr;            // everything of the call (kernels, copies, counter read-back) is complete
    // what rtx_wait needs to fill rtx_stats
    int launches = 0;
    int n_spheres = 0, n_walls = 0;
    bool copies_on_side_stream = false;
};

struct rtx_ctx {
    int device = 0;
    int n_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;           // read-back of finished frames / pixel ranges while the next is traced
    std::string error;
    TraceLaunchState launch_state;

    // scene
    bool have_scene = false;
    SceneDev scene = {};
    void* d_scene_blob = nullptr;
    size_t scene_blob_cap = 0;
    unsigned char* h_scene_blob = nullptr;        // pinned staging of the blob: the upload is stream-ordered, no synchronise
    size_t h_scene_blob_cap = 0;
    cudaEvent_t ev_scene = nullptr;               // the last upload has left h_scene_blob
    double scene_bound = 0.0;          // max over objects of |coordinate| + extent
    // uniform grid (extension, rtx_params.accel): built lazily from the host copy of the spheres
    std::vector<SphereExact> h_sph64;
    bool grid_valid = false;
    GridDev grid = {};
    void* d_grid_blob = nullptr;
    size_t grid_blob_cap = 0;
    void* d_tail_scratch = nullptr;    // trace_kernel's tail rebalance: chain records of every CTA (allocated with the first big scene)
    size_t tail_scratch_cap = 0;
    // scheduling hint (rtx_params.pixel_order): [tiles] order | [tiles] cost | two histograms, fill counters | cells | [tiles] key; the order
    // belongs to the pixel space of the call that produced the costs (order_key)
    void* d_tile_mem = nullptr;
    size_t tile_mem_cap = 0;
    long long order_key[6] = {};
    bool order_valid = false;
    int order_parity = 0;

    Slot slot[kSlots];
    int next_slot = 0;                 // slot the next call takes
    int oldest = 0;                    // oldest pending slot (rtx_wait order)

    double* d_rad_scratch = nullptr;   // radiance buffer for the unfused quantise path (stream-ordered reuse)
    size_t d_rad_scratch_cap = 0;
    void* d_tm_sums = nullptr;         // tone-map extension: per-frame fixed-point log-luminance sums
    size_t d_tm_sums_cap = 0;
    std::vector<long long> h_tm_sums;
    // standalone calls (quantise / tonemap / ...) use slot 0's counters and these events
    cudaEvent_t ev[6] = {};
    void* d_aux_out = nullptr;         // device staging of rtx_quantise / rtx_tonemap with host pointers
    size_t d_aux_out_cap = 0;
    unsigned long long* d_aux_counters = nullptr;
    unsigned long long* h_aux_counters = nullptr;   // pinned
    std::vector<std::pair<void*, size_t>> shared_host;   // rtx_host_shared_open mappings (unmapped by rtx_destroy)
};

namespace {

thread_local std::string g_create_error;

int fail(rtx_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->error = msg;
    return code;
}

int cuda_fail(rtx_ctx* ctx, cudaError_t e, const char* what)
{
    return fail(ctx, RTX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

#define RTX_CUDA(ctx, call)                                     \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// host doubles in the reference's operation order (vec.cpp)
struct h3 { double x, y, z; };
inline h3 H(const rtx_vec3& v) { return h3{v.x, v.y, v.z}; }
inline double hlen(h3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline h3 hdiv(h3 a, double s) { return h3{a.x / s, a.y / s, a.z / s}; }
inline h3 hunit(h3 a) { return hdiv(a, hlen(a)); }
inline h3 hcross(h3 u, h3 v) { return h3{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
inline h3 hsub(h3 a, h3 b) { return h3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline h3 hadd(h3 a, h3 b) { return h3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline h3 hmul(h3 a, double s) { return h3{a.x * s, a.y * s, a.z * s}; }
inline d3 D(h3 a) { return d3{a.x, a.y, a.z}; }
inline double amax3(h3 a) { return std::fmax(std::fabs(a.x), std::fmax(std::fabs(a.y), std::fabs(a.z))); }

// Bounding-sphere entry of a rectangle for the FP32 screen. An unbounded rectangle (non-finite extent) gets r = +inf,
// which passes the screen for every ray — provided the centre is finite: half an infinite extent times a basis vector
// with zero components is NaN, and a NaN test result would REJECT the entry. The centre is irrelevant then: use 0.
inline float4 screen_entry(h3 centre, float r32)
{
    if (std::isinf(r32)) return make_float4(0.f, 0.f, 0.f, r32);
    return make_float4(static_cast<float>(centre.x), static_cast<float>(centre.y), static_cast<float>(centre.z), r32);
}

int grow(rtx_ctx* ctx, void** p, size_t* cap, size_t need)
{
    if (need <= *cap) return RTX_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) return fail(ctx, RTX_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *cap = need;
    return RTX_OK;
}

#ifndef RTX_PAIRS
#define RTX_PAIRS 6
#endif
constexpr int kPad = 2 * RTX_PAIRS;   // entries per hot-loop iteration of trace.cu (kPairsPerIter * 2)

}  // namespace

extern "C" {

int rtx_abi_version(void) { return RTX_ABI_VERSION; }

const char* rtx_status_string(int status)
{
    switch (status) {
        case RTX_OK: return "ok";
        case RTX_ERR_INVALID: return "invalid argument";
        case RTX_ERR_CUDA: return "CUDA error or no usable device";
        case RTX_ERR_NO_SCENE: return "no scene set";
        case RTX_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

int rtx_create(rtx_ctx** out, int device)
{
    if (!out) return RTX_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
        return RTX_ERR_CUDA;   // no CPU fallback, by design
    }
    if (device < 0 || device >= count) return RTX_ERR_INVALID;
    rtx_ctx* ctx = new (std::nothrow) rtx_ctx;
    if (!ctx) return RTX_ERR_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete ctx;
        return RTX_ERR_CUDA;
    }
    if (prop.major < 10) {
        g_create_error = "device is not sm_100-class (kernels are built for sm_100a only)";
        delete ctx;
        return RTX_ERR_CUDA;
    }
    ctx->n_sms = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& ev : ctx->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_scene, cudaEventDisableTiming) == cudaSuccess;
    for (Slot& sl : ctx->slot) {
        for (auto& ev : sl.ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
        for (auto& ev : sl.ev_range) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&sl.d_counters, 8 * sizeof(unsigned long long)) == cudaSuccess;
        ok = ok && cudaHostAlloc(&sl.h_counters, 24 * sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
        if (ok)   // [0..7] read-back area, [8..15] reset template, [16..23] first pixel of each range
            for (int k = 0; k < 8; k++) sl.h_counters[8 + k] = (k >= 4 && k <= 6) ? ~0ull : 0ull;
    }
    ok = ok && cudaMalloc(&ctx->d_aux_counters, 8 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaHostAlloc(&ctx->h_aux_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        g_create_error = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError());
        rtx_destroy(ctx);
        return RTX_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return RTX_OK;
}

void rtx_destroy(rtx_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->ev_scene) cudaEventDestroy(ctx->ev_scene);
    for (Slot& sl : ctx->slot) {
        for (auto& ev : sl.ev)
            if (ev) cudaEventDestroy(ev);
        for (auto& ev : sl.ev_range)
            if (ev) cudaEventDestroy(ev);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.d_cameras) cudaFree(sl.d_cameras);
        if (sl.h_cameras) cudaFreeHost(sl.h_cameras);
        if (sl.d_rays) cudaFree(sl.d_rays);
        if (sl.h_rays) cudaFreeHost(sl.h_rays);
        if (sl.d_counters) cudaFree(sl.d_counters);
        if (sl.h_counters) cudaFreeHost(sl.h_counters);
        for (auto& p : sl.d_out)
            if (p) cudaFree(p);
    }
    if (ctx->d_aux_counters) cudaFree(ctx->d_aux_counters);
    if (ctx->h_aux_counters) cudaFreeHost(ctx->h_aux_counters);
    if (ctx->d_aux_out) cudaFree(ctx->d_aux_out);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d_scene_blob) cudaFree(ctx->d_scene_blob);
    if (ctx->h_scene_blob) cudaFreeHost(ctx->h_scene_blob);
    if (ctx->d_rad_scratch) cudaFree(ctx->d_rad_scratch);
    if (ctx->d_tm_sums) cudaFree(ctx->d_tm_sums);
    if (ctx->d_grid_blob) cudaFree(ctx->d_grid_blob);
    if (ctx->d_tail_scratch) cudaFree(ctx->d_tail_scratch);
    if (ctx->d_tile_mem) cudaFree(ctx->d_tile_mem);
    for (auto& m : ctx->shared_host) {
        cudaHostUnregister(m.first);
        munmap(m.first, m.second);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* rtx_last_error(const rtx_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int rtx_set_stream(rtx_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return RTX_ERR_INVALID;
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    // work already queued (scene upload, calls in flight) stays ordered before anything on the new stream
    RTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RTX_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->stream = cuda_stream == RTX_STREAM_PRIVATE ? ctx->own_stream : static_cast<cudaStream_t>(cuda_stream);
    return RTX_OK;
}

int rtx_set_scene(rtx_ctx* ctx, const rtx_object* objects, int32_t n)
{
    if (!ctx) return RTX_ERR_INVALID;
    if (n < 0 || (n > 0 && !objects)) return fail(ctx, RTX_ERR_INVALID, "rtx_set_scene: null objects or negative count");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<float4> sph32, wall32;
    std::vector<SphereExact> sph64;
    if (n > (1 << 28) - 1) return fail(ctx, RTX_ERR_INVALID, "rtx_set_scene: more than 2^28 - 1 objects");   // key = id * 8 + face
    std::vector<int32_t> sph_key, kind(n), slot(n);
    std::vector<WallDev> walls;
    std::vector<MaterialDev> mats(n);
    double bound = 0.0;
    // Extent of the scene for the screen's error bound. Objects with non-finite geometry do not widen it (the bound
    // would become infinite and send every entry of every ray to the exact test): they either have r = +inf in the
    // screen (always pass) or NaN results there (rejected, and the reference's NaN comparisons never hit them either).
    auto extend = [&bound](double x) { if (std::isfinite(x)) bound = std::fmax(bound, x); };
    for (int k = 0; k < n; k++) {
        const rtx_object& o = objects[k];
        if (o.kind != RTX_SPHERE && o.kind != RTX_WALL && o.kind != RTX_BOX) return fail(ctx, RTX_ERR_INVALID, "rtx_set_scene: unknown object kind");
        MaterialDev& m = mats[k];
        m.color = d3{o.mat.color.x, o.mat.color.y, o.mat.color.z};
        m.ambient = o.mat.ambient;
        m.metallic = o.mat.metallic;
        m.diffuse = o.mat.diffuse;
        m.specular = o.mat.specular;
        m.exponent = o.mat.specular_exponent;
        kind[k] = o.kind;
        if (o.kind == RTX_SPHERE) {
            slot[k] = static_cast<int32_t>(sph64.size());
            sph64.push_back(SphereExact{o.p.x, o.p.y, o.p.z, o.a});
            // float copies for the screen; a non-finite radius gets r = +inf there so that the screen never
            // rejects it and the exact test decides
            float r32 = std::nextafter(static_cast<float>(std::fabs(o.a)), INFINITY);   // radius enters squared: sign is irrelevant
            if (!std::isfinite(o.a)) r32 = INFINITY;
            sph32.push_back(make_float4(static_cast<float>(o.p.x), static_cast<float>(o.p.y), static_cast<float>(o.p.z), r32));
            sph_key.push_back(k * 8);
            extend(amax3(H(o.p)) + std::fabs(o.a));
        } else if (o.kind == RTX_BOX) {
            // EXTENSION (rtx_b200.h, RTX_BOX): six Wall-like faces -x,+x,-y,+y,-z,+z that share the box's id; the trace
            // kernel sees them as walls (same exact test, same bounding-sphere screen), the face index is the low
            // part of the tie-break key and selects the normal when shading.
            slot[k] = static_cast<int32_t>(walls.size());
            const h3 axis[3] = {h3{1, 0, 0}, h3{0, 1, 0}, h3{0, 0, 1}};
            const double size[3] = {o.n.x, o.n.y, o.n.z};
            for (int f = 0; f < 6; f++) {
                const int ax = f >> 1, hi = f & 1;
                const int r_ax = ax == 0 ? 1 : 0, u_ax = ax == 2 ? 1 : 2;
                WallDev w;
                const h3 corner = hi ? hadd(H(o.p), hmul(axis[ax], size[ax])) : H(o.p);
                w.p = D(corner);
                w.n = D(hi ? axis[ax] : hmul(axis[ax], -1.0));
                w.right = D(axis[r_ax]);
                w.up = D(axis[u_ax]);
                w.length = size[r_ax];
                w.width = size[u_ax];
                w.key = k * 8 + f;
                w.pad = 0;
                walls.push_back(w);
                const h3 centre = hadd(hadd(corner, hmul(axis[r_ax], size[r_ax] / 2)), hmul(axis[u_ax], size[u_ax] / 2));
                float r32 = static_cast<float>(0.5 * std::sqrt(size[r_ax] * size[r_ax] + size[u_ax] * size[u_ax]) * (1.0 + 1e-6));
                r32 = std::nextafter(r32, INFINITY);
                if (!std::isfinite(size[r_ax]) || !std::isfinite(size[u_ax])) r32 = INFINITY;
                wall32.push_back(screen_entry(centre, r32));
            }
            extend(amax3(H(o.p)) + std::fabs(size[0]) + std::fabs(size[1]) + std::fabs(size[2]));
        } else {
            slot[k] = static_cast<int32_t>(walls.size());
            WallDev w;
            const h3 nrm = hunit(H(o.n));                               // Wall ctor, scene.h:71
            const h3 right = hunit(hcross(nrm, h3{0, 0, 1}));           // scene.cpp:18
            const h3 up = hunit(hcross(right, nrm));                    // scene.cpp:19
            w.p = D(H(o.p));
            w.n = D(nrm);
            w.right = D(right);
            w.up = D(up);
            w.length = o.a;
            w.width = o.b;
            w.key = k * 8;
            w.pad = 0;
            walls.push_back(w);
            // Bounding sphere of the rectangle for the FP32 screen: every point Wall::intersect can accept
            // (0 <= x <= length, 0 <= y <= width in the (right, up) frame, scene.cpp:25-29) lies inside it.
            const h3 centre = hadd(hadd(H(o.p), hmul(right, o.a / 2)), hmul(up, o.b / 2));
            float r32 = static_cast<float>(0.5 * std::sqrt(o.a * o.a + o.b * o.b) * (1.0 + 1e-6));
            r32 = std::nextafter(r32, INFINITY);
            if (!std::isfinite(o.a) || !std::isfinite(o.b)) r32 = INFINITY;
            wall32.push_back(screen_entry(centre, r32));
            extend(amax3(H(o.p)) + std::fabs(o.a) + std::fabs(o.b));
        }
    }
    const int ns = static_cast<int>(sph64.size()), nw = static_cast<int>(walls.size());
    const int ne = ns + nw;
    const int ns_pad = (ne + kPad - 1) / kPad * kPad;        // padded entry count
    sph32.insert(sph32.end(), wall32.begin(), wall32.end());
    sph32.resize(ns_pad, make_float4(0.f, 0.f, 0.f, -1.f));

    // one device blob, 256-byte aligned sections
    auto align = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
    size_t off = 0;
    const size_t o_s32 = off; off = align(off + sizeof(float4) * std::max(ns_pad, 1));
    const size_t o_s64 = off; off = align(off + sizeof(SphereExact) * std::max(ns, 1));
    const size_t o_sid = off; off = align(off + sizeof(int32_t) * std::max(ns, 1));
    const size_t o_wal = off; off = align(off + sizeof(WallDev) * std::max(nw, 1));
    const size_t o_mat = off; off = align(off + sizeof(MaterialDev) * std::max(n, 1));
    const size_t o_knd = off; off = align(off + sizeof(int32_t) * std::max(n, 1));
    const size_t o_slt = off; off = align(off + sizeof(int32_t) * std::max(n, 1));
    // Pinned staging + stream-ordered upload: the copy queues behind the kernels of earlier calls (they read the old
    // scene) and in front of those of later calls; the host never waits for the device here.
    if (off > ctx->h_scene_blob_cap) {
        if (ctx->h_scene_blob) {
            RTX_CUDA(ctx, cudaEventSynchronize(ctx->ev_scene));
            cudaFreeHost(ctx->h_scene_blob);
        }
        ctx->h_scene_blob = nullptr;
        ctx->h_scene_blob_cap = 0;
        const size_t cap = std::max<size_t>(off, 1 << 16);
        if (cudaHostAlloc(&ctx->h_scene_blob, cap, cudaHostAllocMapped) != cudaSuccess)
            return fail(ctx, RTX_ERR_NOMEM, "rtx_set_scene: pinned staging");
        ctx->h_scene_blob_cap = cap;
    } else {
        RTX_CUDA(ctx, cudaEventSynchronize(ctx->ev_scene));   // the previous upload has left the staging buffer
    }
    unsigned char* blob = ctx->h_scene_blob;
    std::memset(blob, 0, off);
    if (ns_pad) std::memcpy(&blob[o_s32], sph32.data(), sizeof(float4) * ns_pad);
    if (ns) std::memcpy(&blob[o_s64], sph64.data(), sizeof(SphereExact) * ns);
    if (ns) std::memcpy(&blob[o_sid], sph_key.data(), sizeof(int32_t) * ns);
    if (nw) std::memcpy(&blob[o_wal], walls.data(), sizeof(WallDev) * nw);
    if (n) std::memcpy(&blob[o_mat], mats.data(), sizeof(MaterialDev) * n);
    if (n) std::memcpy(&blob[o_knd], kind.data(), sizeof(int32_t) * n);
    if (n) std::memcpy(&blob[o_slt], slot.data(), sizeof(int32_t) * n);

    ctx->have_scene = false;
    if (off > ctx->scene_blob_cap) {                         // the allocation is kept and reused while the new scene fits
        if (ctx->d_scene_blob) {
            RTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // kernels in flight still read the old allocation
            cudaFree(ctx->d_scene_blob);
        }
        ctx->d_scene_blob = nullptr;
        ctx->scene_blob_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->d_scene_blob, off);
        if (e != cudaSuccess) return fail(ctx, RTX_ERR_NOMEM, std::string("cudaMalloc(scene): ") + cudaGetErrorString(e));
        ctx->scene_blob_cap = off;
    }
    if (off <= kSmallUploadBytes) RTX_CUDA(ctx, launch_small_upload(ctx->d_scene_blob, blob, off, ctx->stream));
    else RTX_CUDA(ctx, cudaMemcpyAsync(ctx->d_scene_blob, blob, off, cudaMemcpyHostToDevice, ctx->stream));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev_scene, ctx->stream));
    unsigned char* base = static_cast<unsigned char*>(ctx->d_scene_blob);
    SceneDev& s = ctx->scene;
    s.n_objects = n;
    s.n_spheres = ns;
    s.n_walls = nw;
    s.n_entries = ne;
    s.n_entries_padded = ns_pad;
    s.ent32 = reinterpret_cast<const float4*>(base + o_s32);
    s.sph64 = reinterpret_cast<const SphereExact*>(base + o_s64);
    s.sph_key = reinterpret_cast<const int32_t*>(base + o_sid);
    s.walls = reinterpret_cast<const WallDev*>(base + o_wal);
    s.mats = reinterpret_cast<const MaterialDev*>(base + o_mat);
    s.kind = reinterpret_cast<const int32_t*>(base + o_knd);
    s.slot = reinterpret_cast<const int32_t*>(base + o_slt);
    ctx->scene_bound = bound;
    ctx->h_sph64 = std::move(sph64);
    ctx->grid_valid = false;
    ctx->have_scene = true;
    ctx->error.clear();
    return RTX_OK;
}

int rtx_camera_init(const rtx_camera_desc* d, rtx_camera* out)
{
    if (!d || !out) return RTX_ERR_INVALID;
    // Camera::init, scene.cpp:80-106
    const double image_width = d->image_width;
    const double image_height = static_cast<int>(image_width / d->aspect_ratio);
    const h3 position = H(d->position), lookat = H(d->lookat), vup = H(d->vup);
    const double focal_length = hlen(hsub(position, lookat));
    const double theta = d->vfov * 3.14 / 180.0;
    const double h = std::tan(theta / 2);
    const double fov_height = 2 * h * focal_length;
    const double fov_width = fov_height * (image_width / image_height);
    const h3 w = hunit(hsub(position, lookat));
    const h3 u = hunit(hcross(vup, w));
    const h3 v = hcross(w, u);
    const h3 fov_x = hmul(u, fov_width);
    const h3 fov_y = hmul(v, -fov_height);
    const h3 dx = hdiv(fov_x, image_width);
    const h3 dy = hdiv(fov_y, image_height);
    const h3 top_left = hsub(hsub(hsub(position, hmul(w, focal_length)), hdiv(fov_x, 2)), hdiv(fov_y, 2));
    const h3 itl = hadd(top_left, hmul(hadd(dx, dy), 0.5));
    out->position = d->position;
    out->image_top_left = rtx_vec3{itl.x, itl.y, itl.z};
    out->delta_x = rtx_vec3{dx.x, dx.y, dx.z};
    out->delta_y = rtx_vec3{dy.x, dy.y, dy.z};
    out->width = static_cast<int32_t>(image_width);
    out->height = static_cast<int32_t>(image_height);
    return RTX_OK;
}

void rtx_default_params(rtx_params* p)
{
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->max_depth = 10;
    p->quantise_mode = RTX_QUANT_WRAP;
    p->fuse_quantise = 1;
    p->light_pos = rtx_vec3{0, 0, 0};
    p->ground_color = rtx_vec3{0.025, 0.05, 0.075};
    p->sky_low = rtx_vec3{0.36, 0.45, 0.57};
    p->sky_high = rtx_vec3{0.14, 0.21, 0.49};
    p->reflect_offset = .0001;
    p->sky_exponent = static_cast<double>(static_cast<float>(1. / 4.));
    p->band_rows = 4;
    p->n_ranks = 1;
    p->rank = 0;
    p->frame_offset = 0;
    p->frame_stride = 1;
    p->sun_enabled = 0;                                   // extensions off: the reference's semantics
    p->tonemap = RTX_TONEMAP_NONE;
    p->sun_color = rtx_vec3{1.64, 1.27, 0.99};            // SUN_COLOR      main.cpp:18
    p->sun_direction = rtx_vec3{.7, .4, .7};              // SUN_DIRECTION  main.cpp:19
    p->tonemap_key = 0.18;
    p->tonemap_white = 0.0;
}

int32_t rtx_local_rows(int32_t height, int32_t band_rows, int32_t n_ranks, int32_t rank)
{
    if (height <= 0 || band_rows <= 0 || n_ranks <= 0 || rank < 0 || rank >= n_ranks) return 0;
    if (n_ranks == 1) return height;
    const int32_t n_bands = (height + band_rows - 1) / band_rows;
    int32_t rows = 0;
    for (int32_t b = rank; b < n_bands; b += n_ranks) rows += std::min(band_rows, height - b * band_rows);
    return rows;
}

int32_t rtx_global_row(int32_t local_row, int32_t height, int32_t band_rows, int32_t n_ranks, int32_t rank)
{
    if (local_row < 0 || local_row >= rtx_local_rows(height, band_rows, n_ranks, rank)) return -1;
    if (n_ranks == 1) return local_row;
    const int32_t lb = local_row / band_rows;
    return (lb * n_ranks + rank) * band_rows + (local_row - lb * band_rows);
}

}  // extern "C"

namespace {

// EXTENSION (rtx_params.accel = RTX_ACCEL_GRID): builds the uniform grid of the current scene on first use.
// A sphere is listed in every cell that its bounding box, inflated by `margin`, overlaps. margin = 1e-6 x scene extent:
// six orders of magnitude above the rounding of the double-precision walk and of the exact test's own accept/reject
// boundary, and far below a cell. Spheres that are not finite, or that would land in more than kMaxCellsPerSphere cells,
// go to the always list instead.
int ensure_grid(rtx_ctx* ctx)
{
    if (ctx->grid_valid) return RTX_OK;
    constexpr int kMaxCellsPerSphere = 512, kMaxDim = 512;
    constexpr size_t kMaxCells = size_t(1) << 21;
    const std::vector<SphereExact>& sp = ctx->h_sph64;
    const int ns = static_cast<int>(sp.size());
    const double B = std::fmax(ctx->scene_bound, 1e-3);
    const double margin = 1e-6 * B;
    std::vector<int32_t> always;
    std::vector<int> in_grid;
    std::vector<double> rad(ns, 0.0);
    // a few huge spheres (a "ground" sphere of radius 1000 under a scene of unit spheres) must not stretch the grid: anything
    // above 16 x the median radius is screened for every ray instead of being registered
    double big = 1e300;
    {
        std::vector<double> radii;
        for (const SphereExact& s : sp)
            if (std::isfinite(s.r)) radii.push_back(std::fabs(s.r));
        if (radii.size() >= 8) {
            std::nth_element(radii.begin(), radii.begin() + radii.size() / 2, radii.end());
            big = 16.0 * std::fmax(radii[radii.size() / 2], 1e-6 * B);
        }
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int e = 0; e < ns; e++) {
        const SphereExact& s = sp[e];
        if (!(std::isfinite(s.cx) && std::isfinite(s.cy) && std::isfinite(s.cz) && std::isfinite(s.r)) || std::fabs(s.r) > big) {
            always.push_back(e);
            continue;
        }
        rad[e] = std::fabs(s.r) * (1.0 + 1e-9) + margin;
        in_grid.push_back(e);
        const double c[3] = {s.cx, s.cy, s.cz};
        for (int k = 0; k < 3; k++) {
            lo[k] = std::fmin(lo[k], c[k] - rad[e]);
            hi[k] = std::fmax(hi[k], c[k] + rad[e]);
        }
    }
    GridDev g = {};
    std::vector<uint32_t> cell_start(1, 0u);
    std::vector<uint32_t> items;
    if (!in_grid.empty()) {
        double ext[3];
        for (int k = 0; k < 3; k++) ext[k] = std::fmax(hi[k] - lo[k], 4.0 * margin);
        const double target = std::fmax(1.0, in_grid.size() / 4.0);            // about four spheres per cell
        double cell = std::cbrt(ext[0] * ext[1] * ext[2] / target);
        cell = std::fmax(cell, std::fmax(ext[0], std::fmax(ext[1], ext[2])) / kMaxDim);
        int n[3];
        for (;;) {
            size_t total = 1;
            for (int k = 0; k < 3; k++) {
                n[k] = std::max(1, std::min(kMaxDim, static_cast<int>(std::ceil(ext[k] / cell))));
                total *= static_cast<size_t>(n[k]);
            }
            if (total <= kMaxCells) break;
            cell *= 1.26;
        }
        g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
        g.x0 = lo[0]; g.y0 = lo[1]; g.z0 = lo[2];
        g.x1 = lo[0] + n[0] * cell; g.y1 = lo[1] + n[1] * cell; g.z1 = lo[2] + n[2] * cell;    // covers hi[] (n = ceil(ext / cell))
        g.cell = cell;
        g.inv_cell = 1.0 / cell;
        g.slack = 1e-7 * B;
        const size_t n_cells = static_cast<size_t>(n[0]) * n[1] * n[2];
        auto range = [&](double c, double r, int k, int& a, int& b) {
            a = std::max(0, std::min(n[k] - 1, static_cast<int>(std::floor((c - r - lo[k]) / cell))));
            b = std::max(0, std::min(n[k] - 1, static_cast<int>(std::floor((c + r - lo[k]) / cell))));
        };
        std::vector<uint32_t> count(n_cells + 1, 0u);
        std::vector<int> kept;
        for (int e : in_grid) {
            int a[3], b[3];
            range(sp[e].cx, rad[e], 0, a[0], b[0]);
            range(sp[e].cy, rad[e], 1, a[1], b[1]);
            range(sp[e].cz, rad[e], 2, a[2], b[2]);
            const long long cells = static_cast<long long>(b[0] - a[0] + 1) * (b[1] - a[1] + 1) * (b[2] - a[2] + 1);
            if (cells > kMaxCellsPerSphere) {
                always.push_back(e);
                continue;
            }
            kept.push_back(e);
            for (int z = a[2]; z <= b[2]; z++)
                for (int y = a[1]; y <= b[1]; y++)
                    for (int x = a[0]; x <= b[0]; x++) count[(static_cast<size_t>(z) * n[1] + y) * n[0] + x + 1]++;
        }
        for (size_t k = 0; k < n_cells; k++) count[k + 1] += count[k];
        cell_start = count;
        items.resize(cell_start[n_cells]);
        std::vector<uint32_t> cursor(cell_start.begin(), cell_start.end() - 1);
        for (int e : kept) {
            int a[3], b[3];
            range(sp[e].cx, rad[e], 0, a[0], b[0]);
            range(sp[e].cy, rad[e], 1, a[1], b[1]);
            range(sp[e].cz, rad[e], 2, a[2], b[2]);
            for (int z = a[2]; z <= b[2]; z++)
                for (int y = a[1]; y <= b[1]; y++)
                    for (int x = a[0]; x <= b[0]; x++) items[cursor[(static_cast<size_t>(z) * n[1] + y) * n[0] + x]++] = static_cast<uint32_t>(e);
        }
    }
    g.n_items = static_cast<int32_t>(items.size());
    g.items16 = ns <= 65535 ? 1 : 0;
    g.n_always = static_cast<int32_t>(always.size());
    // one blob: cell_start | items (16 or 32 bit) | always
    auto align = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
    const size_t o_cs = 0;
    const size_t o_it = align(cell_start.size() * 4);
    const size_t it_bytes = g.items16 ? ((items.size() + 1) / 2) * 4 : items.size() * 4;
    const size_t o_al = align(o_it + std::max<size_t>(it_bytes, 4));
    const size_t total = align(o_al + std::max<size_t>(always.size() * 4, 4));
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(&blob[o_cs], cell_start.data(), cell_start.size() * 4);
    if (g.items16) {
        uint16_t* d = reinterpret_cast<uint16_t*>(&blob[o_it]);
        for (size_t k = 0; k < items.size(); k++) d[k] = static_cast<uint16_t>(items[k]);
    } else if (!items.empty()) {
        std::memcpy(&blob[o_it], items.data(), items.size() * 4);
    }
    if (!always.empty()) std::memcpy(&blob[o_al], always.data(), always.size() * 4);
    RTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // a kernel of an earlier call may still walk the old grid
    int rc = grow(ctx, &ctx->d_grid_blob, &ctx->grid_blob_cap, total);
    if (rc != RTX_OK) return rc;
    RTX_CUDA(ctx, cudaMemcpy(ctx->d_grid_blob, blob.data(), total, cudaMemcpyHostToDevice));
    unsigned char* base = static_cast<unsigned char*>(ctx->d_grid_blob);
    g.cell_start = reinterpret_cast<const uint32_t*>(base + o_cs);
    g.items = base + o_it;
    g.always = reinterpret_cast<const int32_t*>(base + o_al);
    ctx->grid = g;
    ctx->grid_valid = true;
    return RTX_OK;
}

// Waits for the call in `sl` and turns its events / counters into rtx_stats.
int finish_slot(rtx_ctx* ctx, Slot& sl, rtx_stats* stats)
{
    sl.pending = false;
    RTX_CUDA(ctx, cudaEventSynchronize(sl.ev_done));
    RTX_CUDA(ctx, cudaGetLastError());
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, sl.ev[0], sl.ev[1]); stats->h2d_ms = ms;
        cudaEventElapsedTime(&ms, sl.ev[1], sl.ev[2]); stats->raytracing_ms = ms;
        cudaEventElapsedTime(&ms, sl.ev[2], sl.ev[3]); stats->surface_update_ms = ms;
        cudaEventElapsedTime(&ms, sl.ev[3], sl.ev[4]); stats->d2h_ms = ms;
        cudaEventElapsedTime(&ms, sl.ev[0], sl.ev[4]); stats->total_ms = ms;
        stats->total_rays = sl.h_counters[1];
        stats->sphere_tests = stats->total_rays * static_cast<uint64_t>(sl.n_spheres);
        stats->wall_tests = stats->total_rays * static_cast<uint64_t>(sl.n_walls);
        stats->over_range_pixels = sl.h_counters[2];
        long long bits = static_cast<long long>(sl.h_counters[3]);
        std::memcpy(&stats->max_luminance, &bits, sizeof bits);
        stats->launches = sl.launches;
        const unsigned long long t0 = sl.h_counters[4], dry = sl.h_counters[5], first = sl.h_counters[6], last = sl.h_counters[7];
        if (t0 != ~0ull && last >= t0) {
            stats->drain_ms = (dry != ~0ull && last >= dry) ? (last - dry) * 1e-6 : 0.0;
            stats->exit_spread_ms = (first != ~0ull && last >= first) ? (last - first) * 1e-6 : 0.0;
        }
    }
    return RTX_OK;
}

// rtx_render / rtx_render_async / rtx_trace_rays: validate, allocate, THEN enqueue (an error after the first enqueue
// synchronises both streams before returning, so the device never writes a caller's buffer after an error return).
int render_impl(rtx_ctx* ctx, const rtx_camera* cams, int32_t n_frames, const rtx_ray* rays, int64_t n_rays, const rtx_params* params,
                const rtx_outputs* outs, rtx_stats* stats, bool async, const char* who)
{
    if (!ctx) return RTX_ERR_INVALID;
    const std::string W_(who);
    if (!ctx->have_scene) return fail(ctx, RTX_ERR_NO_SCENE, W_ + ": call rtx_set_scene first");
    const bool ray_mode = rays != nullptr || cams == nullptr;
    if (!params || !outs) return fail(ctx, RTX_ERR_INVALID, W_ + ": null argument");
    if (ray_mode) {
        if (!rays || n_rays <= 0 || n_rays > 0x7fffffff) return fail(ctx, RTX_ERR_INVALID, W_ + ": null rays or n_rays out of [1, 2^31)");
    } else if (!cams || n_frames <= 0) {
        return fail(ctx, RTX_ERR_INVALID, W_ + ": null argument or n_frames <= 0");
    }
    const rtx_params& p = *params;
    if (p.max_depth < 0 || p.max_depth > RTX_MAX_DEPTH) return fail(ctx, RTX_ERR_INVALID, W_ + ": max_depth out of [0, 254]");
    if (p.n_ranks < 1 || p.rank < 0 || p.rank >= p.n_ranks || (p.n_ranks > 1 && p.band_rows < 1))
        return fail(ctx, RTX_ERR_INVALID, W_ + ": bad band sharding (band_rows, n_ranks, rank)");
    if (p.quantise_mode != RTX_QUANT_WRAP && p.quantise_mode != RTX_QUANT_SATURATE)
        return fail(ctx, RTX_ERR_INVALID, W_ + ": unknown quantise_mode");
    if (outs->memory != RTX_MEM_HOST && outs->memory != RTX_MEM_DEVICE && outs->memory != RTX_MEM_HOST_MAPPED)
        return fail(ctx, RTX_ERR_INVALID, W_ + ": outputs.memory must be RTX_MEM_HOST, RTX_MEM_DEVICE or RTX_MEM_HOST_MAPPED");
    if (p.tonemap != RTX_TONEMAP_NONE && p.tonemap != RTX_TONEMAP_REINHARD) return fail(ctx, RTX_ERR_INVALID, W_ + ": unknown tonemap");
    if (p.accel != RTX_ACCEL_NONE && p.accel != RTX_ACCEL_GRID) return fail(ctx, RTX_ERR_INVALID, W_ + ": unknown accel");
    if (p.pixel_order < RTX_ORDER_AUTO || p.pixel_order > RTX_ORDER_COST) return fail(ctx, RTX_ERR_INVALID, W_ + ": unknown pixel_order");
    const bool tonemap = p.tonemap == RTX_TONEMAP_REINHARD && outs->rgba8;
    if (tonemap && (p.n_ranks != 1 || outs->frame_rgba8 || ray_mode))
        return fail(ctx, RTX_ERR_INVALID, W_ + ": the tone-map operator needs the whole frame on one GPU (n_ranks = 1, no frame_rgba8, no ray batch)");
    if (tonemap && !(p.tonemap_key > 0.0)) return fail(ctx, RTX_ERR_INVALID, W_ + ": tonemap_key must be > 0");
    if (ray_mode && (p.n_ranks != 1 || outs->frame_rgba8))
        return fail(ctx, RTX_ERR_INVALID, W_ + ": a ray batch is not sharded (n_ranks = 1, no frame_rgba8)");
    if (outs->frame_rgba8 && (p.frame_stride < 1 || p.frame_offset < 0))
        return fail(ctx, RTX_ERR_INVALID, W_ + ": frame_offset/frame_stride must be >= 0 / >= 1");
    if (outs->frame_mode != RTX_FRAME_STORE && outs->frame_mode != RTX_FRAME_COPY)
        return fail(ctx, RTX_ERR_INVALID, W_ + ": unknown frame_mode");
    const bool frame_copy = outs->frame_rgba8 && outs->frame_mode == RTX_FRAME_COPY;
    if (frame_copy && outs->rgba8) return fail(ctx, RTX_ERR_INVALID, W_ + ": RTX_FRAME_COPY uses the rgba8 plane as its staging: pass rgba8 = NULL");
    int W, Hh;
    double cam_bound = 0.0;
    if (ray_mode) {
        W = static_cast<int>(n_rays);
        Hh = 1;
        n_frames = 1;
        for (int64_t k = 0; k < n_rays; k++) {
            const double m = amax3(H(rays[k].origin));
            if (std::isfinite(m)) cam_bound = std::fmax(cam_bound, m);    // a far or non-finite origin takes the exact fallback anyway
        }
        cam_bound = std::fmin(cam_bound, 4.0 * std::fmax(ctx->scene_bound, 1e-3));   // keep the screen's E tight: outliers fall back
    } else {
        W = cams[0].width;
        Hh = cams[0].height;
        if (W <= 0 || Hh <= 0) return fail(ctx, RTX_ERR_INVALID, W_ + ": camera width/height must be positive");
        for (int f = 0; f < n_frames; f++) {
            if (cams[f].width != W || cams[f].height != Hh)
                return fail(ctx, RTX_ERR_INVALID, W_ + ": all cameras of one call must share width/height");
            cam_bound = std::fmax(cam_bound, amax3(H(cams[f].position)));
        }
    }
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int band_rows = p.n_ranks > 1 ? p.band_rows : 1;
    const int local_rows = rtx_local_rows(Hh, band_rows, p.n_ranks, p.rank);
    const size_t n_px = static_cast<size_t>(n_frames) * local_rows * W;

    // ---- slot: at most kSlots calls in flight ---------------------------------------------------------------------
    if (!async) {
        // a synchronous call first completes whatever is in flight, oldest first (those calls' stats are dropped)
        while (ctx->slot[ctx->oldest].pending) {
            Slot& o = ctx->slot[ctx->oldest];
            ctx->oldest = (ctx->oldest + 1) % kSlots;
            int rc = finish_slot(ctx, o, nullptr);
            if (rc != RTX_OK) return rc;
        }
        ctx->oldest = ctx->next_slot;
    }
    Slot& sl = ctx->slot[ctx->next_slot];
    if (sl.pending) return fail(ctx, RTX_ERR_INVALID, W_ + ": too many calls in flight (rtx_wait first)");

    // ---- allocations (all before the first enqueue) -----------------------------------------------------------------
    if (!ray_mode && n_frames > sl.cameras_cap) {
        if (sl.d_cameras) cudaFree(sl.d_cameras);
        if (sl.h_cameras) cudaFreeHost(sl.h_cameras);
        sl.d_cameras = nullptr;
        sl.h_cameras = nullptr;
        sl.cameras_cap = 0;
        const int cap = std::max(n_frames, 16);
        // one spare element: the kernel-side upload moves whole 16-byte words (sizeof(rtx_camera) = 104)
        if (cudaMalloc(&sl.d_cameras, sizeof(rtx_camera) * (cap + 1)) != cudaSuccess ||
            cudaHostAlloc(&sl.h_cameras, sizeof(rtx_camera) * (cap + 1), cudaHostAllocMapped) != cudaSuccess)
            return fail(ctx, RTX_ERR_NOMEM, W_ + ": camera buffers");
        sl.cameras_cap = cap;
    }
    if (ray_mode && static_cast<size_t>(n_rays) > sl.rays_cap) {
        if (sl.d_rays) cudaFree(sl.d_rays);
        if (sl.h_rays) cudaFreeHost(sl.h_rays);
        sl.d_rays = nullptr;
        sl.h_rays = nullptr;
        sl.rays_cap = 0;
        const size_t cap = std::max<size_t>(n_rays, 256);
        if (cudaMalloc(&sl.d_rays, sizeof(rtx_ray) * cap) != cudaSuccess ||
            cudaHostAlloc(&sl.h_rays, sizeof(rtx_ray) * cap, cudaHostAllocMapped) != cudaSuccess)
            return fail(ctx, RTX_ERR_NOMEM, W_ + ": ray buffers");
        sl.rays_cap = cap;
    }
    // output planes: the caller's device pointers, the device alias of the caller's mapped host memory, or device
    // staging for plain host pointers
    void* user[kPlanes] = {outs->rgba8, outs->radiance_f32, outs->radiance_f64, outs->object_id, outs->hit_mask, outs->ray_count,
                           outs->hit_distance, outs->hit_normal};
    const size_t elem[kPlanes] = {4, 12, 24, 4, 1, 1, 8, 24};
    void* dev[kPlanes] = {};
    const bool host_out = outs->memory == RTX_MEM_HOST;
    for (int k = 0; k < kPlanes; k++) {
        if (!user[k]) continue;
        if (host_out) {
            int rc = grow(ctx, &sl.d_out[k], &sl.d_out_cap[k], std::max<size_t>(n_px * elem[k], 16));
            if (rc != RTX_OK) return rc;
            dev[k] = sl.d_out[k];
        } else if (outs->memory == RTX_MEM_HOST_MAPPED) {
            // zero copy: the kernel stores straight into the caller's pinned, mapped surface (SDL's surface->pixels, main.cpp:193,344)
            if (cudaHostGetDevicePointer(&dev[k], user[k], 0) != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, RTX_ERR_INVALID, W_ + ": RTX_MEM_HOST_MAPPED needs memory from rtx_host_alloc / rtx_host_register / rtx_host_shared_open");
            }
        } else {
            dev[k] = user[k];
        }
    }
    if (frame_copy) {   // the packed local frame(s) in context staging; bulk copies place them afterwards
        int rc = grow(ctx, &sl.d_out[0], &sl.d_out_cap[0], std::max<size_t>(n_px * 4, 16));
        if (rc != RTX_OK) return rc;
    }
    const bool unfused = (!p.fuse_quantise || tonemap) && user[0] && !outs->frame_rgba8;
    double* rad_for_quant = nullptr;
    if (unfused) {
        if (dev[2]) {
            rad_for_quant = static_cast<double*>(dev[2]);
        } else {
            void* pp = ctx->d_rad_scratch;
            int rc = grow(ctx, &pp, &ctx->d_rad_scratch_cap, std::max<size_t>(n_px * 24, 16));
            ctx->d_rad_scratch = static_cast<double*>(pp);
            if (rc != RTX_OK) return rc;
            rad_for_quant = ctx->d_rad_scratch;
        }
        if (tonemap) {
            int rc = grow(ctx, &ctx->d_tm_sums, &ctx->d_tm_sums_cap, sizeof(long long) * std::max(n_frames, 16));
            if (rc != RTX_OK) return rc;
        }
    }

    if (ctx->scene.n_entries > kSmallSceneEntries) {
        int rc = grow(ctx, &ctx->d_tail_scratch, &ctx->tail_scratch_cap, static_cast<size_t>(ctx->n_sms) * kTailScratchBytesPerCta);
        if (rc != RTX_OK) return rc;
    }
    const bool use_grid = p.accel == RTX_ACCEL_GRID && ctx->scene.n_entries > kSmallSceneEntries;
    if (use_grid) {
        int rc = ensure_grid(ctx);
        if (rc != RTX_OK) return rc;
    }

    // Scheduling hint for the big kernel: tiles of 256 pixels are handed out most expensive first, by the ray counts the
    // previous call with the same pixel space collected (aux_kernels.cu). Only WHEN a pixel is traced changes, nothing else.
    const size_t tiles_all = (n_px + ((size_t{1} << kOrderTileShift) - 1)) >> kOrderTileShift, tiles_full = n_px >> kOrderTileShift;
    const bool hint = !ray_mode && !use_grid && ctx->scene.n_entries > kSmallSceneEntries && tiles_full >= 2 &&
                      (p.pixel_order == RTX_ORDER_COST || (p.pixel_order == RTX_ORDER_AUTO && n_px >= kOrderMinPixels));
    bool hint_reset = false;
    uint32_t *tile_order = nullptr, *tile_cost = nullptr, *tile_hist = nullptr;
    const long long packed_rows = static_cast<long long>(n_frames) * local_rows;
    const size_t order_cells = hint ? tile_order_cells(W, packed_rows) : 0;
    if (hint) {
        const void* before = ctx->d_tile_mem;
        int rc = grow(ctx, &ctx->d_tile_mem, &ctx->tile_mem_cap, (3 * tiles_all + 3 * kOrderKeys + order_cells) * sizeof(uint32_t));
        if (rc != RTX_OK) {
            ctx->order_valid = false;
            return rc;
        }
        const long long key[6] = {n_frames, W, local_rows, band_rows, p.n_ranks, p.rank};
        if (before != ctx->d_tile_mem || std::memcmp(key, ctx->order_key, sizeof key) != 0) {
            hint_reset = true;
            ctx->order_valid = false;
            ctx->order_parity = 0;
            std::memcpy(ctx->order_key, key, sizeof key);
        }
        tile_order = static_cast<uint32_t*>(ctx->d_tile_mem);
        tile_cost = tile_order + tiles_all;
        tile_hist = tile_cost + tiles_all;                 // two histograms, the fill counters, the cells, then 16 bits of key per tile
    }

    if (ray_mode) std::memcpy(sl.h_rays, rays, sizeof(rtx_ray) * n_rays);
    else std::memcpy(sl.h_cameras, cams, sizeof(rtx_camera) * n_frames);

    TraceArgs a = {};
    a.scene = ctx->scene;
    a.cameras = ray_mode ? nullptr : sl.d_cameras;
    a.rays = ray_mode ? sl.d_rays : nullptr;
    a.n_frames = n_frames;
    a.width = W;
    a.height = Hh;
    a.local_rows = local_rows;
    a.band_rows = band_rows;
    a.n_ranks = p.n_ranks;
    a.rank = p.rank;
    a.max_depth = p.max_depth;
    a.quantise_mode = p.quantise_mode;
    a.light = D(H(p.light_pos));
    a.ground = D(H(p.ground_color));
    a.sky_low = D(H(p.sky_low));
    a.sky_high = D(H(p.sky_high));
    a.reflect_offset = p.reflect_offset;
    a.sky_exponent = p.sky_exponent;
    a.sun_enabled = p.sun_enabled ? 1 : 0;
    a.sun_dir = D(hunit(H(p.sun_direction)));             // normalised like every direction of the reference (vec.cpp:21-24)
    a.sun_color = D(H(p.sun_color));
    // FP32 screen error bound (derivation in DESIGN.md §3.2): with B = scene/camera extent and ray origins
    // within 2B, the line-distance error is below 1.7e-6*B; 8e-6*B leaves a 4x margin. Origins beyond 2B
    // (primary-ray overshoot) fall back to exact tests lane by lane.
    const double B = std::fmax(std::fmax(ctx->scene_bound, cam_bound), 1e-3);
    a.filter_eps = static_cast<float>(8e-6 * B);
    a.origin_bound = static_cast<float>(2.0 * B);
    a.rgba8 = unfused ? nullptr : static_cast<uint32_t*>(dev[0]);
    a.rad32 = static_cast<float*>(dev[1]);
    a.rad64 = unfused ? rad_for_quant : static_cast<double*>(dev[2]);
    a.object_id = static_cast<int32_t*>(dev[3]);
    a.hit_mask = static_cast<uint8_t*>(dev[4]);
    a.ray_count = static_cast<uint8_t*>(dev[5]);
    a.hit_distance = static_cast<double*>(dev[6]);
    a.hit_normal = static_cast<double*>(dev[7]);
    a.frame_rgba8 = frame_copy ? nullptr : outs->frame_rgba8;
    if (frame_copy) a.rgba8 = static_cast<uint32_t*>(sl.d_out[0]);
    a.frame_offset = p.frame_offset;
    a.frame_stride = p.frame_stride;
    a.counters = sl.d_counters;
    a.tail_scratch = ctx->d_tail_scratch;
    a.use_grid = use_grid ? 1 : 0;
    if (use_grid) a.grid = ctx->grid;
    a.tile_cost = tile_cost;
    a.tile_order = hint && ctx->order_valid ? tile_order : nullptr;

    // ---- enqueue ----------------------------------------------------------------------------------------------------------
    cudaStream_t cs = ctx->copy_stream;
    auto bail = [&](cudaError_t e, const char* what) {
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(cs);
        ctx->order_valid = false;      // the scheduling hint starts over: its counters may be half-way through a frame
        ctx->order_key[0] = -1;
        return cuda_fail(ctx, e, what);
    };
#define RTX_ENQ(call)                                        \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return bail(e__, #call);     \
    } while (0)

    int launches = 0;
    RTX_ENQ(cudaEventRecord(sl.ev[0], st));
    {
        void* d_in = ray_mode ? static_cast<void*>(sl.d_rays) : static_cast<void*>(sl.d_cameras);
        const void* h_in = ray_mode ? static_cast<const void*>(sl.h_rays) : static_cast<const void*>(sl.h_cameras);
        const size_t in_bytes = ray_mode ? sizeof(rtx_ray) * static_cast<size_t>(n_rays) : sizeof(rtx_camera) * static_cast<size_t>(n_frames);
        if (in_bytes <= kSmallUploadBytes) RTX_ENQ(launch_small_upload(d_in, h_in, in_bytes, st));      // no copy engine: see aux_kernels.cu
        else RTX_ENQ(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, st));
    }
    // counters: [0..3] = 0, [4..6] = ~0 (atomicMin slots), [7] = 0
    RTX_ENQ(launch_reset_counters(sl.d_counters, 0ull, true, st));
    if (hint_reset) {
        RTX_ENQ(launch_tile_reset(tile_order, tile_cost, static_cast<int>(tiles_all), tile_hist, static_cast<int>(3 * kOrderKeys + order_cells), st));
        launches++;
    }
    RTX_ENQ(cudaEventRecord(sl.ev[1], st));
    // A small scene rendered into HOST memory is bound by the PCIe read-back, not by the kernel (1080p: 0.08 ms of
    // tracing, 0.15 ms of copy): trace the frame as kRanges consecutive pixel ranges and copy each finished range on a
    // second stream while the next one is traced. (Not for the big kernel: every launch of it has its own drain tail. Not for
    // rtx_render_async either: there the whole read-back already overlaps the NEXT frame's kernel, and a stream of small
    // frames is bound by the host's launch path — one launch and one copy per frame instead of four of each.)
    const bool ranged = host_out && !async && !unfused && !frame_copy && ctx->scene.n_entries <= kSmallSceneEntries && n_px >= kRangedMinPixels;
    if (ranged) {
        for (int c = 0; c < kRanges; c++) {
            const size_t p0 = (n_px * c / kRanges) & ~static_cast<size_t>(3), p1 = c + 1 == kRanges ? n_px : (n_px * (c + 1) / kRanges) & ~static_cast<size_t>(3);
            a.pixel_begin = p0;
            a.pixel_end = p1;
            if (c > 0) RTX_ENQ(launch_reset_counters(sl.d_counters, p0, false, st));   // the pixel pool of this launch starts at p0
            RTX_ENQ(launch_trace(a, ctx->n_sms, st, &launches, &ctx->launch_state));
            RTX_ENQ(cudaEventRecord(sl.ev_range[c], st));
        }
        // all launches are queued before the first copy: a copy into PAGEABLE host memory blocks this thread, and the
        // kernels behind it should already be running then
        for (int c = 0; c < kRanges; c++) {
            const size_t p0 = (n_px * c / kRanges) & ~static_cast<size_t>(3), p1 = c + 1 == kRanges ? n_px : (n_px * (c + 1) / kRanges) & ~static_cast<size_t>(3);
            RTX_ENQ(cudaStreamWaitEvent(cs, sl.ev_range[c], 0));
            for (int k = 0; k < kPlanes; k++)
                if (user[k])
                    RTX_ENQ(cudaMemcpyAsync(static_cast<char*>(user[k]) + p0 * elem[k], static_cast<char*>(dev[k]) + p0 * elem[k],
                                            (p1 - p0) * elem[k], cudaMemcpyDeviceToHost, cs));
        }
    } else {
        RTX_ENQ(launch_trace(a, ctx->n_sms, st, &launches, &ctx->launch_state));
    }
    RTX_ENQ(cudaEventRecord(sl.ev[2], st));
    if (unfused) {
        RTX_ENQ(cudaMemsetAsync(sl.d_counters + 2, 0, 2 * sizeof(unsigned long long), st));
        if (tonemap) {
            const int64_t ppf = static_cast<int64_t>(local_rows) * W;
            long long* sums = static_cast<long long*>(ctx->d_tm_sums);
            RTX_ENQ(cudaMemsetAsync(sums, 0, sizeof(long long) * n_frames, st));
            RTX_ENQ(launch_tonemap_sums(nullptr, rad_for_quant, ppf, n_frames, sums, ctx->n_sms, st));
            RTX_ENQ(launch_tonemap_apply(nullptr, rad_for_quant, ppf, n_frames, sums, ppf, p.tonemap_key, p.tonemap_white,
                                         p.quantise_mode, static_cast<uint32_t*>(dev[0]), sl.d_counters, ctx->n_sms, st));
            launches += 2;
        } else {
            RTX_ENQ(launch_quantise_f64(rad_for_quant, static_cast<int64_t>(n_px), p.quantise_mode,
                                        static_cast<uint32_t*>(dev[0]), sl.d_counters, ctx->n_sms, st));
            launches++;
        }
    }
    RTX_ENQ(cudaEventRecord(sl.ev[3], st));
    // Read-back. Host outputs leave on the copy stream, so that the NEXT call's kernels (other slot, other staging
    // buffers) run while this frame crosses PCIe; everything else stays on the one stream.
    sl.copies_on_side_stream = host_out || frame_copy;
    if (frame_copy) {
        // RTX_FRAME_COPY: whole bands / whole frames leave the staging buffer for their place in the frame set (pinned
        // host memory, this GPU's or a peer's HBM) as copy-engine transfers on the copy stream — 2-D copies whose rows are
        // the cyclic bands. The next call's kernel (other slot) overlaps them.
        RTX_ENQ(cudaStreamWaitEvent(cs, sl.ev[3], 0));
        const size_t row_bytes = static_cast<size_t>(W) * 4;
        const char* src = static_cast<const char*>(sl.d_out[0]);
        for (int f = 0; f < n_frames; f++) {
            char* dst = reinterpret_cast<char*>(outs->frame_rgba8) +
                        (static_cast<size_t>(p.frame_offset) + static_cast<size_t>(f) * p.frame_stride) * Hh * row_bytes;
            const char* fsrc = src + static_cast<size_t>(f) * local_rows * row_bytes;
            if (p.n_ranks == 1) {
                RTX_ENQ(cudaMemcpyAsync(dst, fsrc, static_cast<size_t>(Hh) * row_bytes, cudaMemcpyDefault, cs));
            } else {
                const int full_bands = local_rows / band_rows;                       // this rank's complete bands
                const size_t band_bytes = static_cast<size_t>(band_rows) * row_bytes;
                char* first = dst + static_cast<size_t>(p.rank) * band_bytes;        // global band `rank`
                if (full_bands)
                    RTX_ENQ(cudaMemcpy2DAsync(first, band_bytes * p.n_ranks, fsrc, band_bytes, band_bytes, full_bands, cudaMemcpyDefault, cs));
                const int tail_rows = local_rows - full_bands * band_rows;           // ragged last band
                if (tail_rows)
                    RTX_ENQ(cudaMemcpyAsync(first + static_cast<size_t>(full_bands) * band_bytes * p.n_ranks, fsrc + static_cast<size_t>(full_bands) * band_bytes,
                                            static_cast<size_t>(tail_rows) * row_bytes, cudaMemcpyDefault, cs));
            }
        }
    }
    if (host_out || frame_copy) {
        if (!host_out) {
            // planes besides the frame stay where the caller put them (device / mapped): nothing more to copy
        } else if (!ranged) {
            RTX_ENQ(cudaStreamWaitEvent(cs, sl.ev[3], 0));
            for (int k = 0; k < kPlanes; k++)
                if (user[k]) RTX_ENQ(cudaMemcpyAsync(user[k], dev[k], n_px * elem[k], cudaMemcpyDeviceToHost, cs));
        }
        RTX_ENQ(cudaStreamWaitEvent(cs, sl.ev[3], 0));
        RTX_ENQ(cudaMemcpyAsync(sl.h_counters, sl.d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, cs));
        RTX_ENQ(cudaEventRecord(sl.ev[4], cs));
        RTX_ENQ(cudaEventRecord(sl.ev_done, cs));
    } else {
        RTX_ENQ(cudaMemcpyAsync(sl.h_counters, sl.d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        RTX_ENQ(cudaEventRecord(sl.ev[4], st));
        RTX_ENQ(cudaEventRecord(sl.ev_done, st));
    }
    if (hint) {
        // behind everything this call waits for: the next frame's order is built while this frame's pixels leave the GPU
        uint32_t* const now = tile_hist + ctx->order_parity * kOrderKeys;
        uint32_t* const next = tile_hist + (ctx->order_parity ^ 1) * kOrderKeys;
        uint32_t* const cells = tile_hist + 3 * kOrderKeys;
        cudaError_t e = launch_tile_order(tile_cost, static_cast<int>(tiles_full), W, packed_rows, cells, reinterpret_cast<uint16_t*>(cells + order_cells),
                                          now, next, tile_hist + 2 * kOrderKeys, tile_order, st);
        if (e != cudaSuccess) return bail(e, "launch_tile_order");
        ctx->order_parity ^= 1;
        ctx->order_valid = true;
        launches += 3;
    }
#undef RTX_ENQ
    sl.pending = true;
    sl.launches = launches;
    sl.n_spheres = ctx->scene.n_spheres;
    sl.n_walls = ctx->scene.n_walls;
    ctx->error.clear();
    if (async) {
        ctx->next_slot = (ctx->next_slot + 1) % kSlots;
        return RTX_OK;
    }
    // Synchronous call: nothing else is in flight (drained above), and the slot is free again when this returns — so the next
    // call takes the SAME slot and finds its staging buffers already allocated (rotating would allocate kSlots sets one call
    // after the other, the later ones in the middle of somebody's timed region).
    return finish_slot(ctx, sl, stats);
}

}  // namespace

extern "C" {

int rtx_render(rtx_ctx* ctx, const rtx_camera* cams, int32_t n_frames, const rtx_params* params, const rtx_outputs* outs,
               rtx_stats* stats)
{
    if (ctx && !cams) return fail(ctx, RTX_ERR_INVALID, "rtx_render: null argument or n_frames <= 0");
    return render_impl(ctx, cams, n_frames, nullptr, 0, params, outs, stats, false, "rtx_render");
}

int rtx_render_async(rtx_ctx* ctx, const rtx_camera* cams, int32_t n_frames, const rtx_params* params, const rtx_outputs* outs)
{
    if (ctx && !cams) return fail(ctx, RTX_ERR_INVALID, "rtx_render_async: null argument or n_frames <= 0");
    return render_impl(ctx, cams, n_frames, nullptr, 0, params, outs, nullptr, true, "rtx_render_async");
}

int rtx_wait(rtx_ctx* ctx, rtx_stats* stats)
{
    if (!ctx) return RTX_ERR_INVALID;
    Slot& o = ctx->slot[ctx->oldest];
    if (!o.pending) return fail(ctx, RTX_ERR_INVALID, "rtx_wait: no call in flight");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->oldest = (ctx->oldest + 1) % kSlots;
    int rc = finish_slot(ctx, o, stats);
    if (rc == RTX_OK) ctx->error.clear();
    return rc;
}

int rtx_trace_rays(rtx_ctx* ctx, const rtx_ray* rays, int64_t n_rays, const rtx_params* params, const rtx_outputs* outs, rtx_stats* stats)
{
    if (ctx && !rays) return fail(ctx, RTX_ERR_INVALID, "rtx_trace_rays: null rays");
    return render_impl(ctx, nullptr, 0, rays, n_rays, params, outs, stats, false, "rtx_trace_rays");
}

int rtx_quantise(rtx_ctx* ctx, const float* rad32, const double* rad64, int64_t n_pixels, int32_t mode, uint32_t* rgba8,
                 int32_t memory, rtx_stats* stats)
{
    if (!ctx) return RTX_ERR_INVALID;
    if ((rad32 == nullptr) == (rad64 == nullptr)) return fail(ctx, RTX_ERR_INVALID, "rtx_quantise: pass exactly one radiance buffer");
    if (n_pixels < 0 || !rgba8) return fail(ctx, RTX_ERR_INVALID, "rtx_quantise: bad size or null output");
    if (mode != RTX_QUANT_WRAP && mode != RTX_QUANT_SATURATE) return fail(ctx, RTX_ERR_INVALID, "rtx_quantise: unknown mode");
    if (memory != RTX_MEM_HOST && memory != RTX_MEM_DEVICE) return fail(ctx, RTX_ERR_INVALID, "rtx_quantise: bad memory kind");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t in_bytes = static_cast<size_t>(n_pixels) * (rad32 ? 12 : 24);
    const void* d_in = rad32 ? static_cast<const void*>(rad32) : static_cast<const void*>(rad64);
    uint32_t* d_o = rgba8;
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    if (memory == RTX_MEM_HOST) {
        void* pp = ctx->d_rad_scratch;
        int rc = grow(ctx, &pp, &ctx->d_rad_scratch_cap, std::max<size_t>(in_bytes, 16));
        ctx->d_rad_scratch = static_cast<double*>(pp);
        if (rc != RTX_OK) return rc;
        rc = grow(ctx, &ctx->d_aux_out, &ctx->d_aux_out_cap, std::max<size_t>(static_cast<size_t>(n_pixels) * 4, 16));
        if (rc != RTX_OK) return rc;
        RTX_CUDA(ctx, cudaMemcpyAsync(ctx->d_rad_scratch, d_in, in_bytes, cudaMemcpyHostToDevice, st));
        d_in = ctx->d_rad_scratch;
        d_o = static_cast<uint32_t*>(ctx->d_aux_out);
    }
    RTX_CUDA(ctx, cudaMemsetAsync(ctx->d_aux_counters, 0, 8 * sizeof(unsigned long long), st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    if (rad32)
        RTX_CUDA(ctx, launch_quantise_f32(static_cast<const float*>(d_in), n_pixels, mode, d_o, ctx->d_aux_counters, ctx->n_sms, st));
    else
        RTX_CUDA(ctx, launch_quantise_f64(static_cast<const double*>(d_in), n_pixels, mode, d_o, ctx->d_aux_counters, ctx->n_sms, st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    if (memory == RTX_MEM_HOST)
        RTX_CUDA(ctx, cudaMemcpyAsync(rgba8, d_o, static_cast<size_t>(n_pixels) * 4, cudaMemcpyDeviceToHost, st));
    RTX_CUDA(ctx, cudaMemcpyAsync(ctx->h_aux_counters, ctx->d_aux_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    RTX_CUDA(ctx, cudaStreamSynchronize(st));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); stats->h2d_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); stats->surface_update_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); stats->d2h_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); stats->total_ms = ms;
        stats->over_range_pixels = ctx->h_aux_counters[2];
        long long bits = static_cast<long long>(ctx->h_aux_counters[3]);
        std::memcpy(&stats->max_luminance, &bits, sizeof bits);
        stats->launches = 1;
    }
    ctx->error.clear();
    return RTX_OK;
}

int rtx_tonemap(rtx_ctx* ctx, const float* rad32, const double* rad64, int64_t pixels_per_frame, int32_t n_frames,
                const rtx_params* params, uint32_t* rgba8, int32_t memory, double* log_avg_luminance, rtx_stats* stats)
{
    if (!ctx) return RTX_ERR_INVALID;
    if ((rad32 == nullptr) == (rad64 == nullptr)) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap: pass exactly one radiance buffer");
    if (pixels_per_frame <= 0 || n_frames <= 0 || !rgba8 || !params) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap: bad size or null argument");
    const rtx_params& p = *params;
    if (p.tonemap != RTX_TONEMAP_REINHARD) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap: params.tonemap must be RTX_TONEMAP_REINHARD");
    if (!(p.tonemap_key > 0.0)) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap: tonemap_key must be > 0");
    if (p.quantise_mode != RTX_QUANT_WRAP && p.quantise_mode != RTX_QUANT_SATURATE) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap: unknown quantise_mode");
    if (memory != RTX_MEM_HOST && memory != RTX_MEM_DEVICE) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap: bad memory kind");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n_px = static_cast<size_t>(pixels_per_frame) * n_frames;
    const size_t in_bytes = n_px * (rad32 ? 12 : 24);
    const void* d_in = rad32 ? static_cast<const void*>(rad32) : static_cast<const void*>(rad64);
    uint32_t* d_o = rgba8;
    int rc = grow(ctx, &ctx->d_tm_sums, &ctx->d_tm_sums_cap, sizeof(long long) * std::max(n_frames, 16));
    if (rc != RTX_OK) return rc;
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    if (memory == RTX_MEM_HOST) {
        void* pp = ctx->d_rad_scratch;
        rc = grow(ctx, &pp, &ctx->d_rad_scratch_cap, std::max<size_t>(in_bytes, 16));
        ctx->d_rad_scratch = static_cast<double*>(pp);
        if (rc != RTX_OK) return rc;
        rc = grow(ctx, &ctx->d_aux_out, &ctx->d_aux_out_cap, std::max<size_t>(n_px * 4, 16));
        if (rc != RTX_OK) return rc;
        RTX_CUDA(ctx, cudaMemcpyAsync(ctx->d_rad_scratch, d_in, in_bytes, cudaMemcpyHostToDevice, st));
        d_in = ctx->d_rad_scratch;
        d_o = static_cast<uint32_t*>(ctx->d_aux_out);
    }
    RTX_CUDA(ctx, cudaMemsetAsync(ctx->d_aux_counters, 0, 8 * sizeof(unsigned long long), st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    long long* sums = static_cast<long long*>(ctx->d_tm_sums);
    const float* in32 = rad32 ? static_cast<const float*>(d_in) : nullptr;
    const double* in64 = rad32 ? nullptr : static_cast<const double*>(d_in);
    RTX_CUDA(ctx, cudaMemsetAsync(sums, 0, sizeof(long long) * n_frames, st));
    RTX_CUDA(ctx, launch_tonemap_sums(in32, in64, pixels_per_frame, n_frames, sums, ctx->n_sms, st));
    RTX_CUDA(ctx, launch_tonemap_apply(in32, in64, pixels_per_frame, n_frames, sums, pixels_per_frame, p.tonemap_key, p.tonemap_white,
                                       p.quantise_mode, d_o, ctx->d_aux_counters, ctx->n_sms, st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    if (memory == RTX_MEM_HOST) RTX_CUDA(ctx, cudaMemcpyAsync(rgba8, d_o, n_px * 4, cudaMemcpyDeviceToHost, st));
    ctx->h_tm_sums.resize(n_frames);
    RTX_CUDA(ctx, cudaMemcpyAsync(ctx->h_tm_sums.data(), sums, sizeof(long long) * n_frames, cudaMemcpyDeviceToHost, st));
    RTX_CUDA(ctx, cudaMemcpyAsync(ctx->h_aux_counters, ctx->d_aux_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    RTX_CUDA(ctx, cudaStreamSynchronize(st));
    if (log_avg_luminance)   // same expression as the kernel and the oracle: exp((sum / 2^32) / n)
        for (int f = 0; f < n_frames; f++)
            log_avg_luminance[f] = std::exp((static_cast<double>(ctx->h_tm_sums[f]) / 4294967296.0) / static_cast<double>(pixels_per_frame));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); stats->h2d_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); stats->surface_update_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); stats->d2h_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); stats->total_ms = ms;
        stats->over_range_pixels = ctx->h_aux_counters[2];
        long long bits = static_cast<long long>(ctx->h_aux_counters[3]);
        std::memcpy(&stats->max_luminance, &bits, sizeof bits);
        stats->launches = 2;
    }
    ctx->error.clear();
    return RTX_OK;
}

int rtx_tonemap_sums(rtx_ctx* ctx, const float* rad32, const double* rad64, int64_t pixels_per_frame, int32_t n_frames, int64_t* sums)
{
    if (!ctx) return RTX_ERR_INVALID;
    if ((rad32 == nullptr) == (rad64 == nullptr)) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap_sums: pass exactly one radiance buffer");
    if (pixels_per_frame < 0 || n_frames <= 0 || !sums) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap_sums: bad size or null argument");
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, launch_tonemap_sums(rad32, rad64, pixels_per_frame, n_frames, reinterpret_cast<long long*>(sums), ctx->n_sms, ctx->stream));
    RTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->error.clear();
    return RTX_OK;
}

int rtx_tonemap_apply(rtx_ctx* ctx, const float* rad32, const double* rad64, int64_t pixels_per_frame, int32_t n_frames, const int64_t* sums,
                      int64_t pixels_per_frame_global, const rtx_params* params, uint32_t* rgba8, rtx_stats* stats)
{
    if (!ctx) return RTX_ERR_INVALID;
    if ((rad32 == nullptr) == (rad64 == nullptr)) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap_apply: pass exactly one radiance buffer");
    if (pixels_per_frame < 0 || n_frames <= 0 || !sums || !rgba8 || !params || pixels_per_frame_global < pixels_per_frame || pixels_per_frame_global <= 0)
        return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap_apply: bad size or null argument");
    const rtx_params& p = *params;
    if (p.tonemap != RTX_TONEMAP_REINHARD) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap_apply: params.tonemap must be RTX_TONEMAP_REINHARD");
    if (!(p.tonemap_key > 0.0)) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap_apply: tonemap_key must be > 0");
    if (p.quantise_mode != RTX_QUANT_WRAP && p.quantise_mode != RTX_QUANT_SATURATE) return fail(ctx, RTX_ERR_INVALID, "rtx_tonemap_apply: unknown quantise_mode");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    RTX_CUDA(ctx, cudaMemsetAsync(ctx->d_aux_counters, 0, 8 * sizeof(unsigned long long), st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    RTX_CUDA(ctx, launch_tonemap_apply(rad32, rad64, pixels_per_frame, n_frames, reinterpret_cast<const long long*>(sums), pixels_per_frame_global,
                                       p.tonemap_key, p.tonemap_white, p.quantise_mode, rgba8, ctx->d_aux_counters, ctx->n_sms, st));
    RTX_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    RTX_CUDA(ctx, cudaMemcpyAsync(ctx->h_aux_counters, ctx->d_aux_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    RTX_CUDA(ctx, cudaStreamSynchronize(st));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        stats->surface_update_ms = stats->total_ms = ms;
        stats->over_range_pixels = ctx->h_aux_counters[2];
        long long bits = static_cast<long long>(ctx->h_aux_counters[3]);
        std::memcpy(&stats->max_luminance, &bits, sizeof bits);
        stats->launches = 1;
    }
    ctx->error.clear();
    return RTX_OK;
}

int rtx_unpermute_bands(rtx_ctx* ctx, const void* band_major, void* row_major, int32_t height, int32_t width, int32_t elem_bytes,
                        int32_t band_rows, int32_t n_ranks, int32_t rows_per_rank)
{
    if (!ctx) return RTX_ERR_INVALID;
    if (!band_major || !row_major || height <= 0 || width <= 0 || band_rows <= 0 || n_ranks <= 0 || (elem_bytes != 1 && elem_bytes != 4))
        return fail(ctx, RTX_ERR_INVALID, "rtx_unpermute_bands: bad argument");
    for (int r = 0; r < n_ranks; r++)
        if (rtx_local_rows(height, band_rows, n_ranks, r) > rows_per_rank)
            return fail(ctx, RTX_ERR_INVALID, "rtx_unpermute_bands: rows_per_rank smaller than a rank's row count");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, launch_unpermute(band_major, row_major, height, width, elem_bytes, band_rows, n_ranks, rows_per_rank, ctx->n_sms, ctx->stream));
    RTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->error.clear();
    return RTX_OK;
}

int rtx_buffer_alloc(rtx_ctx* ctx, uint64_t bytes, void** device_ptr)
{
    if (!ctx || !device_ptr || bytes == 0) return fail(ctx, RTX_ERR_INVALID, "rtx_buffer_alloc: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(device_ptr, bytes);
    if (e != cudaSuccess) return fail(ctx, RTX_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return RTX_OK;
}

int rtx_buffer_free(rtx_ctx* ctx, void* device_ptr)
{
    if (!ctx) return RTX_ERR_INVALID;
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, cudaFree(device_ptr));
    return RTX_OK;
}

int rtx_buffer_export(rtx_ctx* ctx, void* device_ptr, uint8_t handle[RTX_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == RTX_IPC_HANDLE_BYTES, "IPC handle size");
    if (!ctx || !device_ptr || !handle) return fail(ctx, RTX_ERR_INVALID, "rtx_buffer_export: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    RTX_CUDA(ctx, cudaIpcGetMemHandle(&h, device_ptr));
    std::memcpy(handle, &h, sizeof h);
    return RTX_OK;
}

int rtx_buffer_import(rtx_ctx* ctx, const uint8_t handle[RTX_IPC_HANDLE_BYTES], void** device_ptr)
{
    if (!ctx || !device_ptr || !handle) return fail(ctx, RTX_ERR_INVALID, "rtx_buffer_import: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    RTX_CUDA(ctx, cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RTX_OK;
}

int rtx_buffer_read(rtx_ctx* ctx, const void* device_ptr, void* host_ptr, uint64_t bytes)
{
    if (!ctx || !device_ptr || !host_ptr) return fail(ctx, RTX_ERR_INVALID, "rtx_buffer_read: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, cudaMemcpyAsync(host_ptr, device_ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RTX_OK;
}

int rtx_buffer_release(rtx_ctx* ctx, void* imported_ptr)
{
    if (!ctx) return RTX_ERR_INVALID;
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, cudaIpcCloseMemHandle(imported_ptr));
    return RTX_OK;
}

int rtx_enable_peer_access(rtx_ctx* ctx, int peer_device)
{
    if (!ctx) return RTX_ERR_INVALID;
    if (peer_device == ctx->device) return RTX_OK;
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    int can = 0;
    RTX_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return fail(ctx, RTX_ERR_CUDA, "rtx_enable_peer_access: the devices are not peers");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return RTX_OK;
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
    return RTX_OK;
}

int rtx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rtx_host_alloc(rtx_ctx* ctx, uint64_t bytes, void** host_ptr)
{
    if (!ctx || !host_ptr || bytes == 0) return fail(ctx, RTX_ERR_INVALID, "rtx_host_alloc: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostAlloc(host_ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) return fail(ctx, RTX_ERR_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    return RTX_OK;
}

int rtx_host_free(rtx_ctx* ctx, void* host_ptr)
{
    if (!ctx) return RTX_ERR_INVALID;
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, cudaFreeHost(host_ptr));
    return RTX_OK;
}

int rtx_host_register(rtx_ctx* ctx, void* host_ptr, uint64_t bytes)
{
    if (!ctx || !host_ptr || bytes == 0) return fail(ctx, RTX_ERR_INVALID, "rtx_host_register: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return RTX_OK;
}

int rtx_host_unregister(rtx_ctx* ctx, void* host_ptr)
{
    if (!ctx) return RTX_ERR_INVALID;
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, cudaHostUnregister(host_ptr));
    return RTX_OK;
}

int rtx_host_device_pointer(rtx_ctx* ctx, void* host_ptr, void** device_ptr)
{
    if (!ctx || !host_ptr || !device_ptr) return fail(ctx, RTX_ERR_INVALID, "rtx_host_device_pointer: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, cudaHostGetDevicePointer(device_ptr, host_ptr, 0));
    return RTX_OK;
}

int rtx_host_shared_open(rtx_ctx* ctx, const char* name, uint64_t bytes, int32_t create, void** host_ptr)
{
    if (!ctx || !name || name[0] != '/' || bytes == 0 || !host_ptr) return fail(ctx, RTX_ERR_INVALID, "rtx_host_shared_open: bad argument (name must start with '/')");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    const int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return fail(ctx, RTX_ERR_INVALID, std::string("rtx_host_shared_open: shm_open(") + name + "): " + std::strerror(errno));
    if (create && ftruncate(fd, static_cast<off_t>(bytes)) != 0) {
        const std::string msg = std::string("rtx_host_shared_open: ftruncate: ") + std::strerror(errno);
        close(fd);
        return fail(ctx, RTX_ERR_NOMEM, msg);
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return fail(ctx, RTX_ERR_NOMEM, std::string("rtx_host_shared_open: mmap: ") + std::strerror(errno));
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        munmap(p, bytes);
        return cuda_fail(ctx, e, "cudaHostRegister(shared host frame)");
    }
    ctx->shared_host.emplace_back(p, static_cast<size_t>(bytes));
    *host_ptr = p;
    return RTX_OK;
}

int rtx_host_shared_close(rtx_ctx* ctx, void* host_ptr, const char* unlink_name)
{
    if (!ctx || !host_ptr) return fail(ctx, RTX_ERR_INVALID, "rtx_host_shared_close: bad argument");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    for (size_t k = 0; k < ctx->shared_host.size(); k++) {
        if (ctx->shared_host[k].first != host_ptr) continue;
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copy_stream);
        cudaHostUnregister(host_ptr);
        munmap(host_ptr, ctx->shared_host[k].second);
        ctx->shared_host.erase(ctx->shared_host.begin() + k);
        if (unlink_name) shm_unlink(unlink_name);
        return RTX_OK;
    }
    return fail(ctx, RTX_ERR_INVALID, "rtx_host_shared_close: not a mapping of this context");
}

int rtx_ffma_peak(rtx_ctx* ctx, int32_t variant, double* tflops, double* mhz)
{
    if (!ctx) return RTX_ERR_INVALID;
    if (variant < 0 || variant > 7) return fail(ctx, RTX_ERR_INVALID, "rtx_ffma_peak: variant must be 0..7");
    RTX_CUDA(ctx, cudaSetDevice(ctx->device));
    RTX_CUDA(ctx, run_ffma_peak(variant, ctx->n_sms, ctx->stream, tflops, mhz));
    ctx->error.clear();
    return RTX_OK;
}

}  // extern "C"
