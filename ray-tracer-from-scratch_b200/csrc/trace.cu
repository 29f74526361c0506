// trace.cu — the hot path: ray generation, nearest hit over all objects, the reflection chain,
// Blinn-Phong, sky and the 8-bit pack, as ONE persistent sm_100a kernel.
//
// Replaces rt_scene -> recursive_ray_tracing -> find_closest_hit -> SceneGeometry::intersect
// (main.cpp:67-139, scene.cpp:4-78) and, when fused, the quantise loop (main.cpp:338-347).
//
// Design (see DESIGN.md for the derivations):
//
//  * Persistent lanes. One CTA per SM; each lane owns one pixel's reflection chain at a time and fetches
//    the next pixel (warp-aggregated atomic) the moment its chain ends, so the O(N) object loop always
//    runs with 32 live lanes although chains are 1..depth+1 rays long. recursive_ray_tracing's
//    back-to-front lerp (main.cpp:117) becomes a front-to-back accumulation: colour = sum_k W_k(1-m_k)L_k + W_end*T.
//
//  * Exact decisions, FP32 search. The reference is IEEE double. A ray/sphere pair is first screened by a
//    CONSERVATIVE FP32 test, 10 instructions per pair: the squared distance from the sphere centre to the ray's
//    line, |c x d^ - o x d^|^2, against (r + E)^2, where E bounds the FP32 error (filter_eps). The cross-product
//    form has no |oc|^2 - b^2 cancellation; its error grows with |c|, not |c|^2. Survivors (about 2 per ray in
//    the 10k-sphere scene) pass a second FP32 screen (behind the origin? farther than the best so far?) and are
//    then evaluated with the reference's own double arithmetic, operation for operation, and compared with the
//    reference's rule (distance > 0, strictly smaller, lowest scene index on ties). The FP32 stage can only
//    discard pairs the double test would also discard, so ids/distances equal the reference's bit for bit.
//
//  * Spheres live in shared memory as float4 (cx, cy, cz, (r+E)^2): every lane reads the same sphere, one
//    broadcast LDS.128 per pair. 10 000 spheres = 160 KB, resident for the whole launch. Larger scenes stream
//    tiles through the same buffer (CTA-synchronous loop).
//
//  * Walls are few; each is evaluated in double exactly as Wall::intersect, with the ray-independent basis
//    (scene.cpp:18-19) precomputed at rtx_set_scene.
#include "rtx_device.cuh"

namespace rtx {

constexpr int kThreads = 512;         // 16 warps per SM, 4 per scheduler
constexpr int kUnroll = 8;            // spheres per hot-loop iteration
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kMaxSmemBytes = 227 * 1024;

// ---- exact object tests ------------------------------------------------------------------------------

// Sphere::intersect, scene.cpp:40-78. `a` = d.d and `dlen` = |d| are ray constants hoisted by the caller
// (the reference recomputes them per call with the same result). Returns the reference's `distance`
// (projection * |d|, world units; negative when the sphere is behind) or -1 for det < 0.
__device__ __noinline__ double sphere_exact(d3 o, d3 d, double a, double dlen, SphereExact s, d3* normal)
{
    using namespace ex;
    const d3 c = mk(s.cx, s.cy, s.cz);
    const d3 oc = sub(o, c);
    const double b = mul(2.0, dot(d, oc));
    const double cc = sub(len2(oc), mul(s.r, s.r));
    const double det = sub(mul(b, b), mul(mul(4.0, a), cc));
    if (det < 0) return -1.0;
    double projection;
    d3 point;
    if (det == 0) {
        point = add(o, scale(d, div(-b, mul(2.0, a))));
        projection = div(sub(-b, sqrt(det)), a);            // divides by a, not 2a (scene.cpp:65)
    } else {
        const double sq = sqrt(det);
        const double p1 = div(add(-b, sq), mul(2.0, a));
        const double p2 = div(sub(-b, sq), mul(2.0, a));
        projection = p1 < p2 ? p1 : p2;
        point = add(o, scale(d, projection));
    }
    if (normal) *normal = sub(point, c);                    // unnormalised, length r (scene.cpp:77)
    return mul(projection, dlen);
}

// Wall::intersect, scene.cpp:4-35. Returns t (parametric units of the possibly unnormalised d) or -1.
__device__ __forceinline__ double wall_exact(d3 o, d3 d, const WallDev& w)
{
    using namespace ex;
    const double denominator = dot(w.n, d);
    const double t = div(dot(sub(w.p, o), w.n), denominator);
    if (t > 0) {
        const d3 rel = sub(add(o, scale(d, t)), w.p);
        const double px = dot(rel, w.right);
        const double py = dot(rel, w.up);
        if (px >= 0 && px <= w.length && py >= 0 && py <= w.width) return t;
    }
    return -1.0;
}

// ---- per-ray FP32 screening constants -------------------------------------------------------------------
struct RayF {
    float dx, dy, dz;   // d^ = d/|d| rounded to float
    float kx, ky, kz;   // o x d^ (computed in double with the ROUNDED d^, then rounded)
    float d_o;          // d^ . o
    float best_hi;      // float upper bound of the best distance so far
};

__device__ __forceinline__ RayF make_rayf(d3 o, d3 d, double dlen, float origin_bound)
{
    RayF f;
    const double inv = 1.0 / dlen;
    f.dx = static_cast<float>(d.x * inv);
    f.dy = static_cast<float>(d.y * inv);
    f.dz = static_cast<float>(d.z * inv);
    const double ux = f.dx, uy = f.dy, uz = f.dz;
    f.kx = static_cast<float>(o.y * uz - o.z * uy);
    f.ky = static_cast<float>(o.z * ux - o.x * uz);
    f.kz = static_cast<float>(o.x * uy - o.y * ux);
    f.d_o = static_cast<float>(ux * o.x + uy * o.y + uz * o.z);
    f.best_hi = __int_as_float(0x7f800000);
    // The error bound E assumes |o| <= origin_bound. A ray that starts farther out (possible through the
    // reference's primary-ray overshoot, main.cpp:99 with |d| > 1) poisons its constants with NaN: every
    // FP32 screen then answers "maybe" and the lane falls back to exact tests of all spheres.
    const double om = fmax(fabs(o.x), fmax(fabs(o.y), fabs(o.z)));
    if (!(om <= static_cast<double>(origin_bound)) || !(dlen > 0.0) || !(dlen < 1e300)) {
        f.kx = f.ky = f.kz = f.d_o = __int_as_float(0x7fc00000);
    }
    return f;
}

struct Best {
    double dist;
    int id;    // scene index, -1 = none
};

// main.cpp:77 generalised to any evaluation order: accept iff distance > 0 and (distance, id) is
// lexicographically smaller than the best so far — identical to the in-order strict '<' scan.
__device__ __forceinline__ bool better(double dist, int id, const Best& b)
{
    return dist > 0 && (dist < b.dist || (dist == b.dist && id < b.id));
}

// Second FP32 screen + exact evaluation of one filter survivor.
__device__ __forceinline__ void consider(int j, float4 s, RayF& f, d3 o, d3 d, double a, double dlen,
                                         const SceneDev& sc, Best& best)
{
    if (j >= sc.n_spheres) return;
    // b32 ~ d^.(c - o); the reference distance D satisfies b* - r <= D <= b*, |b32 - b*| <= E, and
    // rb >= r + E, so D >= b32 - rb. b* < 0 means both roots are negative: never accepted.
    const float b32 = fmaf(f.dx, s.x, fmaf(f.dy, s.y, fmaf(f.dz, s.z, -f.d_o)));
    const float rb = sqrtf(s.w) * 1.000001f;
    if (b32 < -rb) return;
    if (b32 - rb > f.best_hi) return;
    const SphereExact e = sc.sph64[j];
    const double dist = sphere_exact(o, d, a, dlen, e, nullptr);
    const int id = sc.sph_id[j];
    if (better(dist, id, best)) {
        best.dist = dist;
        best.id = id;
        f.best_hi = __double2float_ru(dist);
    }
}

// The O(N) loop over one shared-memory tile of spheres: the FP32 screen, kUnroll pairs per iteration.
__device__ __forceinline__ void scan_tile(const float4* __restrict__ tile, int count, int base, RayF& f, d3 o, d3 d,
                                          double a, double dlen, const SceneDev& sc, Best& best)
{
    const float dx = f.dx, dy = f.dy, dz = f.dz;
    const float nkx = -f.kx, nky = -f.ky, nkz = -f.kz;
#pragma unroll 1
    for (int j = 0; j < count; j += kUnroll) {
        float4 s[kUnroll];
        float q[kUnroll];
        bool any = false;
#pragma unroll
        for (int u = 0; u < kUnroll; u++) s[u] = tile[j + u];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            // m = c x d^ - o x d^ ; q = |m|^2 = squared distance from the centre to the ray's line
            const float mx = fmaf(s[u].y, dz, fmaf(-s[u].z, dy, nkx));
            const float my = fmaf(s[u].z, dx, fmaf(-s[u].x, dz, nky));
            const float mz = fmaf(s[u].x, dy, fmaf(-s[u].y, dx, nkz));
            q[u] = fmaf(mx, mx, fmaf(my, my, mz * mz));
            any |= !(q[u] > s[u].w);      // NaN-safe: unordered counts as "maybe"
        }
        if (any) {
#pragma unroll
            for (int u = 0; u < kUnroll; u++)
                if (!(q[u] > s[u].w)) consider(base + j + u, s[u], f, o, d, a, dlen, sc, best);
        }
    }
}

// Cooperative tile fill: (cx, cy, cz, r) -> (cx, cy, cz, (r + E)^2); padding entries (r < 0) can never pass.
__device__ __forceinline__ void fill_tile(float4* tile, const float4* __restrict__ src, int count, float eps)
{
    for (int i = threadIdx.x; i < count; i += kThreads) {
        float4 v = __ldg(&src[i]);
        if (v.w >= 0.f) {
            const float re = v.w + eps;
            v.w = re * re;
        } else {
            v = make_float4(0.f, 0.f, 0.f, -1.f);
        }
        tile[i] = v;
    }
}

struct Lane {
    d3 o, d;            // current ray (double, as the reference)
    d3 acc;             // accumulated colour
    double weight;      // product of the metallic factors so far
    unsigned long long pixel;
    int remaining;      // remaining_iterations (main.cpp:89)
    int first_id;       // primary hit id
    int rays;
    bool active;
};

template <bool STREAM>
__global__ void __launch_bounds__(kThreads, 1) trace_kernel(const TraceArgs a, const int tile_capacity)
{
    extern __shared__ float4 s_tile[];
    const SceneDev& sc = a.scene;
    const unsigned lane_id = threadIdx.x & 31u;
    const unsigned long long total_pixels =
        static_cast<unsigned long long>(a.n_frames) * static_cast<unsigned long long>(a.local_rows) * a.width;
    const unsigned long long frame_pixels = static_cast<unsigned long long>(a.local_rows) * a.width;
    const int n_tiles = STREAM ? (sc.n_spheres_padded + tile_capacity - 1) / tile_capacity : 1;

    if (!STREAM) {
        fill_tile(s_tile, sc.sph32, sc.n_spheres_padded, a.filter_eps);
        __syncthreads();
    }

    Lane L;
    L.active = false;
    L.rays = 0;
    L.pixel = 0;
    unsigned long long my_rays = 0, my_over = 0;
    double my_maxlum = 0.0;

    for (;;) {
        __syncwarp();
        // ---- refill idle lanes with fresh pixels (main.cpp:129-134) ------------------------------------
        const unsigned idle = __ballot_sync(kFull, !L.active);
        if (idle) {
            unsigned long long base = 0;
            const int leader = __ffs(idle) - 1;
            if (static_cast<int>(lane_id) == leader) base = atomicAdd(&a.counters[0], static_cast<unsigned long long>(__popc(idle)));
            base = __shfl_sync(kFull, base, leader);
            if (!L.active) {
                const unsigned long long p = base + __popc(idle & ((1u << lane_id) - 1u));
                if (p < total_pixels) {
                    const int frame = static_cast<int>(p / frame_pixels);
                    const unsigned rem = static_cast<unsigned>(p - frame * frame_pixels);
                    const int lrow = rem / a.width;
                    const int col = rem - lrow * a.width;
                    int grow = lrow;
                    if (a.n_ranks > 1) {
                        const int lb = lrow / a.band_rows;
                        grow = (lb * a.n_ranks + a.rank) * a.band_rows + (lrow - lb * a.band_rows);
                    }
                    const rtx_camera& cam = a.cameras[frame];
                    using namespace ex;
                    const d3 centre = add(add(mk(cam.image_top_left), scale(mk(cam.delta_x), static_cast<double>(col))),
                                          scale(mk(cam.delta_y), static_cast<double>(grow)));   // main.cpp:132
                    L.o = mk(cam.position);
                    L.d = sub(mk(cam.position), centre);                                           // main.cpp:133
                    L.acc = d3{0.0, 0.0, 0.0};
                    L.weight = 1.0;
                    L.pixel = p;
                    L.remaining = a.max_depth;
                    L.first_id = -1;
                    L.rays = 0;
                    L.active = true;
                }
            }
        }
        if (STREAM) {
            if (!__syncthreads_or(L.active ? 1 : 0)) break;
        } else {
            if (__ballot_sync(kFull, L.active) == 0u) break;
        }

        // ---- nearest hit (find_closest_hit, main.cpp:67-84) ------------------------------------------------
        Best best;
        best.dist = 1.7976931348623157e308;   // DBL_MAX, main.cpp:70
        best.id = -1;
        double a_dd = 1.0, dlen = 1.0;
        RayF f;
        if (L.active) {
            a_dd = ex::len2(L.d);
            dlen = ex::sqrt(a_dd);
            for (int w = 0; w < sc.n_walls; w++) {
                const WallDev& wd = sc.walls[w];
                const double t = wall_exact(L.o, L.d, wd);
                if (better(t, wd.id, best)) {
                    best.dist = t;
                    best.id = wd.id;
                }
            }
            f = make_rayf(L.o, L.d, dlen, a.origin_bound);
            f.best_hi = __double2float_ru(best.dist);
        } else {
            // idle lane (only while the frame drains): constants that no sphere can pass
            f.dx = f.dy = f.dz = 0.f;
            f.kx = f.ky = f.kz = 1e15f;
            f.d_o = 0.f;
            f.best_hi = -__int_as_float(0x7f800000);
        }

        if (STREAM) {
            for (int t = 0; t < n_tiles; t++) {
                const int begin = t * tile_capacity;
                const int count = min(tile_capacity, sc.n_spheres_padded - begin);
                __syncthreads();
                fill_tile(s_tile, sc.sph32 + begin, count, a.filter_eps);
                __syncthreads();
                scan_tile(s_tile, count, begin, f, L.o, L.d, a_dd, dlen, sc, best);
            }
        } else {
            scan_tile(s_tile, sc.n_spheres_padded, 0, f, L.o, L.d, a_dd, dlen, sc, best);
        }

        if (!L.active) continue;

        // ---- shade this segment (recursive_ray_tracing, main.cpp:89-119) -----------------------------------
        using namespace ex;
        L.rays++;
        if (L.rays == 1) L.first_id = best.id;
        bool done;
        if (best.id < 0) {
            // out_color, main.cpp:28-37 (sign test on the unnormalised z)
            d3 c;
            if (L.d.z < 0.0) {
                c = a.ground;
            } else {
                const double vz = div(L.d.z, dlen);
                // pow(v.z, 0.25): two correctly rounded square roots are within 1 ulp of it
                const double s = (a.sky_exponent == 0.25) ? sqrt(sqrt(vz)) : pow(vz, a.sky_exponent);
                c = lerp(a.sky_low, a.sky_high, s);
            }
            L.acc.x += L.weight * c.x;
            L.acc.y += L.weight * c.y;
            L.acc.z += L.weight * c.z;
            done = true;
        } else {
            d3 normal;
            const int slot = sc.slot[best.id];
            if (sc.kind[best.id] == RTX_SPHERE) {
                sphere_exact(L.o, L.d, a_dd, dlen, sc.sph64[slot], &normal);
            } else {
                normal = sc.walls[slot].n;
            }
            const MaterialDev m = sc.mats[best.id];
            const d3 pos = add(L.o, scale(L.d, best.dist));                    // main.cpp:99
            const d3 ldir = unit(sub(a.light, pos));                           // main.cpp:44,57
            const d3 nn = unit(normal);                                        // main.cpp:46,56
            const d3 dhat = divs(L.d, dlen);                                   // normalize(d); normalize(-d) = -dhat
            const double lambert = dot(ldir, nn);                              // main.cpp:46
            const double di = lambert > 0 ? lambert : 0;
            const d3 half = unit(add(neg(dhat), ldir));                        // main.cpp:59
            const double sp = dot(half, nn);                                   // main.cpp:60
            const double si = pow(sp > 0 ? sp : 0, m.exponent);                // main.cpp:103
            const double k = add(add(mul(di, m.diffuse), mul(si, m.specular)), m.ambient);
            const d3 local = scale(m.color, k);                                // main.cpp:104
            if (L.remaining <= 0) {                                            // main.cpp:105-108
                L.acc.x += L.weight * local.x;
                L.acc.y += L.weight * local.y;
                L.acc.z += L.weight * local.z;
                done = true;
            } else {
                // lerp(local, reflected, metallic) unrolled front to back (main.cpp:117)
                const double wl = L.weight * (1.0 - m.metallic);
                L.acc.x += wl * local.x;
                L.acc.y += wl * local.y;
                L.acc.z += wl * local.z;
                L.weight *= m.metallic;
                const d3 start = add(pos, scale(normal, a.reflect_offset));    // main.cpp:111 (normal unnormalised)
                const double kk = mul(2.0, dot(dhat, nn));                     // vec.cpp:55
                L.d = sub(dhat, scale(nn, kk));                                // vec.cpp:56
                L.o = start;
                L.remaining--;
                done = false;
            }
        }

        if (done) {
            const unsigned long long p = L.pixel;
            if (a.rgba8) a.rgba8[p] = pack_rgba(L.acc.x, L.acc.y, L.acc.z, a.quantise_mode);
            if (a.rad64) {
                a.rad64[3 * p + 0] = L.acc.x;
                a.rad64[3 * p + 1] = L.acc.y;
                a.rad64[3 * p + 2] = L.acc.z;
            }
            if (a.rad32) {
                a.rad32[3 * p + 0] = static_cast<float>(L.acc.x);
                a.rad32[3 * p + 1] = static_cast<float>(L.acc.y);
                a.rad32[3 * p + 2] = static_cast<float>(L.acc.z);
            }
            if (a.object_id) a.object_id[p] = L.first_id;
            if (a.hit_mask) a.hit_mask[p] = L.first_id >= 0 ? 1 : 0;
            if (a.ray_count) a.ray_count[p] = static_cast<uint8_t>(L.rays);
            my_rays += static_cast<unsigned long long>(L.rays);
            if (over_range(L.acc.x, L.acc.y, L.acc.z)) my_over++;
            const double lum = (L.acc.x + L.acc.y + L.acc.z) * (1.0 / 3.0);
            if (lum > my_maxlum) my_maxlum = lum;
            L.active = false;
        }
    }

    // ---- diagnostics: warp-shuffle reduction, one atomic per warp (never alters a pixel) ----------------------
    __syncwarp();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        my_rays += __shfl_down_sync(kFull, my_rays, off);
        my_over += __shfl_down_sync(kFull, my_over, off);
        my_maxlum = fmax(my_maxlum, __shfl_down_sync(kFull, my_maxlum, off));
    }
    if (lane_id == 0) {
        if (my_rays) atomicAdd(&a.counters[1], my_rays);
        if (my_over) atomicAdd(&a.counters[2], my_over);
        // non-negative doubles order like their bit patterns
        if (my_maxlum > 0.0) atomicMax(&a.counters[3], static_cast<unsigned long long>(__double_as_longlong(my_maxlum)));
    }
}

cudaError_t launch_trace(const TraceArgs& args, int n_sms, cudaStream_t stream, int* launches)
{
    const size_t need = static_cast<size_t>(args.scene.n_spheres_padded) * sizeof(float4);
    const bool stream_tiles = need > static_cast<size_t>(kMaxSmemBytes);
    const size_t smem = stream_tiles ? static_cast<size_t>(kMaxSmemBytes) / (kUnroll * sizeof(float4)) * (kUnroll * sizeof(float4))
                                     : (need ? need : sizeof(float4));
    const int tile_capacity = static_cast<int>(smem / sizeof(float4));
    const unsigned long long total =
        static_cast<unsigned long long>(args.n_frames) * static_cast<unsigned long long>(args.local_rows) * args.width;
    if (total == 0) return cudaSuccess;
    unsigned long long blocks = (total + kThreads - 1) / kThreads;
    if (blocks > static_cast<unsigned long long>(n_sms)) blocks = n_sms;
    cudaError_t err;
    if (stream_tiles) {
        err = cudaFuncSetAttribute(trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (err != cudaSuccess) return err;
        trace_kernel<true><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(args, tile_capacity);
    } else {
        err = cudaFuncSetAttribute(trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (err != cudaSuccess) return err;
        trace_kernel<false><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(args, tile_capacity);
    }
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace rtx
