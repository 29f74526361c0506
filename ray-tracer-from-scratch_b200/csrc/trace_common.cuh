// trace_common.cuh — what the trace kernels share: the exact (double) object tests, the per-chain state, the FP32 screen
// constants of a ray, shading (recursive_ray_tracing's body) and primary-ray generation. Included by trace.cu (brute
// force: trace_kernel, trace_small_kernel) and trace_grid.cu (the uniform-grid extension, rtx_params.accel).
#pragma once
#include "rtx_device.cuh"

namespace rtx {


#ifndef RTX_THREADS
#define RTX_THREADS 512
#endif
#ifndef RTX_PAIRS
#define RTX_PAIRS 6
#endif
constexpr int kThreads = RTX_THREADS; // 16 warps per SM, 4 per scheduler
constexpr int kChains = 2;            // pixels in flight per lane
constexpr int kPairsPerIter = RTX_PAIRS;   // entry pairs per hot-loop iteration (6 pairs = 12 entries x 2 chains = 96 FFMA2)
#ifndef RTX_QUEUE_CAP
#define RTX_QUEUE_CAP 24
#endif
constexpr int kQueue = RTX_QUEUE_CAP; // screen survivors buffered per chain before an early flush
#ifndef RTX_COOP_MAX
#define RTX_COOP_MAX 32
#endif
constexpr int kCoopMax = RTX_COOP_MAX;   // cooperative drain when a warp has at most this many live chains (0 = off)
#ifndef RTX_TILE_ORDER
#define RTX_TILE_ORDER 3      // scheduling hint: bit 0 = slot -> pixel mapping, bit 1 = cost collection; 0 compiles it out (A/B builds)
#endif
#ifndef RTX_MBOX_CAP
#define RTX_MBOX_CAP 96
#endif
constexpr int kMboxCap = RTX_MBOX_CAP;   // cooperative drain: survivors one chain may receive per scan
constexpr int kWarps = kThreads / 32;
struct Mailbox {                      // one per warp, in shared memory behind the entry tile
    int count[4];
    int items[4][kMboxCap];
};
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kMaxSmemBytes = 227 * 1024;

// ---- exact object tests ------------------------------------------------------------------------------

// Sphere::intersect, scene.cpp:40-78. `a` = d.d and `dlen` = |d| are ray constants hoisted by the caller
// (the reference recomputes them per call with the same result). Returns the reference's `distance`
// (projection * |d|, world units; negative when the sphere is behind) or -1 for det < 0.
__device__ __forceinline__ double sphere_exact(d3 o, d3 d, double a, double dlen, SphereExact s, d3* normal)
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

// ---- per-chain state (local memory; the hot loop never touches it) ---------------------------------------------
struct Chain {
    d3 o, d;            // current ray (double, as the reference)
    double a_dd, dlen;  // d.d and |d|
    d3 acc;             // accumulated colour
    double weight;      // product of the metallic factors so far
    double best_dist;   // find_closest_hit state (main.cpp:70)
    unsigned long long pixel;
    float fdx, fdy, fdz, fd_o;   // second-screen constants: d^ and d^.o in float
    float inv_dlen_lo;           // float lower bound of 1/|d| (wall distances are parametric)
    float best_hi;               // float upper bound of best_dist
    int best_key;       // scene id * 8 + face of the best hit so far (-1: none)
    int remaining;      // remaining_iterations (main.cpp:89)
    int first_id;       // primary hit id
    int rays;
    int active;
    int fallback;       // 1: origin outside the error bound's assumption -> every entry goes to the exact test
    int qn;
    int queue[kQueue];  // entry indices that passed the FP32 screen
};

// Screen constants of one chain: two unit axes u, v spanning the plane perpendicular to the ray, and -u.o, -v.o.
// The squared distance from a centre c to the ray's line is (u.c - u.o)^2 + (v.c - v.o)^2.
// u = normalize(d^ x z) has no z component for ANY ray, so u.c costs two FMAs instead of three — uniformly across
// the warp, which is what matters in SIMT (v = d^ x u is general). Rays (anti)parallel to z within 1e-10 take the
// exact fallback.
// Kept as scalars and widened with dup() at each use, so that ptxas emits the FFMA2 operand form that broadcasts
// ONE 32-bit register to both halves (".F32") instead of reading a pair: register-file bandwidth, not the FMA
// pipe, bounds the screen.
struct Packed {
    float ux, uy, nuo, vx, vy, vz, nvo;    // u is chosen perpendicular to the z axis: uz == 0 for every ray
};

__device__ __forceinline__ float2 dup(float v) { return make_float2(v, v); }

// main.cpp:77 generalised to any evaluation order: accept iff distance > 0 and (distance, key) is
// lexicographically smaller than the best so far — identical to the in-order strict '<' scan (key = id * 8 + face).
__device__ __forceinline__ bool better(double dist, int key, double best_dist, int best_key)
{
    return dist > 0 && (dist < best_dist || (dist == best_dist && key < best_key));
}

// Ray constants for both screens. `origin_bound`: the error bound E assumes |o| <= origin_bound; a ray that
// starts farther out (possible through the reference's primary-ray overshoot, main.cpp:99 with |d| > 1) gets
// constants under which EVERY entry passes the screens, i.e. the chain falls back to exact tests of everything.
__device__ __forceinline__ Packed setup_chain(Chain& c, float origin_bound)
{
    Packed k;
    if (!c.active) {
        // idle chain (only while the frame drains): m = -k is huge, nothing passes
        k.ux = k.uy = k.vx = k.vy = k.vz = 0.f;
        k.nuo = k.nvo = 1e15f;          // pu = pv = 1e15: nothing passes
        c.qn = 0;
        c.fallback = 0;
        return k;
    }
    c.a_dd = ex::len2(c.d);
    c.dlen = ex::sqrt(c.a_dd);
    c.best_dist = 1.7976931348623157e308;   // DBL_MAX, main.cpp:70
    c.best_key = -1;
    c.best_hi = __int_as_float(0x7f800000);
    c.qn = 0;
    const double inv = 1.0 / c.dlen;
    const double hx = c.d.x * inv, hy = c.d.y * inv, hz = c.d.z * inv;      // d^ (plain double: not a parity value)
    // u = normalize(d^ x z) = (hy, -hx, 0) / sqrt(hx^2 + hy^2); v = d^ x u
    const double hxy2 = hx * hx + hy * hy;
    const double un = rsqrt(hxy2);
    const double ux = hy * un, uy = -hx * un, uz = 0.0;
    const double vx = hy * uz - hz * uy, vy = hz * ux - hx * uz, vz = hx * uy - hy * ux;
    // the FMAs see the ROUNDED axes: the offsets must be computed from those
    const float fux = static_cast<float>(ux), fuy = static_cast<float>(uy);
    const float fvx = static_cast<float>(vx), fvy = static_cast<float>(vy), fvz = static_cast<float>(vz);
    const float fx = static_cast<float>(hx), fy = static_cast<float>(hy), fz = static_cast<float>(hz);
    c.fdx = fx;
    c.fdy = fy;
    c.fdz = fz;
    c.fd_o = static_cast<float>(static_cast<double>(fx) * c.o.x + static_cast<double>(fy) * c.o.y + static_cast<double>(fz) * c.o.z);
    c.inv_dlen_lo = __double2float_rd(inv) * 0.999999f;
    const double om = fmax(fabs(c.o.x), fmax(fabs(c.o.y), fabs(c.o.z)));
    const bool ok = (om <= static_cast<double>(origin_bound)) && (c.dlen > 0.0) && (c.dlen < 1e300) && (hxy2 > 1e-20);
    c.fallback = ok ? 0 : 1;
    if (ok) {
        k.ux = fux; k.uy = fuy;
        k.vx = fvx; k.vy = fvy; k.vz = fvz;
        k.nuo = -static_cast<float>(static_cast<double>(fux) * c.o.x + static_cast<double>(fuy) * c.o.y);
        k.nvo = -static_cast<float>(static_cast<double>(fvx) * c.o.x + static_cast<double>(fvy) * c.o.y + static_cast<double>(fvz) * c.o.z);
    } else {
        // pu = pv = 0 for every entry -> q - w = -w < 0 -> everything passes; NaN makes the second screen pass too
        k.ux = k.uy = k.vx = k.vy = k.vz = k.nuo = k.nvo = 0.f;
        c.fdx = c.fdy = c.fdz = c.fd_o = __int_as_float(0x7fc00000);
    }
    return k;
}

// One screen survivor `e` with its screen entry s = (cx, cy, cz, r): the second FP32 screen (entirely behind the origin, or
// farther than the best hit so far), then the reference's own test in double and its acceptance rule (main.cpp:77).
__device__ __forceinline__ void consider_entry(Chain& c, int e, const float4 s, const SceneDev& sc, const float eps)
{
    // b32 ~ d^.(c - o). With b* the exact value: |b32 - b*| <= E and the hit lies in [b* - r, b* + r].
    const float b32 = fmaf(c.fdx, s.x, fmaf(c.fdy, s.y, fmaf(c.fdz, s.z, -c.fd_o)));
    const float rb = (s.w + eps) * 1.000001f;
    if (b32 < -rb) return;                                 // entirely behind the origin: never accepted
    if (e < sc.n_spheres) {
        if (b32 - rb > c.best_hi) return;                  // sphere distance is in world units (scene.cpp:77)
        const double dist = sphere_exact(c.o, c.d, c.a_dd, c.dlen, sc.sph64[e], nullptr);
        const int key = sc.sph_key[e];
        if (better(dist, key, c.best_dist, c.best_key)) {
            c.best_dist = dist;
            c.best_key = key;
            c.best_hi = __double2float_ru(dist);
        }
    } else {
        if ((b32 - rb) * c.inv_dlen_lo > c.best_hi) return;     // wall distance is t of the unnormalised d
        const WallDev& w = sc.walls[e - sc.n_spheres];
        const double t = wall_exact(c.o, c.d, w);
        if (better(t, w.key, c.best_dist, c.best_key)) {
            c.best_dist = t;
            c.best_key = w.key;
            c.best_hi = __double2float_ru(t);
        }
    }
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}


struct FrameTotals {
    unsigned long long rays, over;
    double maxlum;
};

// acc += w * v, spelled as explicit fused multiply-adds: the front-to-back accumulation is this repo's own arithmetic (the
// reference folds back to front, main.cpp:117), and written with plain * and + the compiler would be free to contract it
// differently in each kernel that inlines it. Explicit operations make every kernel produce the same bits.
__device__ __forceinline__ void accumulate(d3& acc, double w, d3 v)
{
    acc.x = __fma_rn(w, v.x, acc.x);
    acc.y = __fma_rn(w, v.y, acc.y);
    acc.z = __fma_rn(w, v.z, acc.z);
}

// Shade one finished segment of a chain (recursive_ray_tracing, main.cpp:89-119) and either set up the
// reflected ray or write the pixel.
template <bool ORDER = false>   // ORDER: the kernel takes part in the tile-order scheduling hint (trace_kernel only)
__device__ __forceinline__ void shade_body(Chain& c, const TraceArgs& a, const SceneDev& sc, FrameTotals& tot)
{
    using namespace ex;
    c.rays++;
    const int best_id = c.best_key >> 3;          // object id (-1 stays -1); the low bits are the box face
    const bool primary = c.rays == 1;
    if (primary) {
        c.first_id = best_id;
        if (a.hit_distance) a.hit_distance[c.pixel] = c.best_dist;        // DBL_MAX when nothing was hit (main.cpp:70)
    }
    bool done;
    if (best_id < 0) {
        if (primary && a.hit_normal) {
            a.hit_normal[3 * c.pixel + 0] = 0.0;
            a.hit_normal[3 * c.pixel + 1] = 0.0;
            a.hit_normal[3 * c.pixel + 2] = 0.0;
        }
        // out_color, main.cpp:28-37 (sign test on the unnormalised z)
        d3 col;
        if (c.d.z < 0.0) {
            col = a.ground;
        } else {
            const double vz = div(c.d.z, c.dlen);
            // pow(v.z, 0.25): two correctly rounded square roots are within 1 ulp of it
            const double s = (a.sky_exponent == 0.25) ? sqrt(sqrt(vz)) : pow(vz, a.sky_exponent);
            col = lerp(a.sky_low, a.sky_high, s);
        }
        accumulate(c.acc, c.weight, col);
        done = true;
    } else {
        d3 normal;
        const int slot = sc.slot[best_id];
        if (sc.kind[best_id] == RTX_SPHERE) {
            sphere_exact(c.o, c.d, c.a_dd, c.dlen, sc.sph64[slot], &normal);
        } else {
            normal = sc.walls[slot + (c.best_key & 7)].n;      // a wall, or the face of a box that was hit
        }
        if (primary && a.hit_normal) {
            a.hit_normal[3 * c.pixel + 0] = normal.x;
            a.hit_normal[3 * c.pixel + 1] = normal.y;
            a.hit_normal[3 * c.pixel + 2] = normal.z;
        }
        const MaterialDev m = sc.mats[best_id];
        const d3 pos = add(c.o, scale(c.d, c.best_dist));                  // main.cpp:99
        const d3 ldir = unit(sub(a.light, pos));                           // main.cpp:44,57
        const d3 nn = unit(normal);                                        // main.cpp:46,56
        const d3 dhat = divs(c.d, c.dlen);                                 // normalize(d); normalize(-d) = -dhat
        const double lambert = dot(ldir, nn);                              // main.cpp:46
        const double di = lambert > 0 ? lambert : 0;
        const d3 half = unit(add(neg(dhat), ldir));                        // main.cpp:59
        const double sp = dot(half, nn);                                   // main.cpp:60
        const double si = pow(sp > 0 ? sp : 0, m.exponent);                // main.cpp:103
        const double k = add(add(mul(di, m.diffuse), mul(si, m.specular)), m.ambient);
        d3 local = scale(m.color, k);                                      // main.cpp:104
        if (a.sun_enabled) {
            // EXTENSION (rtx_params.sun_enabled; no reference code, specification = oracle.c::trace): the unused
            // SUN_COLOR / SUN_DIRECTION of main.cpp:18-19 as a directional light through the same Blinn-Phong terms
            const double ls = dot(a.sun_dir, nn);
            const double ds = ls > 0 ? ls : 0;
            const d3 hs = unit(add(neg(dhat), a.sun_dir));
            const double sps = dot(hs, nn);
            const double ss = pow(sps > 0 ? sps : 0, m.exponent);
            const double ks = add(mul(ds, m.diffuse), mul(ss, m.specular));
            const d3 tint = d3{mul(m.color.x, a.sun_color.x), mul(m.color.y, a.sun_color.y), mul(m.color.z, a.sun_color.z)};
            local = add(local, scale(tint, ks));
        }
        if (c.remaining <= 0) {                                            // main.cpp:105-108
            accumulate(c.acc, c.weight, local);
            done = true;
        } else {
            // lerp(local, reflected, metallic) unrolled front to back (main.cpp:117)
            accumulate(c.acc, ex::mul(c.weight, ex::sub(1.0, m.metallic)), local);
            c.weight = ex::mul(c.weight, m.metallic);
            const d3 start = add(pos, scale(normal, a.reflect_offset));    // main.cpp:111 (normal unnormalised)
            const double kk = mul(2.0, dot(dhat, nn));                     // vec.cpp:55
            c.d = sub(dhat, scale(nn, kk));                                // vec.cpp:56
            c.o = start;
            c.remaining--;
            done = false;
        }
    }
    if (done) {
        const unsigned long long p = c.pixel;
        const uint32_t word = (a.rgba8 || a.frame_rgba8) ? pack_rgba(c.acc.x, c.acc.y, c.acc.z, a.quantise_mode) : 0u;
        // finished pixels are write-once: streaming stores (evict-first) keep them from pushing the chains' local-memory
        // lines out of L2 (ncu: 1.2 GB of DRAM writes per 8K frame with plain stores, the frame itself is 0.13 GB)
        if (a.rgba8) __stcs(&a.rgba8[p], word);
        if (a.frame_rgba8) {
            // fused gather: store at the pixel's global position (possibly another GPU's memory, over NVLink)
            const unsigned long long frame_pixels = static_cast<unsigned long long>(a.local_rows) * a.width;
            const int frame = static_cast<int>(p / frame_pixels);
            const unsigned rem = static_cast<unsigned>(p - frame * frame_pixels);
            const int lrow = rem / a.width;
            const int col = rem - lrow * a.width;
            int grow = lrow;
            if (a.n_ranks > 1) {
                const int lb = lrow / a.band_rows;
                grow = (lb * a.n_ranks + a.rank) * a.band_rows + (lrow - lb * a.band_rows);
            }
            const unsigned long long gframe = static_cast<unsigned long long>(a.frame_offset) + static_cast<unsigned long long>(frame) * a.frame_stride;
            __stcs(&a.frame_rgba8[(gframe * a.height + grow) * a.width + col], word);
        }
        if (a.rad64) {
            __stcs(&a.rad64[3 * p + 0], c.acc.x);
            __stcs(&a.rad64[3 * p + 1], c.acc.y);
            __stcs(&a.rad64[3 * p + 2], c.acc.z);
        }
        if (a.rad32) {
            __stcs(&a.rad32[3 * p + 0], static_cast<float>(c.acc.x));
            __stcs(&a.rad32[3 * p + 1], static_cast<float>(c.acc.y));
            __stcs(&a.rad32[3 * p + 2], static_cast<float>(c.acc.z));
        }
        if (a.object_id) a.object_id[p] = c.first_id;
        if (a.hit_mask) a.hit_mask[p] = c.first_id >= 0 ? 1 : 0;
        if (a.ray_count) a.ray_count[p] = static_cast<uint8_t>(c.rays);
        if ((RTX_TILE_ORDER & 2) && ORDER && a.tile_cost) atomicAdd(&a.tile_cost[p >> kOrderTileShift], static_cast<uint32_t>(c.rays));
        tot.rays += static_cast<unsigned long long>(c.rays);
        if (over_range(c.acc.x, c.acc.y, c.acc.z)) tot.over++;
        const double lum = (c.acc.x + c.acc.y + c.acc.z) * (1.0 / 3.0);
        if (lum > tot.maxlum) tot.maxlum = lum;
        c.active = 0;
    }
}

// Primary ray of packed pixel index p (main.cpp:129-134).
template <bool ORDER = false>
__device__ __forceinline__ void start_pixel_body(Chain& c, unsigned long long p, const TraceArgs& a)
{
    using namespace ex;
    if ((RTX_TILE_ORDER & 1) && ORDER && a.tile_order) {
        // scheduling hint: pool slot -> pixel, tiles in descending order of the previous frame's cost (a permutation of the
        // tiles in which a last, partial tile keeps its place, so slots and pixels cover the same range)
        p = (static_cast<unsigned long long>(a.tile_order[p >> kOrderTileShift]) << kOrderTileShift) | (p & ((1ull << kOrderTileShift) - 1ull));
    }
    if (a.rays) {
        // rtx_trace_rays: the caller's ray, as recursive_ray_tracing(scene, ray, depth) receives it (main.cpp:89)
        c.o = mk(a.rays[p].origin);
        c.d = mk(a.rays[p].direction);
    } else {
        const unsigned long long frame_pixels = static_cast<unsigned long long>(a.local_rows) * a.width;
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
        const d3 centre = add(add(mk(cam.image_top_left), scale(mk(cam.delta_x), static_cast<double>(col))),
                              scale(mk(cam.delta_y), static_cast<double>(grow)));   // main.cpp:132
        c.o = mk(cam.position);
        c.d = sub(mk(cam.position), centre);                                           // main.cpp:133
    }
    c.acc = d3{0.0, 0.0, 0.0};
    c.weight = 1.0;
    c.pixel = p;
    c.remaining = a.max_depth;
    c.first_id = -1;
    c.rays = 0;
    c.active = 1;
}


// Frame statistics: warp-shuffle reduction, one atomic per warp (never alters a pixel).
__device__ __forceinline__ void commit_totals(FrameTotals tot, unsigned long long* counters, unsigned lane_id)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        tot.rays += __shfl_down_sync(kFull, tot.rays, off);
        tot.over += __shfl_down_sync(kFull, tot.over, off);
        tot.maxlum = fmax(tot.maxlum, __shfl_down_sync(kFull, tot.maxlum, off));
    }
    if (lane_id == 0) {
        if (tot.rays) atomicAdd(&counters[1], tot.rays);
        if (tot.over) atomicAdd(&counters[2], tot.over);
        // non-negative doubles order like their bit patterns
        if (tot.maxlum > 0.0) atomicMax(&counters[3], static_cast<unsigned long long>(__double_as_longlong(tot.maxlum)));
    }
}

}  // namespace rtx
