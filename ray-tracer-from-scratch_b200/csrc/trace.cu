// trace.cu — the hot path: ray generation, nearest hit over all objects, the reflection chain,
// Blinn-Phong, sky and the 8-bit pack, as ONE persistent sm_100a kernel (trace_kernel; scenes of up to 16 objects
// take the compact trace_small_kernel at the end of this file's device code).
//
// Replaces rt_scene -> recursive_ray_tracing -> find_closest_hit -> SceneGeometry::intersect
// (main.cpp:67-139, scene.cpp:4-78) and, when fused, the quantise loop (main.cpp:338-347).
//
// Design (derivations and measurements in DESIGN.md):
//
//  * Persistent lanes, two chains per lane. One CTA per SM; every lane owns two pixels' reflection chains and
//    fetches a new pixel (warp-aggregated atomic) the moment a chain ends, so the O(N) object scan always runs
//    with live lanes although chains are 1..depth+1 rays long. recursive_ray_tracing's back-to-front lerp
//    (main.cpp:117) becomes a front-to-back accumulation: colour = sum_k W_k (1-m_k) L_k + W_end * T.
//
//  * Exact decisions, FP32 search. The reference is IEEE double. Every (ray, object) pair is first screened by a
//    CONSERVATIVE FP32 test: squared distance from the object's centre to the ray's line, computed in a per-ray
//    frame (u, v perpendicular to the ray): (u.c - u.o)^2 + (v.c - v.o)^2, against (r + E)^2, E bounding the FP32
//    error (filter_eps). The projected form has no |oc|^2 - b^2 cancellation: its error grows with |c|, not |c|^2.
//    Walls take part through their bounding sphere. The screen is 7 packed FFMA2 (fma.rn.f32x2, two entries per
//    instruction) + 2 funnel shifts per two pairs; each broadcast LDS.128 feeds four pairs (2 entries x 2 chains).
//    Survivors (a few per ray) are queued per chain and, after the scan, evaluated with the reference's own
//    double arithmetic, operation for operation, many lanes at a time, and compared with the reference's rule
//    (distance > 0, strictly smaller, lowest scene index on ties). The FP32 stage can only discard pairs the
//    double test would also discard, so ids/distances equal the reference's bit for bit.
//
//  * Objects live in shared memory, pair-interleaved for FFMA2, in two planes: A[p] = (cx_a,cx_b,cy_a,cy_b) and
//    B[p] = (cz_a,cz_b,-w_a,-w_b), w = (r+E)^2 (two planes so that both the broadcast loads of the ordinary scan and
//    the lane-distinct loads of the cooperative drain are conflict-free). 10 064 entries = 161 KB, resident for the
//    whole launch. Larger scenes stream tiles through the same buffer (CTA-synchronous loop).
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

// Second FP32 screen + exact evaluation of the queued survivors of one chain. Lanes run this together; each
// iterates over its own queue, so most exact evaluations execute with many lanes active.
__device__ __noinline__ void drain_queue(Chain& c, const SceneDev sc, const float eps)
{
    const int n = c.qn;
    for (int q = 0; q < n; q++) {
        const int e = c.queue[q];
        if (e >= sc.n_entries) continue;                       // padding (only reachable by fallback chains)
        const float4 s = __ldg(&sc.ent32[e]);                  // (cx, cy, cz, r) of the sphere / bounding sphere
        // b32 ~ d^.(c - o). With b* the exact value: |b32 - b*| <= E and the hit lies in [b* - r, b* + r].
        const float b32 = fmaf(c.fdx, s.x, fmaf(c.fdy, s.y, fmaf(c.fdz, s.z, -c.fd_o)));
        const float rb = (s.w + eps) * 1.000001f;
        if (b32 < -rb) continue;                               // entirely behind the origin: never accepted
        if (e < sc.n_spheres) {
            if (b32 - rb > c.best_hi) continue;                // sphere distance is in world units (scene.cpp:77)
            const double dist = sphere_exact(c.o, c.d, c.a_dd, c.dlen, sc.sph64[e], nullptr);
            const int key = sc.sph_key[e];
            if (better(dist, key, c.best_dist, c.best_key)) {
                c.best_dist = dist;
                c.best_key = key;
                c.best_hi = __double2float_ru(dist);
            }
        } else {
            if ((b32 - rb) * c.inv_dlen_lo > c.best_hi) continue;   // wall distance is t of the unnormalised d
            const WallDev& w = sc.walls[e - sc.n_spheres];
            const double t = wall_exact(c.o, c.d, w);
            if (better(t, w.key, c.best_dist, c.best_key)) {
                c.best_dist = t;
                c.best_key = w.key;
                c.best_hi = __double2float_ru(t);
            }
        }
    }
    c.qn = 0;
}

__device__ __forceinline__ void enqueue(Chain& c, int entry, const SceneDev& sc, float eps)
{
    if (c.qn == kQueue) drain_queue(c, sc, eps);   // rare: early flush
    c.queue[c.qn++] = entry;
}

// 128-bit shared load from a 32-bit shared-window address (keeps address arithmetic out of the loop).
__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// The packed screen of kPairsPerIter entry pairs against one chain; returns the sign history (one bit per entry,
// first entry in the highest of the 2*kPairsPerIter low bits).
//   pu = u.c - u.o (u has no z component),  pv = v.c - v.o,  q - w = pu^2 + pv^2 - w      7 FFMA2 per two entries
// Written constant-major (the same per-ray constant through all pairs before the next constant): the FMA pipe
// accepts one FFMA2 per two cycles only if the instruction reads at most four fresh registers.
__device__ __forceinline__ unsigned screen_pairs(const float2 (&cx)[kPairsPerIter], const float2 (&cy)[kPairsPerIter],
                                                 const float2 (&cz)[kPairsPerIter], const float2 (&nw)[kPairsPerIter],
                                                 const Packed& k)
{
    float2 pu[kPairsPerIter], pv[kPairsPerIter], q[kPairsPerIter];
    const float2 ux = dup(k.ux), uy = dup(k.uy), nuo = dup(k.nuo);
    const float2 vx = dup(k.vx), vy = dup(k.vy), vz = dup(k.vz), nvo = dup(k.nvo);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cz[u], vz, nvo);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pu[u] = __ffma2_rn(cy[u], uy, nuo);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cy[u], vy, pv[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pu[u] = __ffma2_rn(cx[u], ux, pu[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cx[u], vx, pv[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) q[u] = __ffma2_rn(pv[u], pv[u], nw[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) q[u] = __ffma2_rn(pu[u], pu[u], q[u]);
    unsigned h = 0u;
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) {
        h = __funnelshift_l(__float_as_uint(q[u].x), h, 1);
        h = __funnelshift_l(__float_as_uint(q[u].y), h, 1);
    }
    return h;
}

// The O(N) scan over one shared-memory tile: n_pairs entry pairs starting at entry index `base`.
__device__ __forceinline__ void scan_tile(unsigned tile_addr, unsigned plane_bytes, int n_pairs, int base, const Packed& k0,
                                          const Packed& k1, Chain& c0, Chain& c1, const SceneDev& sc, float eps)
{
    unsigned addr;
    asm volatile("mov.u32 %0, %1;" : "=r"(addr) : "r"(tile_addr));   // opaque: keeps the shared base in a register
    int entry = base;
#pragma unroll 1
    for (int it = n_pairs / kPairsPerIter; it > 0; --it, addr += kPairsPerIter * 16u, entry += 2 * kPairsPerIter) {
        float2 cx[kPairsPerIter], cy[kPairsPerIter], cz[kPairsPerIter], nw[kPairsPerIter];
#pragma unroll
        for (int u = 0; u < kPairsPerIter; u++) {
            const float4 p0 = lds128(addr + u * 16u), p1 = lds128(addr + plane_bytes + u * 16u);
            cx[u] = make_float2(p0.x, p0.y);
            cy[u] = make_float2(p0.z, p0.w);
            cz[u] = make_float2(p1.x, p1.y);
            nw[u] = make_float2(p1.z, p1.w);
        }
        unsigned h0 = screen_pairs(cx, cy, cz, nw, k0);   // sign history: bit set = q - w < 0 = survivor
        unsigned h1 = screen_pairs(cx, cy, cz, nw, k1);
        if (h0 | h1) {                                    // about 2e-4 of all pairs
            while (h0) {
                const int bit = 31 - __clz(h0);
                h0 &= ~(1u << bit);
                enqueue(c0, entry + (2 * kPairsPerIter - 1 - bit), sc, eps);
            }
            while (h1) {
                const int bit = 31 - __clz(h1);
                h1 &= ~(1u << bit);
                enqueue(c1, entry + (2 * kPairsPerIter - 1 - bit), sc, eps);
            }
        }
    }
}

// ---- cooperative drain ---------------------------------------------------------------------------------------------
// When the pixel pool is empty the frame is finished by the chains still in flight, and a late chain needs up to
// `depth` more scans: scanning all N entries per lane then leaves a constant tail (measured 3.8 ms on B200,
// DESIGN.md §3.5). In this mode the warp takes its live chains four at a time, broadcasts their screen constants,
// and its 32 lanes split the entry array (lane l screens pairs l, l+32, ...). Survivors are posted to the owner
// through a per-warp shared-memory mailbox and then go through the same queue / exact path as always, so results
// are unchanged (the acceptance rule is order independent).
__device__ __forceinline__ float2 screen_one(const float4 p0, const float4 p1, const Packed& k)
{
    const float2 cx = make_float2(p0.x, p0.y), cy = make_float2(p0.z, p0.w);
    const float2 cz = make_float2(p1.x, p1.y), nw = make_float2(p1.z, p1.w);
    const float2 pu = __ffma2_rn(cx, dup(k.ux), __ffma2_rn(cy, dup(k.uy), dup(k.nuo)));
    const float2 pv = __ffma2_rn(cx, dup(k.vx), __ffma2_rn(cy, dup(k.vy), __ffma2_rn(cz, dup(k.vz), dup(k.nvo))));
    return __ffma2_rn(pu, pu, __ffma2_rn(pv, pv, nw));
}

__device__ __forceinline__ void post(Mailbox* mb, int which, int entry)
{
    const int slot = atomicAdd(&mb->count[which], 1);
    if (slot < kMboxCap) mb->items[which][slot] = entry;
}

__device__ __forceinline__ Packed bcast(const Packed& k, int src)
{
    Packed r;
    r.ux = __shfl_sync(kFull, k.ux, src); r.uy = __shfl_sync(kFull, k.uy, src);
    r.nuo = __shfl_sync(kFull, k.nuo, src);
    r.vx = __shfl_sync(kFull, k.vx, src); r.vy = __shfl_sync(kFull, k.vy, src); r.vz = __shfl_sync(kFull, k.vz, src);
    r.nvo = __shfl_sync(kFull, k.nvo, src);
    return r;
}

constexpr int kCoopChains = 4;   // live chains screened per cooperative pass

__device__ __forceinline__ void coop_scan(unsigned tile_addr, unsigned plane_bytes, int total_pairs,
                                          const Packed (&k)[kCoopChains], Mailbox* mb, unsigned lane)
{
#pragma unroll 1
    for (int pair = lane; pair < total_pairs; pair += 32) {
        const float4 p0 = lds128(tile_addr + pair * 16u), p1 = lds128(tile_addr + plane_bytes + pair * 16u);
        float2 q[kCoopChains];
        unsigned any = 0u;
#pragma unroll
        for (int c = 0; c < kCoopChains; c++) {
            q[c] = screen_one(p0, p1, k[c]);
            any |= __float_as_uint(q[c].x) | __float_as_uint(q[c].y);
        }
        if (any & 0x80000000u) {
#pragma unroll
            for (int c = 0; c < kCoopChains; c++) {
                if (__float_as_uint(q[c].x) & 0x80000000u) post(mb, c, 2 * pair);
                if (__float_as_uint(q[c].y) & 0x80000000u) post(mb, c, 2 * pair + 1);
            }
        }
    }
}

// Pops the next live chain: lowest lane of the first-chain mask, then of the second-chain mask.
__device__ __forceinline__ bool pop_chain(unsigned& r0, unsigned& r1, int& src, int& which)
{
    if (r0) { src = __ffs(r0) - 1; r0 &= r0 - 1; which = 0; return true; }
    if (r1) { src = __ffs(r1) - 1; r1 &= r1 - 1; which = 1; return true; }
    src = 0; which = 0;
    return false;
}

// Cooperative tile fill: entries (cx, cy, cz, r) -> planes A[i] = (cx_a,cx_b,cy_a,cy_b), B[i] = (cz_a,cz_b,-w_a,-w_b),
// w = (r + E)^2, plane B starting plane_pairs float4 after plane A. Padding entries (r < 0) get -w = +1: never pass.
__device__ __forceinline__ void fill_tile(float4* tile, int plane_pairs, const float4* __restrict__ src, int n_pairs, float eps)
{
    for (int i = threadIdx.x; i < n_pairs; i += kThreads) {
        const float4 a = __ldg(&src[2 * i]), b = __ldg(&src[2 * i + 1]);
        const float ra = a.w + eps, rb = b.w + eps;
        const float nwa = a.w >= 0.f ? -(ra * ra) : 1.f;
        const float nwb = b.w >= 0.f ? -(rb * rb) : 1.f;
        tile[i] = make_float4(a.x, b.x, a.y, b.y);
        tile[plane_pairs + i] = make_float4(a.z, b.z, nwa, nwb);
    }
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifdef RTX_TAIL_TRACE
// Developer build only (tools/tail_trace.py): what every warp does once the pixel pool is dry — per scan segment the
// start time, the live chains, the mode, and the times after the scan and after drain + shading.
constexpr int kTailRecords = 32;
__device__ unsigned long long g_tail_log[160 * 32 * kTailRecords * 4];
__device__ unsigned long long g_tail_dry;
#endif

struct FrameTotals {
    unsigned long long rays, over;
    double maxlum;
};

// Shade one finished segment of a chain (recursive_ray_tracing, main.cpp:89-119) and either set up the
// reflected ray or write the pixel.
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
        c.acc.x += c.weight * col.x;
        c.acc.y += c.weight * col.y;
        c.acc.z += c.weight * col.z;
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
            c.acc.x += c.weight * local.x;
            c.acc.y += c.weight * local.y;
            c.acc.z += c.weight * local.z;
            done = true;
        } else {
            // lerp(local, reflected, metallic) unrolled front to back (main.cpp:117)
            const double wl = c.weight * (1.0 - m.metallic);
            c.acc.x += wl * local.x;
            c.acc.y += wl * local.y;
            c.acc.z += wl * local.z;
            c.weight *= m.metallic;
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
        if (a.rgba8) a.rgba8[p] = word;
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
            a.frame_rgba8[(gframe * a.height + grow) * a.width + col] = word;
        }
        if (a.rad64) {
            a.rad64[3 * p + 0] = c.acc.x;
            a.rad64[3 * p + 1] = c.acc.y;
            a.rad64[3 * p + 2] = c.acc.z;
        }
        if (a.rad32) {
            a.rad32[3 * p + 0] = static_cast<float>(c.acc.x);
            a.rad32[3 * p + 1] = static_cast<float>(c.acc.y);
            a.rad32[3 * p + 2] = static_cast<float>(c.acc.z);
        }
        if (a.object_id) a.object_id[p] = c.first_id;
        if (a.hit_mask) a.hit_mask[p] = c.first_id >= 0 ? 1 : 0;
        if (a.ray_count) a.ray_count[p] = static_cast<uint8_t>(c.rays);
        tot.rays += static_cast<unsigned long long>(c.rays);
        if (over_range(c.acc.x, c.acc.y, c.acc.z)) tot.over++;
        const double lum = (c.acc.x + c.acc.y + c.acc.z) * (1.0 / 3.0);
        if (lum > tot.maxlum) tot.maxlum = lum;
        c.active = 0;
    }
}

// Primary ray of packed pixel index p (main.cpp:129-134).
__device__ __forceinline__ void start_pixel_body(Chain& c, unsigned long long p, const TraceArgs& a)
{
    using namespace ex;
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

// Out-of-line wrappers for the big kernel (its chains live in local memory; the hot loop should stay small).
__device__ __noinline__ void shade_chain(Chain& c, const TraceArgs& a, FrameTotals& tot) { shade_body(c, a, a.scene, tot); }
__device__ __noinline__ void start_pixel(Chain& c, unsigned long long p, const TraceArgs& a) { start_pixel_body(c, p, a); }

// ---- small scenes ---------------------------------------------------------------------------------------------------
// A handful of objects (the reference's own scene has three) needs no screen: the frame is bound by the double
// shading and, in the big kernel, by instruction-cache misses and local-memory round trips of the chain state
// (ncu on config C2: no_instruction 3.5, long_scoreboard 3.8 warps per issue; profiles/r1_trace_c2_bigkernel_ncu.md).
// This kernel is the same algorithm without the machinery: one chain per lane held in registers, every object tested
// with the exact double routines, persistent lanes refilled from the same pixel counter. Results are identical by
// construction (same sphere_exact / wall_exact / better / shade_body).
constexpr int kSmallScene = kSmallSceneEntries;   // screen entries (spheres + walls + box faces)
constexpr int kSmallThreads = 256;

#ifndef RTX_SMALL_MINBLOCKS
#define RTX_SMALL_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(kSmallThreads, RTX_SMALL_MINBLOCKS) trace_small_kernel(const TraceArgs a)
{
    // The whole scene (<= 16 entries, < 4 KB) is staged in shared memory once per CTA: every lane reads the same
    // object at the same time, so these are broadcast LDS instead of L1/L2 round trips in front of every exact test
    // and every shading step (ncu: long_scoreboard 1.4 warps per issue with the scene in global memory).
    __shared__ SphereExact s_sph[kSmallScene];
    __shared__ int32_t s_sph_key[kSmallScene];
    __shared__ WallDev s_wall[kSmallScene];
    __shared__ MaterialDev s_mat[kSmallScene];
    __shared__ int32_t s_kind[kSmallScene], s_slot[kSmallScene];
    {
        const SceneDev& g = a.scene;
        for (int i = threadIdx.x; i < g.n_spheres; i += kSmallThreads) { s_sph[i] = g.sph64[i]; s_sph_key[i] = g.sph_key[i]; }
        for (int i = threadIdx.x; i < g.n_walls; i += kSmallThreads) s_wall[i] = g.walls[i];
        for (int i = threadIdx.x; i < g.n_objects; i += kSmallThreads) { s_mat[i] = g.mats[i]; s_kind[i] = g.kind[i]; s_slot[i] = g.slot[i]; }
        __syncthreads();
    }
    SceneDev sc = a.scene;
    sc.sph64 = s_sph;
    sc.sph_key = s_sph_key;
    sc.walls = s_wall;
    sc.mats = s_mat;
    sc.kind = s_kind;
    sc.slot = s_slot;
    const unsigned lane_id = threadIdx.x & 31u;
    const unsigned long long total_pixels = a.pixel_end ? a.pixel_end :
        static_cast<unsigned long long>(a.n_frames) * static_cast<unsigned long long>(a.local_rows) * a.width;
    Chain c;
    c.active = 0;
    c.rays = 0;
    FrameTotals tot{0ull, 0ull, 0.0};
    for (;;) {
        __syncwarp();
        const unsigned idle = __ballot_sync(kFull, !c.active);
        if (idle) {
            unsigned long long base = 0;
            if (lane_id == 0) base = atomicAdd(&a.counters[0], static_cast<unsigned long long>(__popc(idle)));
            base = __shfl_sync(kFull, base, 0);
            if (!c.active) {
                const unsigned long long p = base + __popc(idle & ((1u << lane_id) - 1u));
                if (p < total_pixels) start_pixel_body(c, p, a);
            }
        }
        if (__ballot_sync(kFull, c.active) == 0u) break;
        if (c.active) {
            c.a_dd = ex::len2(c.d);
            c.dlen = ex::sqrt(c.a_dd);
            c.best_dist = 1.7976931348623157e308;   // DBL_MAX, main.cpp:70
            c.best_key = -1;
            for (int i = 0; i < sc.n_spheres; i++) {
                const double dist = sphere_exact(c.o, c.d, c.a_dd, c.dlen, sc.sph64[i], nullptr);
                const int key = sc.sph_key[i];
                if (better(dist, key, c.best_dist, c.best_key)) { c.best_dist = dist; c.best_key = key; }
            }
            for (int i = 0; i < sc.n_walls; i++) {
                const WallDev& w = sc.walls[i];
                const double t = wall_exact(c.o, c.d, w);
                if (better(t, w.key, c.best_dist, c.best_key)) { c.best_dist = t; c.best_key = w.key; }
            }
            shade_body(c, a, sc, tot);
        }
    }
    __syncwarp();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        tot.rays += __shfl_down_sync(kFull, tot.rays, off);
        tot.over += __shfl_down_sync(kFull, tot.over, off);
        tot.maxlum = fmax(tot.maxlum, __shfl_down_sync(kFull, tot.maxlum, off));
    }
    if (lane_id == 0) {
        if (tot.rays) atomicAdd(&a.counters[1], tot.rays);
        if (tot.over) atomicAdd(&a.counters[2], tot.over);
        if (tot.maxlum > 0.0) atomicMax(&a.counters[3], static_cast<unsigned long long>(__double_as_longlong(tot.maxlum)));
    }
}

template <bool STREAM>
__global__ void __launch_bounds__(kThreads, 1) trace_kernel(const TraceArgs a, const int tile_pairs)
{
    extern __shared__ float4 s_tile[];
    const SceneDev& sc = a.scene;
    const unsigned lane_id = threadIdx.x & 31u;
    const unsigned long long total_pixels =
        static_cast<unsigned long long>(a.n_frames) * static_cast<unsigned long long>(a.local_rows) * a.width;
    const int total_pairs = sc.n_entries_padded >> 1;
    const int n_tiles = STREAM ? (total_pairs + tile_pairs - 1) / tile_pairs : 1;
    const unsigned tile_addr = static_cast<unsigned>(__cvta_generic_to_shared(s_tile));
    const int plane_pairs = STREAM ? tile_pairs : total_pairs;          // float4 per plane
    const unsigned plane_bytes = static_cast<unsigned>(plane_pairs) * 16u;

    if (!STREAM) {
        fill_tile(s_tile, plane_pairs, sc.ent32, total_pairs, a.filter_eps);
        __syncthreads();
    }

    Chain ch[kChains];
#pragma unroll
    for (int i = 0; i < kChains; i++) {
        ch[i].active = 0;
        ch[i].qn = 0;
        ch[i].rays = 0;
    }
    FrameTotals tot{0ull, 0ull, 0.0};
    bool pool_dry = false;             // warp-uniform: a fetch of this warp found the pixel pool empty
    Mailbox* const mbox = reinterpret_cast<Mailbox*>(s_tile + 2 * plane_pairs) + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(&a.counters[4], globaltimer_ns());   // kernel start

#ifdef RTX_TAIL_TRACE
    int tail_n = 0;
    unsigned long long* const tail_log = g_tail_log + (static_cast<size_t>(blockIdx.x) * 32 + (threadIdx.x >> 5)) * kTailRecords * 4;
#endif
    for (;;) {
        __syncwarp();
        // ---- refill idle chains with fresh pixels ------------------------------------------------------------------
        const unsigned idle0 = __ballot_sync(kFull, !ch[0].active);
        const unsigned idle1 = __ballot_sync(kFull, !ch[1].active);
        if (idle0 | idle1) {
            const int n0 = __popc(idle0), n1 = __popc(idle1);
            unsigned long long base = 0;
            if (lane_id == 0) base = atomicAdd(&a.counters[0], static_cast<unsigned long long>(n0 + n1));
            base = __shfl_sync(kFull, base, 0);
            const unsigned below = (1u << lane_id) - 1u;
            if (base + n0 + n1 > total_pixels) pool_dry = true;
            if (lane_id == 0 && base + n0 + n1 > total_pixels && base <= total_pixels) {  // this fetch emptied the pool
                atomicMin(&a.counters[5], globaltimer_ns());
#ifdef RTX_TAIL_TRACE
                g_tail_dry = globaltimer_ns();
#endif
            }
            if (!ch[0].active) {
                const unsigned long long p = base + __popc(idle0 & below);
                if (p < total_pixels) start_pixel(ch[0], p, a);
            }
            if (!ch[1].active) {
                const unsigned long long p = base + n0 + __popc(idle1 & below);
                if (p < total_pixels) start_pixel(ch[1], p, a);
            }
        }
        const int any_active = ch[0].active | ch[1].active;
        if (STREAM) {
            if (!__syncthreads_or(any_active)) break;
        } else {
            if (__ballot_sync(kFull, any_active) == 0u) break;
        }

#ifdef RTX_TAIL_TRACE
        const unsigned long long tt0 = globaltimer_ns();
        const int tt_live = __popc(__ballot_sync(kFull, ch[0].active)) + __popc(__ballot_sync(kFull, ch[1].active));
        int tt_coop = 0;
#endif
        // ---- nearest hit (find_closest_hit, main.cpp:67-84): FP32 screen of every entry, exact test of survivors ----
        const Packed k0 = setup_chain(ch[0], a.origin_bound);
        const Packed k1 = setup_chain(ch[1], a.origin_bound);
        if (STREAM) {
            for (int t = 0; t < n_tiles; t++) {
                const int first_pair = t * tile_pairs;
                const int count = min(tile_pairs, total_pairs - first_pair);
                __syncthreads();
                fill_tile(s_tile, plane_pairs, sc.ent32 + 2 * first_pair, count, a.filter_eps);
                __syncthreads();
                scan_tile(tile_addr, plane_bytes, count, 2 * first_pair, k0, k1, ch[0], ch[1], sc, a.filter_eps);
            }
        } else {
            const unsigned m0 = __ballot_sync(kFull, ch[0].active), m1 = __ballot_sync(kFull, ch[1].active);
            bool coop = kCoopMax > 0 && pool_dry && (__popc(m0) + __popc(m1)) <= kCoopMax;
            if (coop) coop = !__any_sync(kFull, (ch[0].active && ch[0].fallback) || (ch[1].active && ch[1].fallback));
            if (coop) {
                unsigned r0 = m0, r1 = m1;
                bool overflow = false;
                while (r0 | r1) {
                    int src[kCoopChains], which[kCoopChains];
                    bool have[kCoopChains];
                    Packed kc[kCoopChains];
#pragma unroll
                    for (int c = 0; c < kCoopChains; c++) {
                        have[c] = pop_chain(r0, r1, src[c], which[c]);
                        kc[c] = bcast(which[c] ? k1 : k0, src[c]);
                        if (!have[c]) {
                            kc[c].ux = kc[c].uy = kc[c].vx = kc[c].vy = kc[c].vz = 0.f;
                            kc[c].nuo = kc[c].nvo = 1e15f;
                        }
                    }
                    if (lane_id < kCoopChains) mbox->count[lane_id] = 0;
                    __syncwarp();
                    coop_scan(tile_addr, plane_bytes, total_pairs, kc, mbox, lane_id);
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < kCoopChains; c++) {
                        if (have[c] && static_cast<int>(lane_id) == src[c]) {
                            const int n = mbox->count[c];
                            if (n > kMboxCap) overflow = true;
                            for (int i = 0; i < min(n, kMboxCap); i++) {
                                if (which[c]) enqueue(ch[1], mbox->items[c][i], sc, a.filter_eps);
                                else enqueue(ch[0], mbox->items[c][i], sc, a.filter_eps);
                            }
                        }
                    }
                    __syncwarp();
                }
                if (__any_sync(kFull, overflow)) {   // a mailbox overflowed: redo this segment the ordinary way
                    ch[0].qn = ch[1].qn = 0;
                    if (ch[0].active) { ch[0].best_dist = 1.7976931348623157e308; ch[0].best_key = -1; ch[0].best_hi = __int_as_float(0x7f800000); }
                    if (ch[1].active) { ch[1].best_dist = 1.7976931348623157e308; ch[1].best_key = -1; ch[1].best_hi = __int_as_float(0x7f800000); }
                    coop = false;
                }
            }
#ifdef RTX_TAIL_TRACE
            tt_coop = coop ? 1 : 0;
#endif
            if (!coop) scan_tile(tile_addr, plane_bytes, total_pairs, 0, k0, k1, ch[0], ch[1], sc, a.filter_eps);
        }
#ifdef RTX_TAIL_TRACE
        __syncwarp();
        const unsigned long long tt1 = globaltimer_ns();
#endif
        drain_queue(ch[0], sc, a.filter_eps);
        drain_queue(ch[1], sc, a.filter_eps);

        // ---- shading, bounce or pixel write --------------------------------------------------------------------------------
        if (ch[0].active) shade_chain(ch[0], a, tot);
        if (ch[1].active) shade_chain(ch[1], a, tot);
#ifdef RTX_TAIL_TRACE
        __syncwarp();
        if (pool_dry && lane_id == 0 && tail_n < kTailRecords) {
            tail_log[4 * tail_n + 0] = tt0;
            tail_log[4 * tail_n + 1] = static_cast<unsigned long long>(tt_live) | (static_cast<unsigned long long>(tt_coop) << 8);
            tail_log[4 * tail_n + 2] = tt1;
            tail_log[4 * tail_n + 3] = globaltimer_ns();
            tail_n++;
        }
#endif
    }

    // ---- diagnostics: warp-shuffle reduction, one atomic per warp (never alters a pixel) ----------------------
    __syncwarp();
    if (lane_id == 0) {   // drain profile: when the first and the last warp ran out of work
        const unsigned long long now = globaltimer_ns();
        atomicMin(&a.counters[6], now);
        atomicMax(&a.counters[7], now);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        tot.rays += __shfl_down_sync(kFull, tot.rays, off);
        tot.over += __shfl_down_sync(kFull, tot.over, off);
        tot.maxlum = fmax(tot.maxlum, __shfl_down_sync(kFull, tot.maxlum, off));
    }
    if (lane_id == 0) {
        if (tot.rays) atomicAdd(&a.counters[1], tot.rays);
        if (tot.over) atomicAdd(&a.counters[2], tot.over);
        // non-negative doubles order like their bit patterns
        if (tot.maxlum > 0.0) atomicMax(&a.counters[3], static_cast<unsigned long long>(__double_as_longlong(tot.maxlum)));
    }
}

#ifdef RTX_TAIL_TRACE
extern "C" int rtx_debug_tail_log(unsigned long long* host_dst, unsigned long long* dry)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(host_dst, g_tail_log, sizeof(unsigned long long) * 160 * 32 * kTailRecords * 4) != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(dry, g_tail_dry, sizeof(unsigned long long)) != cudaSuccess) return 1;
    return 0;
}
extern "C" int rtx_debug_tail_clear(void)
{
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_tail_log) != cudaSuccess) return 1;
    return cudaMemset(p, 0, sizeof(unsigned long long) * 160 * 32 * kTailRecords * 4) != cudaSuccess;
}
#endif

cudaError_t launch_trace(const TraceArgs& args, int n_sms, cudaStream_t stream, int* launches, TraceLaunchState* state)
{
    constexpr int iter_bytes = kPairsPerIter * 32;
    const size_t mbox_bytes = sizeof(Mailbox) * kWarps;
    const size_t need = static_cast<size_t>(args.scene.n_entries_padded) * sizeof(float4);
    const bool stream_tiles = need + mbox_bytes > static_cast<size_t>(kMaxSmemBytes);
    const size_t tile_bytes = stream_tiles ? (static_cast<size_t>(kMaxSmemBytes) - mbox_bytes) / iter_bytes * iter_bytes : (need ? need : iter_bytes);
    const int tile_pairs = static_cast<int>(tile_bytes / 32);
    const size_t smem = tile_bytes + mbox_bytes;
    unsigned long long total =
        static_cast<unsigned long long>(args.n_frames) * static_cast<unsigned long long>(args.local_rows) * args.width;
    if (total == 0) return cudaSuccess;
    if (args.scene.n_entries <= kSmallScene) {
        if (args.pixel_end) total = args.pixel_end - args.pixel_begin;      // a range of the frame (see TraceArgs)
        int& per_sm = state->small_per_sm;
        if (per_sm == 0) {
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_small_kernel, kSmallThreads, 0);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) per_sm = 1;
        }
        unsigned long long small_blocks = (total + kSmallThreads - 1) / kSmallThreads;
        if (small_blocks > static_cast<unsigned long long>(n_sms) * per_sm) small_blocks = static_cast<unsigned long long>(n_sms) * per_sm;
        trace_small_kernel<<<static_cast<unsigned>(small_blocks), kSmallThreads, 0, stream>>>(args);
        if (launches) (*launches)++;
        return cudaGetLastError();
    }
    unsigned long long blocks = (total + kThreads * kChains - 1) / (kThreads * kChains);
    if (blocks > static_cast<unsigned long long>(n_sms)) blocks = n_sms;
    cudaError_t err;
    size_t* const smem_set = state->smem_set;          // the attribute is sticky per function AND device: raise it only when needed
    if (stream_tiles) {
        if (smem > smem_set[1]) {
            err = cudaFuncSetAttribute(trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (err != cudaSuccess) return err;
            smem_set[1] = smem;
        }
        trace_kernel<true><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(args, tile_pairs);
    } else {
        if (smem > smem_set[0]) {
            err = cudaFuncSetAttribute(trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (err != cudaSuccess) return err;
            smem_set[0] = smem;
        }
        trace_kernel<false><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(args, tile_pairs);
    }
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace rtx
