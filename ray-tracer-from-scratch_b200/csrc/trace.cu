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
#include "trace_common.cuh"

namespace rtx {

#ifdef RTX_TAIL_TRACE
// Developer build only (tools/tail_trace.py): what every warp does once the pixel pool is dry — per scan segment the
// start time, the live chains, the mode, and the times after the scan and after drain + shading.
constexpr int kTailRecords = 32;
__device__ unsigned long long g_tail_log[160 * 32 * kTailRecords * 4];
__device__ unsigned long long g_tail_dry;
#endif

// Second FP32 screen + exact evaluation of the queued survivors of one chain. Lanes run this together; each
// iterates over its own queue, so most exact evaluations execute with many lanes active.
__device__ __noinline__ void drain_queue(Chain& c, const SceneDev sc, const float eps)
{
    const int n = c.qn;
    for (int q = 0; q < n; q++) {
        const int e = c.queue[q];
        if (e >= sc.n_entries) continue;                       // padding (only reachable by fallback chains)
        consider_entry(c, e, __ldg(&sc.ent32[e]), sc, eps);
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
#ifndef RTX_SCREEN_ORDER
#define RTX_SCREEN_ORDER 1
#endif
    // The ORDER of these loops does not change any value (every pu / pv / q is the same chain of operations); it only nudges
    // ptxas' schedule and with it how many FFMA2 read five fresh registers (tools/sass_ffma2.py counts them).
#if RTX_SCREEN_ORDER == 0
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
#elif RTX_SCREEN_ORDER == 1
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pu[u] = __ffma2_rn(cy[u], uy, nuo);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cz[u], vz, nvo);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pu[u] = __ffma2_rn(cx[u], ux, pu[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cy[u], vy, pv[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cx[u], vx, pv[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) q[u] = __ffma2_rn(pv[u], pv[u], nw[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) q[u] = __ffma2_rn(pu[u], pu[u], q[u]);
#elif RTX_SCREEN_ORDER == 2
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) {       // entry-major
        pv[u] = __ffma2_rn(cz[u], vz, nvo);
        pu[u] = __ffma2_rn(cy[u], uy, nuo);
        pv[u] = __ffma2_rn(cy[u], vy, pv[u]);
        pu[u] = __ffma2_rn(cx[u], ux, pu[u]);
        pv[u] = __ffma2_rn(cx[u], vx, pv[u]);
        q[u] = __ffma2_rn(pv[u], pv[u], nw[u]);
        q[u] = __ffma2_rn(pu[u], pu[u], q[u]);
    }
#elif RTX_SCREEN_ORDER >= 10 && RTX_SCREEN_ORDER < 20
    // all ten interleavings of the two dependent chains  pv: A1 = cz*vz+nvo, A2 = cy*vy+pv, A3 = cx*vx+pv  and
    // pu: B1 = cy*uy+nuo, B2 = cx*ux+pu  (schedule search: tools/schedule_search.py)
#define RTX_A1 _Pragma("unroll") for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cz[u], vz, nvo);
#define RTX_A2 _Pragma("unroll") for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cy[u], vy, pv[u]);
#define RTX_A3 _Pragma("unroll") for (int u = 0; u < kPairsPerIter; u++) pv[u] = __ffma2_rn(cx[u], vx, pv[u]);
#define RTX_B1 _Pragma("unroll") for (int u = 0; u < kPairsPerIter; u++) pu[u] = __ffma2_rn(cy[u], uy, nuo);
#define RTX_B2 _Pragma("unroll") for (int u = 0; u < kPairsPerIter; u++) pu[u] = __ffma2_rn(cx[u], ux, pu[u]);
#if RTX_SCREEN_ORDER == 10
    RTX_A1 RTX_A2 RTX_A3 RTX_B1 RTX_B2
#elif RTX_SCREEN_ORDER == 11
    RTX_A1 RTX_A2 RTX_B1 RTX_A3 RTX_B2
#elif RTX_SCREEN_ORDER == 12
    RTX_A1 RTX_A2 RTX_B1 RTX_B2 RTX_A3
#elif RTX_SCREEN_ORDER == 13
    RTX_A1 RTX_B1 RTX_A2 RTX_A3 RTX_B2
#elif RTX_SCREEN_ORDER == 14
    RTX_A1 RTX_B1 RTX_A2 RTX_B2 RTX_A3
#elif RTX_SCREEN_ORDER == 15
    RTX_A1 RTX_B1 RTX_B2 RTX_A2 RTX_A3
#elif RTX_SCREEN_ORDER == 16
    RTX_B1 RTX_A1 RTX_A2 RTX_A3 RTX_B2
#elif RTX_SCREEN_ORDER == 17
    RTX_B1 RTX_A1 RTX_A2 RTX_B2 RTX_A3
#elif RTX_SCREEN_ORDER == 18
    RTX_B1 RTX_A1 RTX_B2 RTX_A2 RTX_A3
#elif RTX_SCREEN_ORDER == 19
    RTX_B1 RTX_B2 RTX_A1 RTX_A2 RTX_A3
#endif
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) q[u] = __ffma2_rn(pv[u], pv[u], nw[u]);
#pragma unroll
    for (int u = 0; u < kPairsPerIter; u++) q[u] = __ffma2_rn(pu[u], pu[u], q[u]);
#endif
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

// Out-of-line wrappers for the big kernel (its chains live in local memory; the hot loop should stay small).
__device__ __noinline__ void shade_chain(Chain& c, const TraceArgs& a, FrameTotals& tot) { shade_body<true>(c, a, a.scene, tot); }
__device__ __noinline__ void start_pixel(Chain& c, unsigned long long p, const TraceArgs& a) { start_pixel_body<true>(c, p, a); }

// ---- small scenes ---------------------------------------------------------------------------------------------------
// A handful of objects (the reference's own scene has three) needs no screen: the frame is bound by the double
// shading and, in the big kernel, by instruction-cache misses and local-memory round trips of the chain state
// (ncu on config C2: no_instruction 3.5, long_scoreboard 3.8 warps per issue; profiles/r1_trace_c2_bigkernel_ncu.md).
// This kernel is the same algorithm without the machinery: one chain per lane held in registers, every object tested
// with the exact double routines, persistent lanes refilled from the same pixel counter. Results are identical by
// construction (same sphere_exact / wall_exact / better / shade_body).
// Wall::intersect (scene.cpp:4-35) for the small kernel, where every ray tests every wall exactly and half of those tests
// end at "t > 0" being false: the quotient t = num / den can only be positive if num is neither zero nor NaN, den is not
// NaN, and both carry the same sign bit (+-0 and +-inf included: num / +-0 = +-inf with the product of the signs) — a
// necessary condition, checked on the bit patterns BEFORE the division (a ~25-instruction sequence in double). When it
// holds, t is computed and tested exactly as in wall_exact; when it does not, the reference's own test would have failed
// and -1 is returned as there. Same result for every input.
__device__ __forceinline__ double wall_exact_lazy(d3 o, d3 d, const WallDev& w)
{
    using namespace ex;
    const double denominator = dot(w.n, d);
    const double numerator = dot(sub(w.p, o), w.n);
    const bool same_sign = ((__double2hiint(numerator) ^ __double2hiint(denominator)) >= 0);
    if (!(same_sign && numerator != 0.0 && numerator == numerator && denominator == denominator)) return -1.0;
    const double t = div(numerator, denominator);
    if (t > 0) {
        const d3 rel = sub(add(o, scale(d, t)), w.p);
        const double px = dot(rel, w.right);
        const double py = dot(rel, w.up);
        if (px >= 0 && px <= w.length && py >= 0 && py <= w.width) return t;
    }
    return -1.0;
}

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
                const double t = wall_exact_lazy(c.o, c.d, w);
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

// ---- rebalancing the tail --------------------------------------------------------------------------------------------
// tools/tail_trace.py: after the pixel pool runs dry a 4K frame still holds 0.25 ms of work per SM, yet the
// last warp finishes 1.6 ms later — the chains that are left sit in the warps they started in, and a warp with 40 live
// chains needs ten cooperative passes while its neighbours idle. So ONCE, when every warp of the CTA has found the pool dry,
// the CTA deals its live chains out again: every live chain is written to a per-CTA scratch area in global memory (104 B:
// ray, accumulated colour, weight, pixel, depth budget — the rest is recomputed per segment), and chain i is picked up by
// warp i mod 16. After that every warp holds the same number of chains (±1) and carries on exactly as before (cooperative
// scan, queue, exact tests, shading). Which lane traces a chain has no influence on any result.
#ifndef RTX_TAIL_REBALANCE
#define RTX_TAIL_REBALANCE 1
#endif
constexpr bool kTailRebalance = RTX_TAIL_REBALANCE != 0;      // 0: developer switch (round 1's tail)
struct TailShared {
    int arrived;          // warps of this CTA that have found the pixel pool dry
    int cnt[kWarps];      // live chains per warp at the rebalance
    int pad[3];
};
struct ChainDump {        // what survives between two segments of a chain
    d3 o, d, acc;
    double weight;
    unsigned long long pixel;
    int remaining, first_id, rays, pad;
};
static_assert(sizeof(ChainDump) == 104, "ChainDump layout");
static_assert(static_cast<size_t>(kThreads) * kChains * sizeof(ChainDump) <= kTailScratchBytesPerCta,
              "api.cu allocates kTailScratchBytesPerCta per CTA for the tail rebalance");

__device__ __forceinline__ void dump_chain(ChainDump* dst, const Chain& c)
{
    dst->o = c.o; dst->d = c.d; dst->acc = c.acc;
    dst->weight = c.weight;
    dst->pixel = c.pixel;
    dst->remaining = c.remaining; dst->first_id = c.first_id; dst->rays = c.rays; dst->pad = 0;
}

__device__ __forceinline__ void load_chain(Chain& c, const ChainDump* src)
{
    c.o = src->o; c.d = src->d; c.acc = src->acc;
    c.weight = src->weight;
    c.pixel = src->pixel;
    c.remaining = src->remaining; c.first_id = src->first_id; c.rays = src->rays;
    c.active = 1;
    c.qn = 0;
}

// Every thread of the CTA must call this once (it synchronises the CTA twice).
#ifndef RTX_REBALANCE_INLINE
#define RTX_REBALANCE_INLINE 0
#endif
#if RTX_REBALANCE_INLINE
__device__ __forceinline__ void rebalance_tail(
#else
__device__ __noinline__ void rebalance_tail(
#endif
Chain (&ch)[kChains], TailShared* ts, ChainDump* scratch)
{
    const unsigned lane_id = threadIdx.x & 31u;
    const int warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane_id) - 1u;
    const unsigned m0 = __ballot_sync(kFull, ch[0].active), m1 = __ballot_sync(kFull, ch[1].active);
    if (lane_id == 0) ts->cnt[warp] = __popc(m0) + __popc(m1);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        const int n = ts->cnt[w];
        if (w < warp) base += n;
        total += n;
    }
    if (ch[0].active) dump_chain(scratch + base + __popc(m0 & below), ch[0]);
    if (ch[1].active) dump_chain(scratch + base + __popc(m0) + __popc(m1 & below), ch[1]);
    ch[0].active = ch[1].active = 0;
    __syncthreads();                       // the dumps are visible to the whole CTA
#pragma unroll
    for (int i = 0; i < kChains; i++) {
        const int idx = i * kThreads + static_cast<int>(lane_id) * kWarps + warp;      // chain idx -> warp idx mod 16
        if (idx < total) load_chain(ch[i], scratch + idx);
    }
}

#ifdef RTX_MAXNREG
#define RTX_KERNEL_BOUNDS __maxnreg__(RTX_MAXNREG)
#else
#define RTX_KERNEL_BOUNDS __launch_bounds__(kThreads, 1)
#endif
template <bool STREAM>
__global__ void RTX_KERNEL_BOUNDS trace_kernel(const TraceArgs a, const int tile_pairs, ChainDump* const tail_scratch)
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
        // behind the tile: one mailbox per warp (cooperative drain), then the bookkeeping of the tail rebalance
        if (tail_scratch != nullptr && threadIdx.x == 0)
            reinterpret_cast<TailShared*>(reinterpret_cast<Mailbox*>(s_tile + 2 * plane_pairs) + kWarps)->arrived = 0;
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
    // warp-uniform: 0 = pixels left; 1 = a fetch of this warp found the pixel pool empty; 2 = ... and the warp has told the CTA;
    // 3 = ... and the CTA's chains have been rebalanced (ONE variable on purpose: anything else that lives across the scan
    // changes ptxas' schedule of the hot loop, see tools/sass_ffma2.py)
    int pool_dry = 0;
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
            if (base + n0 + n1 > total_pixels && pool_dry == 0) pool_dry = 1;
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
        int any_active = ch[0].active | ch[1].active;
        if (STREAM) {
            if (!__syncthreads_or(any_active)) break;
        } else {
            bool warp_done = __ballot_sync(kFull, any_active) == 0u;
            if (tail_scratch != nullptr && (pool_dry == 1 || pool_dry == 2)) {
                TailShared* const tail = reinterpret_cast<TailShared*>(mbox - (threadIdx.x >> 5) + kWarps);
                if (pool_dry == 1) {
                    pool_dry = 2;
                    if (lane_id == 0) atomicAdd(&tail->arrived, 1);
                }
                // once EVERY warp of the CTA has found the pool dry, the live chains are dealt out evenly (once); a warp
                // that has run out of chains waits for that, a warp that still has some keeps working until then
                if (warp_done || *reinterpret_cast<volatile int*>(&tail->arrived) == kWarps) {
                    rebalance_tail(ch, tail, tail_scratch + static_cast<size_t>(blockIdx.x) * (kThreads * kChains));
                    pool_dry = 3;
                    any_active = ch[0].active | ch[1].active;
                    warp_done = __ballot_sync(kFull, any_active) == 0u;
                }
            }
            if (warp_done) break;
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
            bool coop = kCoopMax > 0 && pool_dry != 0 && (__popc(m0) + __popc(m1)) <= kCoopMax;
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
    const bool stream_tiles = need + mbox_bytes + sizeof(TailShared) > static_cast<size_t>(kMaxSmemBytes);
    const size_t tile_bytes = stream_tiles ? (static_cast<size_t>(kMaxSmemBytes) - mbox_bytes) / iter_bytes * iter_bytes : (need ? need : iter_bytes);
    const int tile_pairs = static_cast<int>(tile_bytes / 32);
    const size_t smem = tile_bytes + mbox_bytes + (stream_tiles ? 0 : sizeof(TailShared));
    unsigned long long total =
        static_cast<unsigned long long>(args.n_frames) * static_cast<unsigned long long>(args.local_rows) * args.width;
    if (total == 0) return cudaSuccess;
    if (args.use_grid && args.scene.n_entries > kSmallScene) return launch_trace_grid(args, n_sms, stream, launches, state);
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
    // cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (function, device) pair of the PROCESS, not to a context. It is
    // only an upper limit, so every context sets the same maximum once. (Raising it "when needed" per context let a second
    // context with a small scene lower it under the first one's feet: cudaErrorInvalidValue at the big scene's next launch.)
    if (!state->smem_opt_in) {
        cudaError_t err = cudaFuncSetAttribute(trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
        if (err != cudaSuccess) return err;
        state->smem_opt_in = true;
    }
    if (stream_tiles)
        trace_kernel<true><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(args, tile_pairs, nullptr);
    else
        trace_kernel<false><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(args, tile_pairs, kTailRebalance ? static_cast<ChainDump*>(args.tail_scratch) : nullptr);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace rtx
