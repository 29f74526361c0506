// aux_kernels.cu — the HBM-side kernels around the trace kernel:
//
//   quantise_*      the reference's 8-bit pack loop (main.cpp:338-347) as a vectorised streaming kernel
//                   (4 pixels per thread: 128-bit loads, one 128-bit uchar4x4 store) with warp-shuffle
//                   reductions for the two frame statistics (over-range pixel count, max luminance);
//   unpermute_*     scatter of all-gathered cyclic row bands into a row-major frame (multi-GPU epilogue);
//   ffma_peak_*     FP32 FMA throughput microbenchmark — the roofline denominator of the trace kernel,
//                   which MEASURED_PEAKS.json does not carry.
#include "rtx_device.cuh"

namespace rtx {

constexpr unsigned kFullMask = 0xFFFFFFFFu;

struct FrameStats {
    unsigned long long over;
    double maxlum;
};

// One pixel: RGBA8888 word + statistics (over-range count, max luminance) from one set of products.
__device__ __forceinline__ uint32_t quantise_pixel(FrameStats& s, double r, double g, double b, int mode)
{
    bool over;
    const uint32_t word = pack_rgba_flag(r, g, b, mode, over);
    if (over) s.over++;
    const double lum = (r + g + b) * (1.0 / 3.0);
    if (lum > s.maxlum) s.maxlum = lum;
    return word;
}

__device__ __forceinline__ void stats_commit(FrameStats s, unsigned long long* counters)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s.over += __shfl_down_sync(kFullMask, s.over, off);
        s.maxlum = fmax(s.maxlum, __shfl_down_sync(kFullMask, s.maxlum, off));
    }
    if ((threadIdx.x & 31) == 0 && counters) {
        if (s.over) atomicAdd(&counters[2], s.over);
        if (s.maxlum > 0.0) atomicMax(&counters[3], static_cast<unsigned long long>(__double_as_longlong(s.maxlum)));
    }
}

// float radiance: 4 pixels = 12 floats = three float4 loads -> one uint4 store.
__global__ void __launch_bounds__(256) quantise_f32_kernel(const float* __restrict__ rad, long long n_pixels, int mode,
                                                           uint32_t* __restrict__ out, unsigned long long* counters)
{
    FrameStats st{0ull, 0.0};
    const long long n_quads = n_pixels >> 2;
    const float4* __restrict__ in4 = reinterpret_cast<const float4*>(rad);
    uint4* __restrict__ out4 = reinterpret_cast<uint4*>(out);
    for (long long qd = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; qd < n_quads;
         qd += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 a = __ldcs(&in4[3 * qd + 0]);
        const float4 b = __ldcs(&in4[3 * qd + 1]);
        const float4 c = __ldcs(&in4[3 * qd + 2]);
        uint4 o;
        o.x = quantise_pixel(st, a.x, a.y, a.z, mode);
        o.y = quantise_pixel(st, a.w, b.x, b.y, mode);
        o.z = quantise_pixel(st, b.z, b.w, c.x, mode);
        o.w = quantise_pixel(st, c.y, c.z, c.w, mode);
        __stcs(&out4[qd], o);
    }
    // ragged tail (n_pixels not a multiple of 4)
    if (blockIdx.x == 0 && threadIdx.x < (n_pixels & 3)) {
        const long long p = (n_quads << 2) + threadIdx.x;
        const double r = rad[3 * p], g = rad[3 * p + 1], b = rad[3 * p + 2];
        out[p] = quantise_pixel(st, r, g, b, mode);
    }
    stats_commit(st, counters);
}

// double radiance: 4 pixels = 12 doubles = six double2 loads -> one uint4 store.
__global__ void __launch_bounds__(256) quantise_f64_kernel(const double* __restrict__ rad, long long n_pixels, int mode,
                                                           uint32_t* __restrict__ out, unsigned long long* counters)
{
    FrameStats st{0ull, 0.0};
    const long long n_quads = n_pixels >> 2;
    const double2* __restrict__ in2 = reinterpret_cast<const double2*>(rad);
    uint4* __restrict__ out4 = reinterpret_cast<uint4*>(out);
    for (long long qd = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; qd < n_quads;
         qd += static_cast<long long>(gridDim.x) * blockDim.x) {
        double2 v[6];
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = __ldcs(&in2[6 * qd + k]);
        uint4 o;
        o.x = quantise_pixel(st, v[0].x, v[0].y, v[1].x, mode);
        o.y = quantise_pixel(st, v[1].y, v[2].x, v[2].y, mode);
        o.z = quantise_pixel(st, v[3].x, v[3].y, v[4].x, mode);
        o.w = quantise_pixel(st, v[4].y, v[5].x, v[5].y, mode);
        __stcs(&out4[qd], o);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n_pixels & 3)) {
        const long long p = (n_quads << 2) + threadIdx.x;
        const double r = rad[3 * p], g = rad[3 * p + 1], b = rad[3 * p + 2];
        out[p] = quantise_pixel(st, r, g, b, mode);
    }
    stats_commit(st, counters);
}

// Any 4- / 8-byte aligned pointers (tensor slices, odd frame offsets): one pixel per thread, scalar loads and stores.
template <typename T>
__global__ void __launch_bounds__(256) quantise_scalar_kernel(const T* __restrict__ rad, long long n_pixels, int mode,
                                                              uint32_t* __restrict__ out, unsigned long long* counters)
{
    FrameStats st{0ull, 0.0};
    for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n_pixels;
         p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double r = rad[3 * p], g = rad[3 * p + 1], b = rad[3 * p + 2];
        out[p] = quantise_pixel(st, r, g, b, mode);
    }
    stats_commit(st, counters);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int stream_grid(long long work_items, int threads, int n_sms)
{
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = static_cast<long long>(n_sms) * 8;   // 8 x 256 threads = full occupancy, whole waves
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

cudaError_t launch_quantise_f64(const double* rad, int64_t n_pixels, int mode, uint32_t* rgba8,
                                unsigned long long* counters, int n_sms, cudaStream_t stream)
{
    if (n_pixels <= 0) return cudaSuccess;
    if (!aligned16(rad) || !aligned16(rgba8))
        quantise_scalar_kernel<double><<<stream_grid(n_pixels, 256, n_sms), 256, 0, stream>>>(rad, n_pixels, mode, rgba8, counters);
    else
    quantise_f64_kernel<<<stream_grid((n_pixels + 3) / 4, 256, n_sms), 256, 0, stream>>>(rad, n_pixels, mode, rgba8, counters);
    return cudaGetLastError();
}

cudaError_t launch_quantise_f32(const float* rad, int64_t n_pixels, int mode, uint32_t* rgba8,
                                unsigned long long* counters, int n_sms, cudaStream_t stream)
{
    if (n_pixels <= 0) return cudaSuccess;
    if (!aligned16(rad) || !aligned16(rgba8))
        quantise_scalar_kernel<float><<<stream_grid(n_pixels, 256, n_sms), 256, 0, stream>>>(rad, n_pixels, mode, rgba8, counters);
    else
    quantise_f32_kernel<<<stream_grid((n_pixels + 3) / 4, 256, n_sms), 256, 0, stream>>>(rad, n_pixels, mode, rgba8, counters);
    return cudaGetLastError();
}

// ---- small uploads without the copy engine ---------------------------------------------------------------------------
// Scene blob, cameras and the counter reset of a call are a few hundred bytes. As cudaMemcpyAsync they queue on a copy
// engine — behind the 8 MB read-back of the PREVIOUS frame when frames are streamed (rtx_render_async): measured 96 us of
// stall per 1080p frame in front of an 80 us kernel. A one-CTA kernel that reads the pinned, mapped staging buffer over
// PCIe has no such queue.
__global__ void __launch_bounds__(256) small_upload_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src_host, int n16)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src_host[i];
}

cudaError_t launch_small_upload(void* dst, const void* src_host_mapped, size_t bytes, cudaStream_t stream)
{
    const int n16 = static_cast<int>((bytes + 15) / 16);       // both buffers are allocated in multiples of 16 bytes
    if (n16 <= 0) return cudaSuccess;
    const int blocks = n16 > 256 * 8 ? 8 : (n16 + 255) / 256;
    small_upload_kernel<<<blocks, 256, 0, stream>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src_host_mapped), n16);
    return cudaGetLastError();
}

// counters of a call: [0] = first pixel of the launch, [1..3] = 0, [4..6] = ~0 (atomicMin slots), [7] = 0
__global__ void reset_counters_kernel(unsigned long long* counters, unsigned long long first_pixel, int all)
{
    const int k = threadIdx.x;
    if (k == 0) counters[0] = first_pixel;
    else if (all && k < 8) counters[k] = (k >= 4 && k <= 6) ? ~0ull : 0ull;
}

cudaError_t launch_reset_counters(unsigned long long* counters, unsigned long long first_pixel, bool all, cudaStream_t stream)
{
    reset_counters_kernel<<<1, 8, 0, stream>>>(counters, first_pixel, all ? 1 : 0);
    return cudaGetLastError();
}

// ---- scheduling hint: the next frame's tile order ---------------------------------------------------------------------
// trace_kernel hands out pixels through a pool; the frame ends when the chains still in flight at the moment the pool runs
// dry are finished. If the LAST pixels handed out are cheap ones (sky: one ray) that tail is short, so the
// pool serves tiles of 256 pixels in descending order of what they cost in the previous frame — longest processing time
// first. No result depends on the order. The kernel adds every finished pixel's ray count to tile_cost; three small
// launches turn the costs into a permutation and clear them for the next frame:
//   1. cells: the largest tile cost in every cell of a coarse grid over the packed frame (256 pixels x 32 rows);
//   2. key of a tile = its own cost class (1/8 ray per pixel) x 4 + how expensive its surroundings are (the cells up to 3
//      above/below and 2 to either side: >= 96 rows, >= 512 pixels) — among equally cheap tiles the ones a moving or turning
//      camera cannot have filled with objects go last; histogram of the keys;
//   3. counting sort, highest key first (the order inside a key is whatever the atomics give); costs and cells cleared.
// n = the number of COMPLETE tiles: a last partial tile keeps the last place, so that pool slots and pixels cover the same
// range.
__device__ __forceinline__ uint32_t cost_class(uint32_t cost)
{
    const uint32_t b = cost >> 5;                       // 256 pixels x (1 .. depth + 1) rays
    return b < kOrderClasses ? b : kOrderClasses - 1;
}

struct TileGrid {
    int width;       // pixels per packed row
    int cells_x;     // ceil(width / 256)
    int cells_y;     // ceil(packed rows / 32)
};

__device__ __forceinline__ void tile_cell(const TileGrid& g, int tile, int& cx, int& cy)
{
    const long long q = (static_cast<long long>(tile) << kOrderTileShift) + (1 << (kOrderTileShift - 1));   // the tile's middle pixel
    const long long row = q / g.width;
    cx = static_cast<int>(q - row * g.width) >> 8;
    cy = static_cast<int>(row >> 5);
    if (cy >= g.cells_y) cy = g.cells_y - 1;
}

__global__ void __launch_bounds__(1024) tile_cells_kernel(const uint32_t* __restrict__ cost, int n, TileGrid g, uint32_t* cells)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx, cy;
    tile_cell(g, i, cx, cy);
    atomicMax(&cells[cy * g.cells_x + cx], cost[i]);
}

__global__ void __launch_bounds__(1024) tile_key_kernel(const uint32_t* __restrict__ cost, int n, TileGrid g,
                                                        const uint32_t* __restrict__ cells, uint16_t* __restrict__ key,
                                                        uint32_t* hist, uint32_t* fill)
{
    __shared__ uint32_t sh[kOrderKeys];
    for (int k = threadIdx.x; k < kOrderKeys; k += blockDim.x) sh[k] = 0u;
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < kOrderKeys; k += blockDim.x) fill[k] = 0u;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int cx, cy;
        tile_cell(g, i, cx, cy);
        uint32_t around = 0u;
        for (int y = max(cy - 3, 0); y <= min(cy + 3, g.cells_y - 1); y++)
            for (int x = max(cx - 2, 0); x <= min(cx + 2, g.cells_x - 1); x++) around = max(around, cells[y * g.cells_x + x]);
        // surroundings: 0 = nothing but single-ray pixels (up to 1/8 extra ray per pixel), 1 = up to 2 rays per pixel, 2 = up to 4, 3 = more
        const uint32_t risk = around < 288u ? 0u : around < 512u ? 1u : around < 1024u ? 2u : 3u;
        const uint32_t k = cost_class(cost[i]) * 4u + risk;
        key[i] = static_cast<uint16_t>(k);
        atomicAdd(&sh[k], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kOrderKeys; k += blockDim.x)
        if (sh[k]) atomicAdd(&hist[k], sh[k]);
}

__global__ void __launch_bounds__(1024) tile_scatter_kernel(uint32_t* __restrict__ cost, int n, const uint16_t* __restrict__ key,
                                                            const uint32_t* __restrict__ hist, uint32_t* hist_next, uint32_t* fill,
                                                            uint32_t* __restrict__ order, uint32_t* cells, int n_cells)
{
    __shared__ uint32_t start[kOrderKeys], cnt[kOrderKeys], base[kOrderKeys];
    const int t = threadIdx.x;
    for (int k = t; k < kOrderKeys; k += blockDim.x) {
        base[k] = hist[k];
        cnt[k] = 0u;
        if (blockIdx.x == 0) hist_next[k] = 0u;                        // the other histogram of the pair, for the next frame
    }
    for (int c = blockIdx.x * blockDim.x + t; c < n_cells; c += gridDim.x * blockDim.x) cells[c] = 0u;
    __syncthreads();
    for (int k = t; k < kOrderKeys; k += blockDim.x) {
        uint32_t s = 0u;
        for (int b = k + 1; b < kOrderKeys; b++) s += base[b];         // descending: the keys above this one come first
        start[k] = s;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + t;
    int b = 0;
    uint32_t rank = 0u;
    if (i < n) {
        b = key[i];
        rank = atomicAdd(&cnt[b], 1u);
        cost[i] = 0u;
    }
    __syncthreads();
    for (int k = t; k < kOrderKeys; k += blockDim.x) base[k] = cnt[k] ? atomicAdd(&fill[k], cnt[k]) : 0u;
    __syncthreads();
    if (i < n) order[start[b] + base[b] + rank] = static_cast<uint32_t>(i);
}

__global__ void __launch_bounds__(256) tile_identity_kernel(uint32_t* order, uint32_t* cost, int n_all, uint32_t* hist_fill_cells, int n_aux)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_all) {
        order[i] = static_cast<uint32_t>(i);
        cost[i] = 0u;
    }
    if (i < n_aux) hist_fill_cells[i] = 0u;              // both histograms, the fill counters and the cells
}

static TileGrid tile_grid(int width, long long packed_rows)
{
    TileGrid g;
    g.width = width;
    g.cells_x = (width + 255) / 256;
    g.cells_y = static_cast<int>((packed_rows + 31) / 32);
    if (g.cells_y < 1) g.cells_y = 1;
    return g;
}

size_t tile_order_cells(int width, long long packed_rows)
{
    const TileGrid g = tile_grid(width, packed_rows);
    return static_cast<size_t>(g.cells_x) * g.cells_y;
}

cudaError_t launch_tile_reset(uint32_t* tile_order, uint32_t* tile_cost, int n_tiles_all, uint32_t* hist_fill_cells, int n_aux, cudaStream_t stream)
{
    const int n = n_tiles_all > n_aux ? n_tiles_all : n_aux;
    tile_identity_kernel<<<(n + 255) / 256, 256, 0, stream>>>(tile_order, tile_cost, n_tiles_all, hist_fill_cells, n_aux);
    return cudaGetLastError();
}

cudaError_t launch_tile_order(uint32_t* tile_cost, int n_tiles, int width, long long packed_rows, uint32_t* cells, uint16_t* tile_key,
                              uint32_t* hist_now, uint32_t* hist_next, uint32_t* fill, uint32_t* tile_order, cudaStream_t stream)
{
    if (n_tiles <= 0) return cudaSuccess;
    const TileGrid g = tile_grid(width, packed_rows);
    const int blocks = (n_tiles + 1023) / 1024;
    tile_cells_kernel<<<blocks, 1024, 0, stream>>>(tile_cost, n_tiles, g, cells);
    tile_key_kernel<<<blocks, 1024, 0, stream>>>(tile_cost, n_tiles, g, cells, tile_key, hist_now, fill);
    tile_scatter_kernel<<<blocks, 1024, 0, stream>>>(tile_cost, n_tiles, tile_key, hist_now, hist_next, fill, tile_order, cells, g.cells_x * g.cells_y);
    return cudaGetLastError();
}

// ---- multi-GPU epilogue ----------------------------------------------------------------------------------
// dst row i lives in band b = i / band_rows, owned by rank b % n_ranks, at packed local row
// (b / n_ranks) * band_rows + i % band_rows of that rank's block.
template <typename T>
__global__ void __launch_bounds__(256) unpermute_kernel(const T* __restrict__ src, T* __restrict__ dst, int height,
                                                        int row_elems, int band_rows, int n_ranks, int rows_per_rank)
{
    const long long total = static_cast<long long>(height) * row_elems;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(e / row_elems);
        const int x = static_cast<int>(e - static_cast<long long>(i) * row_elems);
        const int band = i / band_rows;
        const int r = band % n_ranks;
        const int lrow = (band / n_ranks) * band_rows + (i - band * band_rows);
        dst[e] = src[(static_cast<long long>(r) * rows_per_rank + lrow) * row_elems + x];
    }
}

cudaError_t launch_unpermute(const void* band_major, void* row_major, int height, int width, int elem_bytes,
                             int band_rows, int n_ranks, int rows_per_rank, int n_sms, cudaStream_t stream)
{
    const long long row_bytes = static_cast<long long>(width) * elem_bytes;
    const bool vec16 = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(band_major) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(row_major) % 16 == 0);
    if (vec16) {
        const int row_elems = static_cast<int>(row_bytes / 16);
        const long long total = static_cast<long long>(height) * row_elems;
        unpermute_kernel<uint4><<<stream_grid(total, 256, n_sms), 256, 0, stream>>>(
            static_cast<const uint4*>(band_major), static_cast<uint4*>(row_major), height, row_elems, band_rows, n_ranks,
            rows_per_rank);
    } else if (elem_bytes == 4) {
        const long long total = static_cast<long long>(height) * width;
        unpermute_kernel<uint32_t><<<stream_grid(total, 256, n_sms), 256, 0, stream>>>(
            static_cast<const uint32_t*>(band_major), static_cast<uint32_t*>(row_major), height, width, band_rows, n_ranks,
            rows_per_rank);
    } else {
        const long long total = static_cast<long long>(height) * width;
        unpermute_kernel<uint8_t><<<stream_grid(total, 256, n_sms), 256, 0, stream>>>(
            static_cast<const uint8_t*>(band_major), static_cast<uint8_t*>(row_major), height, width, band_rows, n_ranks,
            rows_per_rank);
    }
    return cudaGetLastError();
}

// ---- FP32 peak microbenchmark --------------------------------------------------------------------------------
constexpr int kPeakChains = 16;
constexpr int kPeakIters = 4096;

__global__ void __launch_bounds__(512) ffma_peak_scalar(float* sink, float a, float b, long long* clocks)
{
    float acc[kPeakChains];
#pragma unroll
    for (int k = 0; k < kPeakChains; k++) acc[k] = threadIdx.x * 1e-3f + k;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int k = 0; k < kPeakChains; k++) acc[k] = fmaf(acc[k], a, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kPeakChains; k++) s += acc[k];
    if (s == 123.456f) sink[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0;
}

__global__ void __launch_bounds__(512) ffma_peak_packed(float* sink, float a, float b, long long* clocks)
{
    unsigned long long acc[kPeakChains / 2];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int k = 0; k < kPeakChains / 2; k++) {
        const float lo = threadIdx.x * 1e-3f + k, hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[k]) : "f"(lo), "f"(hi));
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int k = 0; k < kPeakChains / 2; k++)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(av), "l"(bv));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kPeakChains / 2; k++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
        s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0;
}

// Register-file pressure variants of the packed FMA: what limits the trace kernel's screen is not the FMA pipe
// but how many distinct registers an FFMA2 reads (three 64-bit operands = 6 registers over 2 issue cycles).
//   variant 2: d_k = fma2(a_k, b_k, d_k)   three distinct pairs per instruction, nothing shared
//   variant 3: d_k = fma2(a_k, B,   d_k)   one pair shared by consecutive instructions (operand reuse cache)
//   variant 4: d_k = fma2(a_k, a_k, d_k)   two distinct pairs
template <int MODE>
__global__ void __launch_bounds__(512) ffma_peak_operands(float* sink, float a, float b, long long* clocks)
{
    constexpr int N = 8;
    unsigned long long acc[N], x[N], y[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        const float lo = threadIdx.x * 1e-3f + k, hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[k]) : "f"(lo), "f"(hi));
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[k]) : "f"(a + k * 1e-9f), "f"(a - k * 1e-9f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(y[k]) : "f"(b + k * 1e-9f), "f"(b - k * 1e-9f));
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k]) : "l"(x[k]), "l"(y[k]));
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k]) : "l"(x[k]), "l"(y[0]));
            if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc[k]) : "l"(x[k]));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < N; k++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
        s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0;
}

// variant 5: packed and scalar FMAs interleaved (8 FFMA2 + NS scalar FFMA per iteration, all operands in the
// full-rate forms): is there FP32 capacity beyond what FFMA2 alone reaches (a second pipe for scalar FFMA)?
template <int NS>
__global__ void __launch_bounds__(512) ffma_peak_mixed(float* sink, float a, float b, long long* clocks)
{
    unsigned long long acc[8];
    float sc[NS > 0 ? NS : 1];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float lo = threadIdx.x * 1e-3f + k, hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[k]) : "f"(lo), "f"(hi));
    }
#pragma unroll
    for (int k = 0; k < NS; k++) sc[k] = threadIdx.x * 2e-3f + k;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(av), "l"(bv));
            if (k < NS) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(sc[k]) : "f"(a), "f"(b));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
        s += lo + hi;
    }
#pragma unroll
    for (int k = 0; k < NS; k++) s += sc[k];
    if (s == 123.456f) sink[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0;
}

// variant 7: packed HALF FMAs (fma.rn.f16x2: two half FMAs per 32-bit register and instruction), 16 independent
// chains, two operands shared by all instructions — does the CUDA-core half path run at twice the FP32 rate on this
// chip? (The data point a half-precision pre-screen would stand on; the trace kernel does not use it.)
__global__ void __launch_bounds__(512) hfma2_peak(float* sink, float a, float b, long long* clocks)
{
    unsigned acc[kPeakChains], av, bv;
    asm("{ .reg .f16 h; cvt.rn.f16.f32 h, %1; mov.b32 %0, {h, h}; }" : "=r"(av) : "f"(a));
    asm("{ .reg .f16 h; cvt.rn.f16.f32 h, %1; mov.b32 %0, {h, h}; }" : "=r"(bv) : "f"(b));
#pragma unroll
    for (int k = 0; k < kPeakChains; k++)
        asm("{ .reg .f16 h; cvt.rn.f16.f32 h, %1; mov.b32 %0, {h, h}; }" : "=r"(acc[k]) : "f"(threadIdx.x * 1e-3f + k));
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kPeakIters; it++) {
#pragma unroll
        for (int k = 0; k < kPeakChains; k++) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(av), "r"(bv));
    }
    const long long t1 = clock64();
    unsigned s = 0u;
#pragma unroll
    for (int k = 0; k < kPeakChains; k++) s ^= acc[k];
    if (s == 0x12345678u) sink[0] = 1.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0;
}

cudaError_t run_ffma_peak(int variant, int n_sms, cudaStream_t stream, double* tflops, double* mhz)
{
    float* sink = nullptr;
    long long* clocks = nullptr;
    cudaError_t err = cudaMalloc(&sink, sizeof(float));
    if (err != cudaSuccess) return err;
    err = cudaMalloc(&clocks, sizeof(long long));
    if (err != cudaSuccess) { cudaFree(sink); return err; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = n_sms * 4 * 8, threads = 512;   // 8 waves of 4 CTAs per SM
    double best_ms = 1e30;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, stream);
        if (variant == 1)
            ffma_peak_packed<<<blocks, threads, 0, stream>>>(sink, 1.0000001f, 1e-7f, clocks);
        else if (variant == 2)
            ffma_peak_operands<2><<<blocks, threads, 0, stream>>>(sink, 1.0000001f, 1e-7f, clocks);
        else if (variant == 3)
            ffma_peak_operands<3><<<blocks, threads, 0, stream>>>(sink, 1.0000001f, 1e-7f, clocks);
        else if (variant == 4)
            ffma_peak_operands<4><<<blocks, threads, 0, stream>>>(sink, 1.0000001f, 1e-7f, clocks);
        else if (variant == 5)
            ffma_peak_mixed<4><<<blocks, threads, 0, stream>>>(sink, 1.0000001f, 1e-7f, clocks);
        else if (variant == 6)
            ffma_peak_mixed<8><<<blocks, threads, 0, stream>>>(sink, 1.0000001f, 1e-7f, clocks);
        else if (variant == 7)
            hfma2_peak<<<blocks, threads, 0, stream>>>(sink, 0.999f, 1e-3f, clocks);
        else
            ffma_peak_scalar<<<blocks, threads, 0, stream>>>(sink, 1.0000001f, 1e-7f, clocks);
        cudaEventRecord(e1, stream);
        err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err == cudaSuccess) {
        const double per_iter = variant == 5 ? 2.0 * (16 + 4) : variant == 6 ? 2.0 * (16 + 8) : variant == 7 ? 4.0 * kPeakChains : 2.0 * kPeakChains;
        const double flops = per_iter * kPeakIters * static_cast<double>(blocks) * threads;
        if (tflops) *tflops = flops / (best_ms * 1e-3) / 1e12;
        long long h_clocks = 0;
        cudaMemcpy(&h_clocks, clocks, sizeof h_clocks, cudaMemcpyDeviceToHost);
        // one CTA's loop in SM cycles vs. the whole launch in time: 32 CTAs per SM run in 8 waves
        if (mhz) *mhz = static_cast<double>(h_clocks) * 8.0 / (best_ms * 1e-3) / 1e6;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(clocks);
    return err;
}

}  // namespace rtx
