// tonemap.cu — EXTENSION (rtx_params.tonemap = RTX_TONEMAP_REINHARD; off by default, the reference packs radiance
// straight to 8 bits, main.cpp:338-347). README.md:13 speaks of tone mapping, the snapshot has no operator: parity is
// unpinned, the specification is oracle/oracle.c::orc_tonemap, which these kernels follow operation for operation.
//
//   pass 1  logsum   per frame, sum over pixels of log(1e-4 + L), L = .2126 R + .7152 G + .0722 B (negative / NaN -> 0)
//   pass 2  map+pack Lavg = exp(sum / n), Ls = key / Lavg * L, Ld = Ls (1 + Ls / white^2) / (1 + Ls), rgb *= Ld / L,
//                    then the same 8-bit pack as the quantise kernel
//
// Both are HBM-bound streaming kernels (pass 1 reads 12 or 24 B per pixel, pass 2 reads the same and writes 4 B):
// 4 pixels per thread with 128-bit loads and one 128-bit store, grid = 8 CTAs per SM, warp-shuffle reductions and one
// atomic per warp for the global statistic. The statistic is accumulated in 32.32 FIXED POINT with integer atomics:
// integer addition is associative, so the frame's log-average — and with it every pixel — is identical from run to
// run whatever the order in which warps finish (a floating-point atomicAdd would not be).
//
// FLOAT radiance (rtx_tonemap / rtx_tonemap_sums / _apply on an f32 buffer) is processed in FLOAT: luminance, logf, the map
// and the scale are single-precision operations with every rounding spelled out (the specification says so: oracle.c,
// branch rgb32), only the per-frame constants (exp of the mean, key / Lavg) and the 8-bit pack stay double. With double
// log / divide per pixel the f32 kernels were bound by the FP64 pipe at 46 % of the HBM roofline.
#include <algorithm>

#include "rtx_device.cuh"

namespace rtx {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr double kFix = 4294967296.0;   // 2^32

// Rec. 709 luminance with every rounding spelled out; negative and NaN luminances count as 0.
__device__ __forceinline__ double luminance(double r, double g, double b)
{
    const double l = ex::add(ex::add(ex::mul(0.2126, r), ex::mul(0.7152, g)), ex::mul(0.0722, b));
    return l > 0.0 ? l : 0.0;
}

__device__ __forceinline__ long long log_fixed(double r, double g, double b)
{
    return __double2ll_rn(ex::mul(log(ex::add(1e-4, luminance(r, g, b))), kFix));
}

__device__ __forceinline__ float luminance_f(float r, float g, float b)
{
    const float l = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, r), __fmul_rn(0.7152f, g)), __fmul_rn(0.0722f, b));
    return l > 0.0f ? l : 0.0f;
}

__device__ __forceinline__ long long log_fixed(float r, float g, float b)
{
    return __float2ll_rn(__fmul_rn(logf(__fadd_rn(1e-4f, luminance_f(r, g, b))), 4294967296.0f));
}

struct MapConsts {
    double key_over_avg;   // key / Lavg of this frame
    double inv_white2;     // 1 / white^2, or 0 when there is no burn-out term
};

__device__ __forceinline__ MapConsts map_consts(long long sum_fixed, long long pixels, double key, double white)
{
    const double mean = ex::div(ex::div(static_cast<double>(sum_fixed), kFix), static_cast<double>(pixels));
    MapConsts c;
    c.key_over_avg = ex::div(key, exp(mean));
    c.inv_white2 = white > 0.0 ? ex::div(1.0, ex::mul(white, white)) : 0.0;
    return c;
}

struct PackStats {
    unsigned long long over;
    double maxlum;
};

__device__ __forceinline__ uint32_t map_pixel(PackStats& s, const MapConsts& c, double r, double g, double b, int mode)
{
    const double l = luminance(r, g, b);
    const double ls = ex::mul(c.key_over_avg, l);
    const double ld = ex::div(ex::mul(ls, ex::add(1.0, ex::mul(ls, c.inv_white2))), ex::add(1.0, ls));
    const double k = l > 0.0 ? ex::div(ld, l) : 0.0;
    const double R = ex::mul(r, k), G = ex::mul(g, k), B = ex::mul(b, k);
    bool over;
    const uint32_t word = pack_rgba_flag(R, G, B, mode, over);
    if (over) s.over++;
    const double lum = (R + G + B) * (1.0 / 3.0);
    if (lum > s.maxlum) s.maxlum = lum;
    return word;
}

// float radiance: the map in single precision (see the header comment); pack and statistics as for doubles
__device__ __forceinline__ uint32_t map_pixel(PackStats& s, const MapConsts& c, float r, float g, float b, int mode)
{
    const float koa = static_cast<float>(c.key_over_avg), iw2 = static_cast<float>(c.inv_white2);
    const float l = luminance_f(r, g, b);
    const float ls = __fmul_rn(koa, l);
    const float ld = __fdiv_rn(__fmul_rn(ls, __fadd_rn(1.0f, __fmul_rn(ls, iw2))), __fadd_rn(1.0f, ls));
    const float k = l > 0.0f ? __fdiv_rn(ld, l) : 0.0f;
    const double R = __fmul_rn(r, k), G = __fmul_rn(g, k), B = __fmul_rn(b, k);
    bool over;
    const uint32_t word = pack_rgba_flag(R, G, B, mode, over);
    if (over) s.over++;
    const double lum = (R + G + B) * (1.0 / 3.0);
    if (lum > s.maxlum) s.maxlum = lum;
    return word;
}

// Working type of a radiance type: float stays float, double stays double.
template <typename T> struct Work { using type = double; };
template <> struct Work<float> { using type = float; };

// Four pixels (12 channels) of frame-relative quad q. VEC: 128-bit loads (the frame's base is 16-byte aligned).
template <typename T, bool VEC>
__device__ __forceinline__ void load_quad(const T* __restrict__ frame, long long q, typename Work<T>::type (&v)[12])
{
    if constexpr (VEC && sizeof(T) == 4) {
        const float4* in4 = reinterpret_cast<const float4*>(frame);
        const float4 a = __ldcs(&in4[3 * q]), b = __ldcs(&in4[3 * q + 1]), c = __ldcs(&in4[3 * q + 2]);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
    } else if constexpr (VEC) {
        const double2* in2 = reinterpret_cast<const double2*>(frame);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double2 d = __ldcs(&in2[6 * q + k]);
            v[2 * k] = d.x;
            v[2 * k + 1] = d.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) v[k] = frame[12 * q + k];
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) tonemap_logsum_kernel(const T* __restrict__ rad, long long pixels, long long* __restrict__ sums)
{
    const T* __restrict__ frame = rad + 3 * pixels * blockIdx.y;
    const long long n_quads = pixels >> 2;
    long long acc = 0;
    for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n_quads;
         q += static_cast<long long>(gridDim.x) * blockDim.x) {
        typename Work<T>::type v[12];
        load_quad<T, VEC>(frame, q, v);
#pragma unroll
        for (int k = 0; k < 4; k++) acc += log_fixed(v[3 * k], v[3 * k + 1], v[3 * k + 2]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (pixels & 3)) {   // ragged tail
        const long long p = (n_quads << 2) + threadIdx.x;
        acc += log_fixed(frame[3 * p], frame[3 * p + 1], frame[3 * p + 2]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(kFull, acc, off);
    if ((threadIdx.x & 31) == 0 && acc != 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(&sums[blockIdx.y]), static_cast<unsigned long long>(acc));
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) tonemap_pack_kernel(const T* __restrict__ rad, long long pixels, const long long* __restrict__ sums,
                                                           long long pixels_global, double key, double white, int mode,
                                                           uint32_t* __restrict__ out, unsigned long long* counters)
{
    // `pixels` = this buffer's pixels per frame; `pixels_global` = the pixels per frame the sums were taken over (more
    // than `pixels` when the frame's rows are spread over several GPUs and the sums were all-reduced)
    const T* __restrict__ frame = rad + 3 * pixels * blockIdx.y;
    uint32_t* __restrict__ oframe = out + pixels * blockIdx.y;
    const MapConsts c = map_consts(sums[blockIdx.y], pixels_global, key, white);
    PackStats st{0ull, 0.0};
    const long long n_quads = pixels >> 2;
    for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n_quads;
         q += static_cast<long long>(gridDim.x) * blockDim.x) {
        typename Work<T>::type v[12];
        load_quad<T, VEC>(frame, q, v);
        uint4 o;
        o.x = map_pixel(st, c, v[0], v[1], v[2], mode);
        o.y = map_pixel(st, c, v[3], v[4], v[5], mode);
        o.z = map_pixel(st, c, v[6], v[7], v[8], mode);
        o.w = map_pixel(st, c, v[9], v[10], v[11], mode);
        if constexpr (VEC) {
            __stcs(reinterpret_cast<uint4*>(oframe) + q, o);
        } else {
            oframe[4 * q] = o.x; oframe[4 * q + 1] = o.y; oframe[4 * q + 2] = o.z; oframe[4 * q + 3] = o.w;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (pixels & 3)) {
        const long long p = (n_quads << 2) + threadIdx.x;
        oframe[p] = map_pixel(st, c, frame[3 * p], frame[3 * p + 1], frame[3 * p + 2], mode);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        st.over += __shfl_down_sync(kFull, st.over, off);
        st.maxlum = fmax(st.maxlum, __shfl_down_sync(kFull, st.maxlum, off));
    }
    if ((threadIdx.x & 31) == 0 && counters) {
        if (st.over) atomicAdd(&counters[2], st.over);
        if (st.maxlum > 0.0) atomicMax(&counters[3], static_cast<unsigned long long>(__double_as_longlong(st.maxlum)));
    }
}

template <typename T>
dim3 stream_grid2(long long pixels, int n_frames, int n_sms)
{
    // whole waves: 8 CTAs of 256 threads per SM, shared between the frames of the call
    long long per_frame = (pixels / 4 + 255) / 256;
    const long long cap = std::max<long long>(1, static_cast<long long>(n_sms) * 8 / n_frames);
    per_frame = std::max<long long>(1, std::min(per_frame, cap));
    return dim3(static_cast<unsigned>(per_frame), static_cast<unsigned>(n_frames));
}

// vector path: every frame's base must be 16-byte aligned
inline bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

template <typename T>
cudaError_t launch_sums_typed(const T* rad, long long pixels, int n_frames, long long* sums, int n_sms, cudaStream_t stream)
{
    const dim3 grid = stream_grid2<T>(pixels, n_frames, n_sms);
    if ((pixels % 4 == 0) && aligned16(rad))
        tonemap_logsum_kernel<T, true><<<grid, 256, 0, stream>>>(rad, pixels, sums);
    else
        tonemap_logsum_kernel<T, false><<<grid, 256, 0, stream>>>(rad, pixels, sums);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_apply_typed(const T* rad, long long pixels, int n_frames, const long long* sums, long long pixels_global, double key,
                               double white, int mode, uint32_t* rgba8, unsigned long long* counters, int n_sms, cudaStream_t stream)
{
    const dim3 grid = stream_grid2<T>(pixels, n_frames, n_sms);
    if ((pixels % 4 == 0) && aligned16(rad) && aligned16(rgba8))
        tonemap_pack_kernel<T, true><<<grid, 256, 0, stream>>>(rad, pixels, sums, pixels_global, key, white, mode, rgba8, counters);
    else
        tonemap_pack_kernel<T, false><<<grid, 256, 0, stream>>>(rad, pixels, sums, pixels_global, key, white, mode, rgba8, counters);
    return cudaGetLastError();
}

}  // namespace

// Pass 1: ADDS this buffer's per-frame fixed-point sums to sums[n_frames] (the caller zeroes them first).
cudaError_t launch_tonemap_sums(const float* rad32, const double* rad64, int64_t pixels_per_frame, int n_frames, long long* sums,
                                int n_sms, cudaStream_t stream)
{
    if (pixels_per_frame <= 0 || n_frames <= 0) return cudaSuccess;
    return rad32 ? launch_sums_typed<float>(rad32, pixels_per_frame, n_frames, sums, n_sms, stream)
                 : launch_sums_typed<double>(rad64, pixels_per_frame, n_frames, sums, n_sms, stream);
}

// Pass 2: maps and packs with the given sums, taken over pixels_global pixels per frame.
cudaError_t launch_tonemap_apply(const float* rad32, const double* rad64, int64_t pixels_per_frame, int n_frames, const long long* sums,
                                 int64_t pixels_global, double key, double white, int mode, uint32_t* rgba8, unsigned long long* counters,
                                 int n_sms, cudaStream_t stream)
{
    if (pixels_per_frame <= 0 || n_frames <= 0) return cudaSuccess;
    return rad32 ? launch_apply_typed<float>(rad32, pixels_per_frame, n_frames, sums, pixels_global, key, white, mode, rgba8, counters, n_sms, stream)
                 : launch_apply_typed<double>(rad64, pixels_per_frame, n_frames, sums, pixels_global, key, white, mode, rgba8, counters, n_sms, stream);
}

}  // namespace rtx
