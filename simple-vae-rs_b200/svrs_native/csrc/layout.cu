// layout.cu - layout glue (NCHW-flat <-> NHWC), weight packing, casts, error plumbing.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace svrs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Every kernel launch of this library passes through check_launch() -> note_kernel(): a process-wide launch counter
// (bench.py's gpu_launches) and, between svrs_trace_reset() and svrs_trace(), the names of the kernels launched by the
// calling thread (profile attribution: which kernel a C-ABI entry point actually dispatched to).
static long long g_launches = 0;
static thread_local char g_trace[2048] = "";
static thread_local int g_trace_len = -1;   // -1 = tracing off

bool pdl_enabled() {
    // measured on B200: neutral for the graph-replayed step (graph nodes already launch back to back), so off by default
    static const bool on = [] { const char* e = getenv("SVRS_PDL"); return e && e[0] == '1'; }();
    return on;
}

void note_kernel(const char* what) {
    __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED);
    if (g_trace_len < 0) return;
    int n = (int)strlen(what);
    if (g_trace_len + n + 2 >= (int)sizeof(g_trace)) return;
    if (g_trace_len > 0) g_trace[g_trace_len++] = ',';
    memcpy(g_trace + g_trace_len, what, n + 1);
    g_trace_len += n;
}

// One image: src is [C][HW] (row stride HW), dst is [HW][C].  32x32 smem tile transpose.
template <typename TS, typename TD>
__global__ void nchw_to_nhwc_kernel(const TS* __restrict__ src, long long src_ld, TD* __restrict__ dst,
                                    int C, int HW) {
    pdl_entry();
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const TS* s = src + (long long)n * src_ld;
    TD* d = dst + (long long)n * C * HW;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? Cvt<TS>::to_f(s[(long long)c * HW + p]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) d[(long long)p * C + c] = Cvt<TD>::from_f(tile[threadIdx.x][i]);
    }
}

template <typename TS, typename TD>
__global__ void nhwc_to_nchw_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long dst_ld,
                                    int C, int HW, int accumulate) {
    pdl_entry();
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const TS* s = src + (long long)n * C * HW;
    TD* d = dst + (long long)n * dst_ld;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? Cvt<TS>::to_f(s[(long long)p * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) {
            long long o = (long long)c * HW + p;
            float v = tile[threadIdx.x][i];
            if (accumulate) v += Cvt<TD>::to_f(d[o]);
            d[o] = Cvt<TD>::from_f(v);
        }
    }
}

template <typename TD>
__global__ void pack_weights_kernel(const float* __restrict__ w, int d0, int d1, int kk, TD* __restrict__ p01,
                                    TD* __restrict__ p10) {
    pdl_entry();
    long long total = (long long)d0 * d1 * kk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int t = (int)(i % kk);
        long long r = i / kk;
        int b = (int)(r % d1);
        int a = (int)(r / d1);
        float v = w[i];
        if (p01) p01[((long long)t * d0 + a) * d1 + b] = Cvt<TD>::from_f(v);
        if (p10) p10[((long long)t * d1 + b) * d0 + a] = Cvt<TD>::from_f(v);
    }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long n) {
    pdl_entry();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        d[i] = Cvt<TD>::from_f(Cvt<TS>::to_f(s[i]));
}

template <typename TS, typename TD>
__global__ void copy2d_kernel(const TS* __restrict__ s, long long sld, TD* __restrict__ d, long long dld,
                              long long rows, int cols, int accumulate) {
    pdl_entry();
    long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / cols;
        int c = (int)(i % cols);
        float v = Cvt<TS>::to_f(s[r * sld + c]);
        if (accumulate) v += Cvt<TD>::to_f(d[r * dld + c]);
        d[r * dld + c] = Cvt<TD>::from_f(v);
    }
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, int act, long long n) {
    pdl_entry();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float yv = Cvt<T>::to_f(y[i]), g = Cvt<T>::to_f(dy[i]);
        float o = g;
        if (act == SVRS_ACT_SIGMOID) o = g * yv * (1.f - yv);
        else if (act == SVRS_ACT_HARDTANH7) o = (yv > -7.f && yv < 7.f) ? g : 0.f;
        dx[i] = Cvt<T>::from_f(o);
    }
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
    pdl_entry();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = fmaf(a, x[i], y[i]);
}

static inline unsigned grid_for(long long n, int block = 256) {
    long long b = (n + block - 1) / block;
    long long cap = 16LL * num_sms();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace svrs

using namespace svrs;

extern "C" const char* svrs_last_error(void) { return svrs::g_err; }
extern "C" int svrs_abi_version(void) { return 1; }
extern "C" int svrs_device_cc(int dev) {
    int ma = 0, mi = 0;
    if (cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return SVRS_E_CUDA;
    if (cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return SVRS_E_CUDA;
    return ma * 10 + mi;
}

#define DISPATCH2(sd, dd, CALL)                                                            \
    do {                                                                                   \
        if (sd == SVRS_F32 && dd == SVRS_F32) { CALL(float, float); }                      \
        else if (sd == SVRS_F32 && dd == SVRS_BF16) { CALL(float, __nv_bfloat16); }        \
        else if (sd == SVRS_BF16 && dd == SVRS_F32) { CALL(__nv_bfloat16, float); }        \
        else if (sd == SVRS_BF16 && dd == SVRS_BF16) { CALL(__nv_bfloat16, __nv_bfloat16); } \
        else { set_error("bad dtype pair %d,%d", sd, dd); return SVRS_E_ARG; }             \
    } while (0)

extern "C" int svrs_nchw_to_nhwc(const void* src, int src_dtype, int64_t src_ld, void* dst, int dst_dtype,
                                 int N, int C, int H, int W, void* stream) {
    SVRS_CHECK_ARG(src && dst && N >= 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad args");
    if (N == 0) return 0;
    SVRS_CHECK_ARG(N <= 65535, "nchw_to_nhwc: N > 65535");
    int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(TS, TD) SVRS_LAUNCH((nchw_to_nhwc_kernel<TS, TD>), grid, block, 0, st, (const TS*)src, src_ld, (TD*)dst, C, HW)
    DISPATCH2(src_dtype, dst_dtype, CALL);
#undef CALL
    return check_launch("nchw_to_nhwc");
}

extern "C" int svrs_nhwc_to_nchw(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t dst_ld,
                                 int N, int C, int H, int W, int accumulate, void* stream) {
    SVRS_CHECK_ARG(src && dst && N >= 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad args");
    if (N == 0) return 0;
    SVRS_CHECK_ARG(N <= 65535, "nhwc_to_nchw: N > 65535");
    int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(TS, TD) SVRS_LAUNCH((nhwc_to_nchw_kernel<TS, TD>), grid, block, 0, st, (const TS*)src, (TD*)dst, dst_ld, C, HW, accumulate)
    DISPATCH2(src_dtype, dst_dtype, CALL);
#undef CALL
    return check_launch("nhwc_to_nchw");
}

extern "C" int svrs_pack_weights(const float* w, int d0, int d1, int kk, void* p01, void* p10, int dtype,
                                 void* stream) {
    SVRS_CHECK_ARG(w && d0 > 0 && d1 > 0 && kk > 0, "pack_weights: bad args");
    long long total = (long long)d0 * d1 * kk;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32)
        SVRS_LAUNCH((pack_weights_kernel<float>), grid_for(total), 256, 0, st, w, d0, d1, kk, (float*)p01, (float*)p10);
    else if (dtype == SVRS_BF16)
        SVRS_LAUNCH((pack_weights_kernel<__nv_bfloat16>), grid_for(total), 256, 0, st, w, d0, d1, kk, (__nv_bfloat16*)p01,
                                                                             (__nv_bfloat16*)p10);
    else { set_error("pack_weights: bad dtype"); return SVRS_E_ARG; }
    return check_launch("pack_weights");
}

extern "C" int svrs_fill_zero(void* p, int64_t bytes, void* stream) {
    if (bytes <= 0) return 0;
    SVRS_CHECK_ARG(p, "fill_zero: null");
    cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("fill_zero: %s", cudaGetErrorString(e)); return SVRS_E_CUDA; }
    return 0;
}

extern "C" int svrs_axpy_f32(float* y, const float* x, float a, int64_t n, void* stream) {
    if (n <= 0) return 0;
    SVRS_CHECK_ARG(x && y, "axpy: null");
    SVRS_LAUNCH((axpy_kernel), grid_for(n), 256, 0, (cudaStream_t)stream, y, x, a, n);
    return check_launch("axpy");
}

extern "C" int svrs_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
    if (n <= 0) return 0;
    SVRS_CHECK_ARG(src && dst, "cast: null");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(TS, TD) SVRS_LAUNCH((cast_kernel<TS, TD>), grid_for(n), 256, 0, st, (const TS*)src, (TD*)dst, n)
    DISPATCH2(src_dtype, dst_dtype, CALL);
#undef CALL
    return check_launch("cast");
}

extern "C" int svrs_copy2d(const void* src, int src_dtype, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld,
                           int64_t rows, int cols, int accumulate, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    SVRS_CHECK_ARG(src && dst && (src_ld >= cols || src_ld == 0) && dst_ld >= cols, "copy2d: bad args (src_ld 0 = broadcast)");
    cudaStream_t st = (cudaStream_t)stream;
    long long n = rows * cols;
#define CALL(TS, TD) SVRS_LAUNCH((copy2d_kernel<TS, TD>), grid_for(n), 256, 0, st, (const TS*)src, src_ld, (TD*)dst, dst_ld, rows, cols, accumulate)
    DISPATCH2(src_dtype, dst_dtype, CALL);
#undef CALL
    return check_launch("copy2d");
}

extern "C" int svrs_act_bwd(const void* y, const void* dy, void* dx, int dtype, int act, int64_t n, void* stream) {
    if (n <= 0) return 0;
    SVRS_CHECK_ARG(y && dy && dx, "act_bwd: null");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32) SVRS_LAUNCH((act_bwd_kernel<float>), grid_for(n), 256, 0, st, (const float*)y, (const float*)dy, (float*)dx, act, n);
    else if (dtype == SVRS_BF16) SVRS_LAUNCH((act_bwd_kernel<__nv_bfloat16>), grid_for(n), 256, 0, st, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, act, n);
    else { set_error("act_bwd: bad dtype"); return SVRS_E_ARG; }
    return check_launch("act_bwd");
}

// ------------------------------------------------------------------------------------------------
// All layers' weight packs in ONE launch.  A CTA owns a 32 (d0) x 16 (d1) tile of one layer for all kk taps: it reads
// 32 contiguous runs of 16*kk fp32 (coalesced), transposes through shared memory and writes both packs with
// contiguous 16/32-element runs.  `jobs` is a device array built once by the host runtime.
// ------------------------------------------------------------------------------------------------
namespace svrs {
struct PackJob {
    const float* w;
    void* p01;
    void* p10;
    int d0, d1, kk;
    int tile0;     // first global tile index of this job
    int tiles_b;   // tiles along d1
    int pad_;
};
constexpr int PK_TA = 32, PK_TB = 16;

// job of a tile: binary search over the jobs' first-tile indices (staged in shared memory by one coalesced load; the
// former linear scan cost every CTA up to njobs dependent global loads - more than its whole 5k-element tile)
__device__ __forceinline__ int find_job(const PackJob* __restrict__ jobs, int njobs, int tile, int* s_tile0) {
    for (int i = threadIdx.x; i < njobs; i += blockDim.x) s_tile0[i] = jobs[i].tile0;
    __syncthreads();
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_tile0[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// i -> (b = i / kk, t = i % kk) for the two kernel sizes on the path without a runtime division
__device__ __forceinline__ void split_kk(int i, int kk, int& b, int& t) {
    if (kk == 16) { b = i >> 4; t = i & 15; }
    else if (kk == 9) { b = i / 9; t = i - 9 * b; }
    else { b = i / kk; t = i - kk * b; }
}
constexpr int PK_MAX_JOBS = 128;

template <typename TD>
__global__ void __launch_bounds__(256) pack_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
    pdl_entry();
    extern __shared__ float tile[];
    __shared__ int s_tile0[PK_MAX_JOBS];
    const PackJob jb = jobs[find_job(jobs, njobs, blockIdx.x, s_tile0)];
    const int lt = blockIdx.x - jb.tile0;
    const int a0 = (lt / jb.tiles_b) * PK_TA, b0 = (lt % jb.tiles_b) * PK_TB;
    const int kk = jb.kk, d0 = jb.d0, d1 = jb.d1;
    const int na = d0 - a0 < PK_TA ? d0 - a0 : PK_TA, nb = d1 - b0 < PK_TB ? d1 - b0 : PK_TB;
    const int ROW = PK_TB * (kk + 1) + 1;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int run = nb * kk;
    for (int a = warp; a < na; a += 8) {                           // read torch layout: contiguous runs of nb*kk floats
        const float* src = jb.w + ((long long)(a0 + a) * d1 + b0) * kk;
        for (int i = lane; i < run; i += 32) {
            int b, t;
            split_kk(i, kk, b, t);
            tile[a * ROW + b * (kk + 1) + t] = src[i];
        }
    }
    __syncthreads();
    TD* p01 = reinterpret_cast<TD*>(jb.p01);
    TD* p10 = reinterpret_cast<TD*>(jb.p10);
    if (p01) {                                                     // [t][a][b]: half-warps write 16 consecutive b
        const int tx = threadIdx.x % PK_TB, ty = threadIdx.x / PK_TB;
        if (tx < nb)
            for (int t = 0; t < kk; ++t)
                for (int a = ty; a < na; a += 256 / PK_TB)
                    p01[((long long)t * d0 + a0 + a) * d1 + b0 + tx] = Cvt<TD>::from_f(tile[a * ROW + tx * (kk + 1) + t]);
    }
    if (p10) {                                                     // [t][b][a]: warps write 32 consecutive a
        if (lane < na)
            for (int t = 0; t < kk; ++t)
                for (int b = warp; b < nb; b += 8)
                    p10[((long long)t * d1 + b0 + b) * d0 + a0 + lane] = Cvt<TD>::from_f(tile[lane * ROW + b * (kk + 1) + t]);
    }
}
}  // namespace svrs

namespace svrs {
// gradients: packed fp32 scratch [tap][d1][d0] -> torch layout [d0][d1][tap] (+=), same tiling as pack_multi_kernel
__global__ void __launch_bounds__(256) unpack_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
    pdl_entry();
    extern __shared__ float tile[];
    __shared__ int s_tile0[PK_MAX_JOBS];
    const PackJob jb = jobs[find_job(jobs, njobs, blockIdx.x, s_tile0)];
    const int lt = blockIdx.x - jb.tile0;
    const int a0 = (lt / jb.tiles_b) * PK_TA, b0 = (lt % jb.tiles_b) * PK_TB;
    const int kk = jb.kk, d0 = jb.d0, d1 = jb.d1;
    const int na = d0 - a0 < PK_TA ? d0 - a0 : PK_TA, nb = d1 - b0 < PK_TB ? d1 - b0 : PK_TB;
    const int ROW = PK_TB * (kk + 1) + 1;
    const float* src = reinterpret_cast<const float*>(jb.p01);
    float* dst = const_cast<float*>(jb.w);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (lane < na)                                                 // read packed [kk][d1][d0]: warps read 32 consecutive a
        for (int t = 0; t < kk; ++t)
            for (int b = warp; b < nb; b += 8)
                tile[lane * ROW + b * (kk + 1) + t] = src[((long long)t * d1 + b0 + b) * d0 + a0 + lane];
    __syncthreads();
    const int run = nb * kk;
    for (int a = warp; a < na; a += 8) {                          // write torch layout: contiguous runs of nb*kk
        float* d = dst + ((long long)(a0 + a) * d1 + b0) * kk;
        for (int i = lane; i < run; i += 32) {
            int b, t;
            split_kk(i, kk, b, t);
            d[i] += tile[a * ROW + b * (kk + 1) + t];
        }
    }
}
}  // namespace svrs

extern "C" int svrs_unpack_grads_multi(const void* jobs, int njobs, int total_tiles, int max_kk, void* stream) {
    SVRS_CHECK_ARG(jobs && njobs > 0 && njobs <= svrs::PK_MAX_JOBS && total_tiles > 0 && max_kk > 0 && max_kk <= 16, "unpack_grads_multi: bad args (at most 128 jobs)");
    size_t smem = (size_t)svrs::PK_TA * (svrs::PK_TB * (max_kk + 1) + 1) * sizeof(float);
    SVRS_LAUNCH((svrs::unpack_multi_kernel), total_tiles, 256, smem, (cudaStream_t)stream, (const svrs::PackJob*)jobs, njobs);
    return check_launch("unpack_grads_multi");
}

extern "C" int svrs_pack_job_bytes(void) { return (int)sizeof(svrs::PackJob); }

extern "C" int svrs_pack_weights_multi(const void* jobs, int njobs, int total_tiles, int max_kk, int dtype, void* stream) {
    SVRS_CHECK_ARG(jobs && njobs > 0 && njobs <= svrs::PK_MAX_JOBS && total_tiles > 0 && max_kk > 0 && max_kk <= 16, "pack_weights_multi: bad args (at most 128 jobs)");
    size_t smem = (size_t)svrs::PK_TA * (svrs::PK_TB * (max_kk + 1) + 1) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32) SVRS_LAUNCH((svrs::pack_multi_kernel<float>), total_tiles, 256, smem, st, (const svrs::PackJob*)jobs, njobs);
    else if (dtype == SVRS_BF16) SVRS_LAUNCH((svrs::pack_multi_kernel<__nv_bfloat16>), total_tiles, 256, smem, st, (const svrs::PackJob*)jobs, njobs);
    else { set_error("pack_weights_multi: bad dtype"); return SVRS_E_ARG; }
    return check_launch("pack_weights_multi");
}

extern "C" int64_t svrs_launch_count(void) { return __atomic_load_n(&svrs::g_launches, __ATOMIC_RELAXED); }
extern "C" void svrs_trace_reset(int enable) {
    svrs::g_trace[0] = 0;
    svrs::g_trace_len = enable ? 0 : -1;
}
extern "C" const char* svrs_trace(void) { return svrs::g_trace; }
