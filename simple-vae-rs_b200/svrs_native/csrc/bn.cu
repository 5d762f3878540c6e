// bn.cu - train/eval BatchNorm2d (+ReLU) forward and backward on NHWC activations viewed as [M][C].
// Statistics are accumulated in double (per-thread double partials, double atomics) so that the
// E[x^2]-E[x]^2 form is exact to fp32 output precision (fp32 parity mode needs 1e-5).
#include "common.cuh"
#include <stdlib.h>
#include "tc_ptx.cuh"

namespace svrs {

// 16-byte vector access: V = 4 floats or 8 bf16 (V = 4 bf16 = 8 bytes only when C is not a multiple of 8).
template <int V> __device__ __forceinline__ void ldv(const float* p, float* o) {
    static_assert(V == 4, "fp32 vectors are 4 wide");
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <int V> __device__ __forceinline__ void ldv(const __nv_bfloat16* p, float* o) {
    if (V == 8) {
        uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { o[2 * j] = __uint_as_float(w[j] << 16); o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
    } else {
        uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        const uint32_t w[2] = {r.x, r.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) { o[2 * j] = __uint_as_float(w[j] << 16); o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
    }
}
// raw 16-byte (or 8-byte) row fragments: kept unconverted while several loads are in flight (half the registers of floats)
template <typename T, int V> struct RawVec { uint32_t w[V * sizeof(T) / 4]; };
template <typename T, int V> __device__ __forceinline__ void ld_raw(const T* p, RawVec<T, V>& r) {
    constexpr int NW = V * sizeof(T) / 4;
    if (NW == 4) { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[NW - 1] = v.w; }
    else { uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); r.w[0] = v.x; r.w[1] = v.y; }
}
template <int V> __device__ __forceinline__ void raw_cvt(const RawVec<float, V>& r, float* o) {
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = __uint_as_float(r.w[j]);
}
template <int V> __device__ __forceinline__ void raw_cvt(const RawVec<__nv_bfloat16, V>& r, float* o) {
#pragma unroll
    for (int j = 0; j < V / 2; ++j) { o[2 * j] = __uint_as_float(r.w[j] << 16); o[2 * j + 1] = __uint_as_float(r.w[j] & 0xffff0000u); }
}

template <int V> __device__ __forceinline__ void stv(float* p, const float* o) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <int V> __device__ __forceinline__ void stv(__nv_bfloat16* p, const float* o) {
    uint32_t w[V / 2];
#pragma unroll
    for (int j = 0; j < V / 2; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    if (V == 8) *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[V / 2 - 2], w[V / 2 - 1]);
    else *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
}

// Forward statistics of bf16 tensors are accumulated as fp32 sums of (x - k) and (x - k)^2 with k = the channel's first
// sample (FP64 throughput, not HBM, bound the per-element double version: 2.7 TB/s).  A shift within a few sigma of the mean
// removes the E[x^2] - E[x]^2 cancellation that rules out plain fp32 sums; the shift is undone in double per thread:
//   sum x = S1 + n k ,  sum x^2 = S2 + 2 k S1 + n k^2.   fp32 tensors (parity mode) keep the per-element double path.
template <int V>
__device__ __forceinline__ void unshift_sums(double* s0, double* s1, const float* k, long long n) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const double kd = (double)k[j], S1 = s0[j], S2 = s1[j];
        s0[j] = S1 + (double)n * kd;
        s1[j] = S2 + 2.0 * kd * S1 + (double)n * kd * kd;
    }
}

// End of a reduce block: fold the row lanes through shared memory (`red`: 256 x V doubles) and add the block's 2C sums
// to ONE of SVRS_BN_REPLICAS copies of the `sums` scratch (replica = block index mod R); every reader adds the replicas
// up (bn_load_sums).  Measured: double atomics on the same 128-byte line serialise at ~3 ns each, so ~300 blocks x 16
// channels per line put a ~15 us floor under every reduce launch, whatever the tensor size; 8 replicas cut it 8x.
// (Thread-block clusters + DSMEM gave the same reduction in atomics but halved the bandwidth of the large launches:
// clusters of 8 one-CTA-per-SM blocks leave SMs of every GPC unused.)
template <int V>
__device__ __forceinline__ void block_sums_to_global(const double* s0, const double* s1, double* red, int C, int cg, int q,
                                                     int lane, int lanes, int c, double* __restrict__ sums) {
    double* dst = sums + (size_t)(blockIdx.x % SVRS_BN_REPLICAS) * 2 * C;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < V; ++j) red[threadIdx.x * V + j] = half ? s1[j] : s0[j];
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                double acc = 0;
                for (int l = 0; l < lanes; ++l) acc += red[(l * cg + q) * V + j];
                atomicAdd(&dst[half * C + c + j], acc);
            }
        }
    }
}
// entry e (0 <= e < 2C) of the replicated sums
__device__ __forceinline__ double bn_load_sums(const double* __restrict__ sums, int C, int e) {
    double acc = 0;
#pragma unroll
    for (int r = 0; r < SVRS_BN_REPLICAS; ++r) acc += sums[(size_t)r * 2 * C + e];
    return acc;
}

// thread t owns channel group (t % cg) of V channels and row lane (t / cg); cg = C/V divides 256.  Four rows are loaded
// per trip so every thread keeps 4 (forward) or 8 (backward) 16-byte loads in flight - these kernels are pure streaming.
template <typename T, int V, bool BWD>
__global__ void __launch_bounds__(256, 2) bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                         long long M, int C, long long rows_per_block,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                                         int relu, double* __restrict__ sums) {
    pdl_entry();
    constexpr int U = 4;
    __shared__ double red[256 * 8];
    const int cg = C / V;
    const int q = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = 256 / cg;
    const int c = q * V;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    double s0[V], s1[V];
    float sc[V], sh[V], mu[V], is[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { s0[j] = 0; s1[j] = 0; }
    if (BWD) {
#pragma unroll
        for (int j = 0; j < V; ++j) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; mu[j] = mean[c + j]; is[j] = invstd[c + j]; }
    }
    // Backward sums (sum g, sum g*xhat): fp32 partials over at most 16 rows folded into the double accumulators.  The
    // forward statistics stay in double per element: var = E[x^2] - E[x]^2 amplifies any rounding of E[x^2] by
    // mean^2/var, and fp32 partial sums measurably broke fp32 gradient parity (1e-3) on channels with |mean| >> std.
    float p0[V], p1[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { p0[j] = 0.f; p1[j] = 0.f; }
    float kshift[V];
    long long nrows_t = 0;
    if (!BWD && sizeof(T) == 2) ldv<V>(x + c, kshift);
    auto accumulate = [&](const float* xs, const float* gs) {
        if (!BWD && sizeof(T) == 4) {
#pragma unroll
            for (int j = 0; j < V; ++j) { s0[j] += (double)xs[j]; s1[j] += (double)xs[j] * (double)xs[j]; }
        } else if (!BWD) {
#pragma unroll
            for (int j = 0; j < V; ++j) { const float d = xs[j] - kshift[j]; p0[j] += d; p1[j] = fmaf(d, d, p1[j]); }
            ++nrows_t;
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float g = gs[j];
                if (relu && !(fmaf(xs[j], sc[j], sh[j]) > 0.f)) g = 0.f;
                float xh = (xs[j] - mu[j]) * is[j];
                p0[j] += g;
                p1[j] = fmaf(g, xh, p1[j]);
            }
        }
    };
    auto fold = [&]() {
#pragma unroll
        for (int j = 0; j < V; ++j) { s0[j] += (double)p0[j]; s1[j] += (double)p1[j]; p0[j] = 0.f; p1[j] = 0.f; }
    };
    long long r = r0 + lane;
    int trips = 0;
    for (; r + (long long)(U - 1) * lanes < r1; r += (long long)U * lanes) {
        RawVec<T, V> xr[U], gr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ld_raw<T, V>(x + (r + (long long)u * lanes) * C + c, xr[u]);
            if (BWD) ld_raw<T, V>(dy + (r + (long long)u * lanes) * C + c, gr[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float xs[V], gs[V];
            raw_cvt<V>(xr[u], xs);
            if (BWD) raw_cvt<V>(gr[u], gs);
            accumulate(xs, gs);
        }
        if ((++trips & 3) == 0) fold();
    }
    for (; r < r1; r += lanes) {
        float xs[V], gs[V];
        ldv<V>(x + r * C + c, xs);
        if (BWD) ldv<V>(dy + r * C + c, gs);
        accumulate(xs, gs);
    }
    fold();
    if (!BWD && sizeof(T) == 2) unshift_sums<V>(s0, s1, kshift, nrows_t);
    block_sums_to_global<V>(s0, s1, red, C, cg, q, lane, lanes, c, sums);
}

// ------------------------------------------------------------------------------------------------
// Streaming variant of bn_reduce_kernel for large tensors.  The register-file version above can keep at most
// threads x 4 rows x 16 B in flight and drains that window every trip (measured 2.0-2.7 TB/s); here ONE thread streams
// the block's contiguous row range through a shared-memory ring with 1-D bulk copies (cp.async.bulk, mbarrier
// complete_tx), so 64 KB (forward) / 128 KB (backward) per SM are in flight continuously while all 256 threads only
// read shared memory and accumulate.  Same thread <-> (channel group, row lane) mapping, same final reduction.
// ------------------------------------------------------------------------------------------------
constexpr int BNS_CHUNK = 16384;      // bytes per tensor per stage
constexpr int BNS_STAGES = 4;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <typename T, int V, bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_stream_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                long long M, int C, long long rows_per_block,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                int relu, double* __restrict__ sums) {
    pdl_entry();
    extern __shared__ __align__(128) uint8_t bns_smem[];
    __shared__ __align__(8) unsigned long long bars[BNS_STAGES];
    constexpr int NT = BWD ? 2 : 1;
    const int cg = C / V;
    const int q = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = 256 / cg;
    const int c = q * V;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    const int row_bytes = C * (int)sizeof(T);
    const int rows_per_chunk = BNS_CHUNK / row_bytes;
    const long long nrows = r1 > r0 ? r1 - r0 : 0;
    const int nchunks = (int)((nrows + rows_per_chunk - 1) / rows_per_chunk);
    const uint32_t smem0 = smem_u32(bns_smem);
    auto bar = [&](int s_) { return smem_u32(&bars[s_]); };
    auto issue = [&](int chunk) {          // thread 0 only
        const int s_ = chunk % BNS_STAGES;
        const long long row = r0 + (long long)chunk * rows_per_chunk;
        long long nr = r1 - row;
        if (nr > rows_per_chunk) nr = rows_per_chunk;
        const uint32_t bytes = (uint32_t)nr * (uint32_t)row_bytes;
        mbar_expect_tx(bar(s_), bytes * NT);
        bulk_load_1d(smem0 + (uint32_t)(s_ * NT) * BNS_CHUNK, x + row * C, bytes, bar(s_));
        if (BWD) bulk_load_1d(smem0 + (uint32_t)(s_ * NT + 1) * BNS_CHUNK, dy + row * C, bytes, bar(s_));
    };
    if (threadIdx.x == 0) {
        for (int s_ = 0; s_ < BNS_STAGES; ++s_) mbar_init(bar(s_), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < BNS_STAGES && k < nchunks; ++k) issue(k);

    double s0[V], s1[V];
    float p0[V], p1[V];
    float sc[V], sh[V], mu[V], is[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { s0[j] = 0; s1[j] = 0; p0[j] = 0.f; p1[j] = 0.f; }
    if (BWD) {
#pragma unroll
        for (int j = 0; j < V; ++j) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; mu[j] = mean[c + j]; is[j] = invstd[c + j]; }
    }
    float kshift[V];
    long long nrows_t = 0;
    if (!BWD && sizeof(T) == 2) ldv<V>(x + c, kshift);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int s_ = chunk % BNS_STAGES;
        mbar_wait(bar(s_), (uint32_t)((chunk / BNS_STAGES) & 1));
        long long nr = nrows - (long long)chunk * rows_per_chunk;
        if (nr > rows_per_chunk) nr = rows_per_chunk;
        const T* xs_ = reinterpret_cast<const T*>(bns_smem + (size_t)(s_ * NT) * BNS_CHUNK);
        const T* gs_ = reinterpret_cast<const T*>(bns_smem + (size_t)(s_ * NT + 1) * BNS_CHUNK);
        for (int r = lane; r < (int)nr; r += lanes) {
            RawVec<T, V> xr, gr;
            constexpr int NW = V * sizeof(T) / 4;
            const uint32_t* px = reinterpret_cast<const uint32_t*>(xs_ + (size_t)r * C + c);
            if (NW == 4) { uint4 v = *reinterpret_cast<const uint4*>(px); xr.w[0] = v.x; xr.w[1] = v.y; xr.w[2] = v.z; xr.w[NW - 1] = v.w; }
            else { uint2 v = *reinterpret_cast<const uint2*>(px); xr.w[0] = v.x; xr.w[1] = v.y; }
            float xs[V], gs[V];
            raw_cvt<V>(xr, xs);
            if (BWD) {
                const uint32_t* pg = reinterpret_cast<const uint32_t*>(gs_ + (size_t)r * C + c);
                if (NW == 4) { uint4 v = *reinterpret_cast<const uint4*>(pg); gr.w[0] = v.x; gr.w[1] = v.y; gr.w[2] = v.z; gr.w[NW - 1] = v.w; }
                else { uint2 v = *reinterpret_cast<const uint2*>(pg); gr.w[0] = v.x; gr.w[1] = v.y; }
                raw_cvt<V>(gr, gs);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float g = gs[j];
                    if (relu && !(fmaf(xs[j], sc[j], sh[j]) > 0.f)) g = 0.f;
                    float xh = (xs[j] - mu[j]) * is[j];
                    p0[j] += g;
                    p1[j] = fmaf(g, xh, p1[j]);
                }
            } else if (sizeof(T) == 4) {
#pragma unroll
                for (int j = 0; j < V; ++j) { s0[j] += (double)xs[j]; s1[j] += (double)xs[j] * (double)xs[j]; }
            } else {
                // bf16 input: fp32 sums of (x - k), k = the channel's first sample (see bn_reduce_kernel)
#pragma unroll
                for (int j = 0; j < V; ++j) { const float d = xs[j] - kshift[j]; p0[j] += d; p1[j] = fmaf(d, d, p1[j]); }
                ++nrows_t;
            }
        }
        if ((BWD || sizeof(T) == 2) && (chunk & 3) == 3) {
#pragma unroll
            for (int j = 0; j < V; ++j) { s0[j] += (double)p0[j]; s1[j] += (double)p1[j]; p0[j] = 0.f; p1[j] = 0.f; }
        }
        __syncthreads();                                   // everyone is done with stage s_
        if (threadIdx.x == 0 && chunk + BNS_STAGES < nchunks) issue(chunk + BNS_STAGES);
    }
    if (BWD || sizeof(T) == 2) {
#pragma unroll
        for (int j = 0; j < V; ++j) { s0[j] += (double)p0[j]; s1[j] += (double)p1[j]; }
    }
    if (!BWD && sizeof(T) == 2) unshift_sums<V>(s0, s1, kshift, nrows_t);
    // cross-lane reduction through the (now idle) ring: 256 x 8 doubles = 16 KB
    __syncthreads();
    block_sums_to_global<V>(s0, s1, reinterpret_cast<double*>(bns_smem), C, cg, q, lane, lanes, c, sums);
}

// scale/shift (+ mean, invstd, unbiased variance) of one channel from its batch sums: the single definition used by the
// stand-alone finalize kernel and by the fused statistics->apply kernel, so both give bit-identical coefficients
__device__ __forceinline__ void bn_channel_coeffs(const double* __restrict__ sums, long long M, int C, int c,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                  float& sc, float& sh, float& meanf, float& is, float& unb) {
    double m = bn_load_sums(sums, C, c) / (double)M;
    double var = bn_load_sums(sums, C, C + c) / (double)M - m * m;
    if (var < 0) var = 0;
    meanf = (float)m;
    float varf = (float)var;
    is = 1.0f / sqrtf(varf + eps);
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    sc = g * is;
    sh = b - meanf * sc;
    unb = (M > 1) ? (float)(var * (double)M / (double)(M - 1)) : varf;
}

__device__ __forceinline__ void bn_update_running(float* __restrict__ rmean, float* __restrict__ rvar, int c, float meanf,
                                                  float unb, float momentum, int n_updates) {
    float rm = rmean[c], rv = rvar[c];
    for (int u = 0; u < n_updates; ++u) {
        rm = (1.f - momentum) * rm + momentum * meanf;
        rv = (1.f - momentum) * rv + momentum * unb;
    }
    rmean[c] = rm;
    rvar[c] = rv;
}

__global__ void bn_finalize_train_kernel(const double* __restrict__ sums, long long M, int C,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float eps, float momentum, float* __restrict__ rmean, float* __restrict__ rvar,
                                         long long* __restrict__ nbt, int n_updates,
                                         float* __restrict__ scale, float* __restrict__ shift,
                                         float* __restrict__ mean_out, float* __restrict__ invstd_out) {
    pdl_entry();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += n_updates;
    if (c >= C) return;
    float sc, sh, meanf, is, unb;
    bn_channel_coeffs(sums, M, C, c, gamma, beta, eps, sc, sh, meanf, is, unb);
    scale[c] = sc;
    shift[c] = sh;
    if (mean_out) mean_out[c] = meanf;
    if (invstd_out) invstd_out[c] = is;
    if (rmean && rvar) bn_update_running(rmean, rvar, c, meanf, unb, momentum, n_updates);
}

__global__ void bn_finalize_eval_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                        const float* __restrict__ rmean, const float* __restrict__ rvar,
                                        float* __restrict__ scale, float* __restrict__ shift) {
    pdl_entry();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float is = 1.0f / sqrtf(rvar[c] + eps);
    float sc = (gamma ? gamma[c] : 1.f) * is;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - rmean[c] * sc;
}

// The grid stride (gridDim * 256 vectors) is a multiple of cg, so a thread's channel group never changes: its per-channel
// coefficients are loaded once.  Two vectors per trip for memory-level parallelism.
template <typename T, int V>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec, int cg,
                                                        const float* __restrict__ scale, const float* __restrict__ shift, int relu) {
    pdl_entry();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(i % cg) * V;
    float sc[V], sh[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; }
    auto apply = [&](float* v) {
#pragma unroll
        for (int j = 0; j < V; ++j) { v[j] = fmaf(v[j], sc[j], sh[j]); if (relu) v[j] = fmaxf(v[j], 0.f); }
    };
    for (; i + stride < nvec; i += 2 * stride) {
        float a[V], b[V];
        ldv<V>(x + i * V, a);
        ldv<V>(x + (i + stride) * V, b);
        apply(a); apply(b);
        stv<V>(y + i * V, a);
        stv<V>(y + (i + stride) * V, b);
    }
    if (i < nvec) {
        float a[V];
        ldv<V>(x + i * V, a);
        apply(a);
        stv<V>(y + i * V, a);
    }
}

// Train-mode finalize + apply in one launch: every block derives the per-channel coefficients from the batch sums into
// shared memory (C <= 1024), block 0 also publishes them (backward needs scale/shift/mean/invstd) and advances the
// running statistics.  Saves one tiny launch per BatchNorm layer.
template <typename T, int V>
__global__ void __launch_bounds__(256) bn_apply_train_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec, int cg,
                                                              long long M, int C, const double* __restrict__ sums,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, float momentum, float* __restrict__ rmean,
                                                              float* __restrict__ rvar, long long* __restrict__ nbt, int n_updates,
                                                              int relu, float* __restrict__ scale, float* __restrict__ shift,
                                                              float* __restrict__ mean_out, float* __restrict__ invstd_out) {
    pdl_entry();
    __shared__ float s_sc[1024], s_sh[1024];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sc, sh, meanf, is, unb;
        bn_channel_coeffs(sums, M, C, c, gamma, beta, eps, sc, sh, meanf, is, unb);
        s_sc[c] = sc;
        s_sh[c] = sh;
        if (blockIdx.x == 0) {
            scale[c] = sc;
            shift[c] = sh;
            if (mean_out) mean_out[c] = meanf;
            if (invstd_out) invstd_out[c] = is;
            if (rmean && rvar) bn_update_running(rmean, rvar, c, meanf, unb, momentum, n_updates);
            if (c == 0 && nbt) *nbt += n_updates;
        }
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (int)(i % cg) * V;
    float sc[V], sh[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { sc[j] = s_sc[c0 + j]; sh[j] = s_sh[c0 + j]; }
    auto apply = [&](float* v) {
#pragma unroll
        for (int j = 0; j < V; ++j) { v[j] = fmaf(v[j], sc[j], sh[j]); if (relu) v[j] = fmaxf(v[j], 0.f); }
    };
    for (; i + stride < nvec; i += 2 * stride) {
        float a[V], b[V];
        ldv<V>(x + i * V, a);
        ldv<V>(x + (i + stride) * V, b);
        apply(a); apply(b);
        stv<V>(y + i * V, a);
        stv<V>(y + (i + stride) * V, b);
    }
    if (i < nvec) {
        float a[V];
        ldv<V>(x + i * V, a);
        apply(a);
        stv<V>(y + i * V, a);
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                            long long nvec, int cg, long long M, int C,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, int relu,
                                                            const double* __restrict__ sums,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
    pdl_entry();
    // collapse the replicated sums once per block (C x 2 x R loads) instead of once per thread
    __shared__ float s_mg[1024], s_mgx[1024];
    const float invM = 1.0f / (float)M;
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
        const double sg = bn_load_sums(sums, C, cc), sgx = bn_load_sums(sums, C, C + cc);
        s_mg[cc] = (float)sg * invM;
        s_mgx[cc] = (float)sgx * invM;
        if (blockIdx.x == 0) {
            if (dbeta) dbeta[cc] += (float)sg;
            if (dgamma) dgamma[cc] += (float)sgx;
        }
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(i % cg) * V;
    float sc[V], sh[V], mu[V], is[V], mg[V], mgx[V], gi[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        sc[j] = scale[c + j]; sh[j] = shift[c + j]; mu[j] = mean[c + j]; is[j] = invstd[c + j];
        mg[j] = s_mg[c + j];
        mgx[j] = s_mgx[c + j];
        gi[j] = (gamma ? gamma[c + j] : 1.f) * is[j];
    }
    auto apply = [&](const float* xs, float* gs) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float g = gs[j];
            if (relu && !(fmaf(xs[j], sc[j], sh[j]) > 0.f)) g = 0.f;
            float xh = (xs[j] - mu[j]) * is[j];
            gs[j] = gi[j] * (g - mg[j] - xh * mgx[j]);
        }
    };
    for (; i + stride < nvec; i += 2 * stride) {
        float xa[V], ga[V], xb[V], gb[V];
        ldv<V>(x + i * V, xa);
        ldv<V>(dy + i * V, ga);
        ldv<V>(x + (i + stride) * V, xb);
        ldv<V>(dy + (i + stride) * V, gb);
        apply(xa, ga); apply(xb, gb);
        stv<V>(dx + i * V, ga);
        stv<V>(dx + (i + stride) * V, gb);
    }
    if (i < nvec) {
        float xa[V], ga[V];
        ldv<V>(x + i * V, xa);
        ldv<V>(dy + i * V, ga);
        apply(xa, ga);
        stv<V>(dx + i * V, ga);
    }
}

static bool c_ok(int C) { return C >= 4 && C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0; }
// bf16 with C % 8 == 0 moves 8 channels (16 bytes) per access
static bool wide(int dtype, int C) { return dtype == SVRS_BF16 && C % 8 == 0 && 256 % (C / 8) == 0; }

static void reduce_grid(long long M, int C, int V, unsigned& blocks, long long& rpb) {
    int lanes = 256 / (C / V);
    long long b = (M + (long long)lanes * 8 - 1) / ((long long)lanes * 8);       // >= 8 rows per thread
    // every block ends with 2C double atomics on the same few cache lines: keep the block count low (measured: with
    // 8 blocks per SM the same-line atomics, not the streaming loop, set the kernel time)
    long long cap = 2LL * num_sms();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    rpb = (M + b - 1) / b;
    blocks = (unsigned)((M + rpb - 1) / rpb);
}

// element-wise BN kernels: every thread first loads its V channels' coefficients (up to 56 scalar loads), so give each
// thread >= 8 vectors of work where the tensor allows and cap the grid at 4 blocks per SM (measured: with 16 blocks per
// SM the coefficient prologue, not the streaming loop, dominated the small layers)
static unsigned ew_grid(long long n) {
    long long b = (n + 256 * 8 - 1) / (256 * 8), cap = 4LL * num_sms();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// large tensors go through the bulk-copy streaming kernel (one or two blocks per SM); returns false if not applicable
template <typename T, int V, bool BWD>
static bool launch_reduce_stream(const void* x, const void* dy, long long M, int C, const float* scale, const float* shift,
                                 const float* mean, const float* invstd, int relu, double* sums, cudaStream_t st) {
    const long long bytes = M * C * (long long)sizeof(T);
    const int row_bytes = C * (int)sizeof(T);
    static const bool off = getenv("SVRS_BN_STREAM") && getenv("SVRS_BN_STREAM")[0] == '0';
    if (off || bytes < (4ll << 20) || row_bytes > BNS_CHUNK || row_bytes % 16 != 0) return false;
    const int smem = BNS_STAGES * BNS_CHUNK * (BWD ? 2 : 1);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(bn_reduce_stream_kernel<T, V, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    long long blocks = (long long)num_sms() * (BWD ? 1 : 2);
    const int rows_per_chunk = BNS_CHUNK / row_bytes;
    long long rpb = (M + blocks - 1) / blocks;
    rpb = (rpb + rows_per_chunk - 1) / rows_per_chunk * rows_per_chunk;     // whole chunks per block
    blocks = (M + rpb - 1) / rpb;
    SVRS_LAUNCH((bn_reduce_stream_kernel<T, V, BWD>), (unsigned)blocks, 256, smem, st, (const T*)x, (const T*)dy, M, C, rpb,
                scale, shift, mean, invstd, relu, sums);
    return true;
}

}  // namespace svrs

using namespace svrs;

extern "C" int svrs_bn_stats(const void* x, int dtype, int64_t M, int C, double* sums, void* stream) {
    SVRS_CHECK_ARG(x && sums && M > 0 && c_ok(C), "bn_stats: bad args (C=%d must be 4*2^k <= 1024)", C);
    unsigned blocks; long long rpb;
    const bool w8 = wide(dtype, C);
    reduce_grid(M, C, w8 ? 8 : 4, blocks, rpb);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32 ? launch_reduce_stream<float, 4, false>(x, nullptr, M, C, nullptr, nullptr, nullptr, nullptr, 0, sums, st)
        : (w8 ? launch_reduce_stream<__nv_bfloat16, 8, false>(x, nullptr, M, C, nullptr, nullptr, nullptr, nullptr, 0, sums, st)
              : launch_reduce_stream<__nv_bfloat16, 4, false>(x, nullptr, M, C, nullptr, nullptr, nullptr, nullptr, 0, sums, st)))
        return check_launch("bn_stats");
    if (dtype == SVRS_F32)
        SVRS_LAUNCH((bn_reduce_kernel<float, 4, false>), blocks, 256, 0, st, (const float*)x, nullptr, M, C, rpb, nullptr, nullptr, nullptr, nullptr, 0, sums);
    else if (dtype == SVRS_BF16 && w8)
        SVRS_LAUNCH((bn_reduce_kernel<__nv_bfloat16, 8, false>), blocks, 256, 0, st, (const __nv_bfloat16*)x, nullptr, M, C, rpb, nullptr, nullptr, nullptr, nullptr, 0, sums);
    else if (dtype == SVRS_BF16)
        SVRS_LAUNCH((bn_reduce_kernel<__nv_bfloat16, 4, false>), blocks, 256, 0, st, (const __nv_bfloat16*)x, nullptr, M, C, rpb, nullptr, nullptr, nullptr, nullptr, 0, sums);
    else { set_error("bn_stats: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_stats");
}

extern "C" int svrs_bn_finalize_train(const double* sums, int64_t M, int C, const float* gamma, const float* beta,
                                      float eps, float momentum, float* running_mean, float* running_var,
                                      int64_t* num_batches_tracked, int n_updates,
                                      float* scale, float* shift, float* mean, float* invstd, void* stream) {
    SVRS_CHECK_ARG(sums && scale && shift && M > 0 && C > 0, "bn_finalize_train: bad args");
    SVRS_LAUNCH((bn_finalize_train_kernel), (C + 127) / 128, 128, 0, (cudaStream_t)stream, 
        sums, M, C, gamma, beta, eps, momentum, running_mean, running_var, (long long*)num_batches_tracked, n_updates,
        scale, shift, mean, invstd);
    return check_launch("bn_finalize_train");
}

extern "C" int svrs_bn_finalize_eval(int C, const float* gamma, const float* beta, float eps,
                                     const float* running_mean, const float* running_var,
                                     float* scale, float* shift, void* stream) {
    SVRS_CHECK_ARG(running_mean && running_var && scale && shift && C > 0, "bn_finalize_eval: bad args");
    SVRS_LAUNCH((bn_finalize_eval_kernel), (C + 127) / 128, 128, 0, (cudaStream_t)stream, C, gamma, beta, eps, running_mean, running_var, scale, shift);
    return check_launch("bn_finalize_eval");
}

extern "C" int svrs_bn_apply(const void* x, void* y, int dtype, int64_t M, int C, const float* scale,
                             const float* shift, int relu, void* stream) {
    SVRS_CHECK_ARG(x && y && scale && shift && M > 0 && c_ok(C), "bn_apply: bad args (C=%d must be 4*2^k <= 1024)", C);
    const bool w8 = wide(dtype, C);
    long long nvec = M * C / (w8 ? 8 : 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32)
        SVRS_LAUNCH((bn_apply_kernel<float, 4>), ew_grid(nvec), 256, 0, st, (const float*)x, (float*)y, nvec, C / 4, scale, shift, relu);
    else if (dtype == SVRS_BF16 && w8)
        SVRS_LAUNCH((bn_apply_kernel<__nv_bfloat16, 8>), ew_grid(nvec), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, nvec, C / 8, scale, shift, relu);
    else if (dtype == SVRS_BF16)
        SVRS_LAUNCH((bn_apply_kernel<__nv_bfloat16, 4>), ew_grid(nvec), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, nvec, C / 4, scale, shift, relu);
    else { set_error("bn_apply: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_apply");
}

extern "C" int svrs_bn_apply_train(const void* x, void* y, int dtype, int64_t M, int64_t M_stat, int C, const double* sums,
                                   const float* gamma, const float* beta, float eps, float momentum,
                                   float* running_mean, float* running_var, int64_t* num_batches_tracked, int n_updates,
                                   int relu, float* scale, float* shift, float* mean, float* invstd, void* stream) {
    SVRS_CHECK_ARG(x && y && sums && scale && shift && M > 0 && c_ok(C) && C <= 1024, "bn_apply_train: bad args (C=%d must be 4*2^k <= 1024)", C);
    const bool w8 = wide(dtype, C);
    const long long Ms = M_stat > 0 ? M_stat : M;       // rows the statistics in `sums` were taken over (sync_bn: global batch)
    long long nvec = M * C / (w8 ? 8 : 4);
    cudaStream_t st = (cudaStream_t)stream;
    long long* nbt = (long long*)num_batches_tracked;
    if (dtype == SVRS_F32)
        SVRS_LAUNCH((bn_apply_train_kernel<float, 4>), ew_grid(nvec), 256, 0, st, (const float*)x, (float*)y, nvec, C / 4, Ms, C, sums, gamma, beta, eps, momentum, running_mean, running_var, nbt, n_updates, relu, scale, shift, mean, invstd);
    else if (dtype == SVRS_BF16 && w8)
        SVRS_LAUNCH((bn_apply_train_kernel<__nv_bfloat16, 8>), ew_grid(nvec), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, nvec, C / 8, Ms, C, sums, gamma, beta, eps, momentum, running_mean, running_var, nbt, n_updates, relu, scale, shift, mean, invstd);
    else if (dtype == SVRS_BF16)
        SVRS_LAUNCH((bn_apply_train_kernel<__nv_bfloat16, 4>), ew_grid(nvec), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, nvec, C / 4, Ms, C, sums, gamma, beta, eps, momentum, running_mean, running_var, nbt, n_updates, relu, scale, shift, mean, invstd);
    else { set_error("bn_apply_train: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_apply_train");
}

extern "C" int svrs_bn_bwd_reduce(const void* x, const void* dy, int dtype, int64_t M, int C, const float* scale,
                                  const float* shift, const float* mean, const float* invstd, int relu,
                                  double* sums, void* stream) {
    SVRS_CHECK_ARG(x && dy && sums && scale && shift && mean && invstd && M > 0 && c_ok(C), "bn_bwd_reduce: bad args");
    unsigned blocks; long long rpb;
    const bool w8 = wide(dtype, C);
    reduce_grid(M, C, w8 ? 8 : 4, blocks, rpb);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32 ? launch_reduce_stream<float, 4, true>(x, dy, M, C, scale, shift, mean, invstd, relu, sums, st)
        : (w8 ? launch_reduce_stream<__nv_bfloat16, 8, true>(x, dy, M, C, scale, shift, mean, invstd, relu, sums, st)
              : launch_reduce_stream<__nv_bfloat16, 4, true>(x, dy, M, C, scale, shift, mean, invstd, relu, sums, st)))
        return check_launch("bn_bwd_reduce");
    if (dtype == SVRS_F32)
        SVRS_LAUNCH((bn_reduce_kernel<float, 4, true>), blocks, 256, 0, st, (const float*)x, (const float*)dy, M, C, rpb, scale, shift, mean, invstd, relu, sums);
    else if (dtype == SVRS_BF16 && w8)
        SVRS_LAUNCH((bn_reduce_kernel<__nv_bfloat16, 8, true>), blocks, 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, M, C, rpb, scale, shift, mean, invstd, relu, sums);
    else if (dtype == SVRS_BF16)
        SVRS_LAUNCH((bn_reduce_kernel<__nv_bfloat16, 4, true>), blocks, 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, M, C, rpb, scale, shift, mean, invstd, relu, sums);
    else { set_error("bn_bwd_reduce: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_bwd_reduce");
}

extern "C" int svrs_bn_bwd_apply(const void* x, const void* dy, void* dx, int dtype, int64_t M, int64_t M_stat, int C,
                                 const float* scale, const float* shift, const float* mean, const float* invstd,
                                 const float* gamma, int relu, const double* sums, float* dgamma, float* dbeta,
                                 void* stream) {
    SVRS_CHECK_ARG(x && dy && dx && sums && scale && shift && mean && invstd && M > 0 && c_ok(C), "bn_bwd_apply: bad args");
    const bool w8 = wide(dtype, C);
    const long long Ms = M_stat > 0 ? M_stat : M;
    long long nvec = M * C / (w8 ? 8 : 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32)
        SVRS_LAUNCH((bn_bwd_apply_kernel<float, 4>), ew_grid(nvec), 256, 0, st, (const float*)x, (const float*)dy, (float*)dx, nvec, C / 4, Ms, C, scale, shift, mean, invstd, gamma, relu, sums, dgamma, dbeta);
    else if (dtype == SVRS_BF16 && w8)
        SVRS_LAUNCH((bn_bwd_apply_kernel<__nv_bfloat16, 8>), ew_grid(nvec), 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, nvec, C / 8, Ms, C, scale, shift, mean, invstd, gamma, relu, sums, dgamma, dbeta);
    else if (dtype == SVRS_BF16)
        SVRS_LAUNCH((bn_bwd_apply_kernel<__nv_bfloat16, 4>), ew_grid(nvec), 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, nvec, C / 4, Ms, C, scale, shift, mean, invstd, gamma, relu, sums, dgamma, dbeta);
    else { set_error("bn_bwd_apply: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_bwd_apply");
}
