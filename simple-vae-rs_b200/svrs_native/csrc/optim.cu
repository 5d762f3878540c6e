// optim.cu - global-norm clip + Adam on flat fp32 buffers (models/base.py:106-107, train.py:65).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace svrs {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ acc) {
    pdl_entry();
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;
    double s = 0.0;
    const bool al = ((uintptr_t)g & 15) == 0;
    long long nvec = al ? n / 4 : 0;
    for (long long i = gtid; i < nvec; i += gsize) {
        float4 v = *reinterpret_cast<const float4*>(g + i * 4);
        s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (long long i = nvec * 4 + gtid; i < n; i += gsize) s += (double)g[i] * g[i];
    __shared__ double red[8];
    s = warp_sum(s);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(acc, t);
    }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v, long long n,
                                                         const double* __restrict__ sumsq, float max_norm, float grad_scale,
                                                         float lr, float b1, float b2, float eps,
                                                         const long long* __restrict__ step_ptr) {
    pdl_entry();
    __shared__ float s_coef, s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        float coef = grad_scale;
        if (sumsq) {
            // clip_grad_norm_: total = ||g||_2 ; coef = min(1, max_norm / (total + 1e-6))
            float total = (float)sqrt(*sumsq) * grad_scale;
            float c = max_norm / (total + 1e-6f);
            coef = grad_scale * fminf(c, 1.0f);
        }
        double t = (double)(*step_ptr);
        double bc1 = 1.0 - pow((double)b1, t);
        double bc2 = 1.0 - pow((double)b2, t);
        s_coef = coef;
        s_step_size = (float)((double)lr / bc1);
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float coef = s_coef, step_size = s_step_size, bc2s = s_bc2_sqrt;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2s + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// adam_multi_kernel: the whole optimiser tail of the fused step in ONE launch over a device job table.
//   * conv-weight jobs (d1 > 0): a CTA owns a 16 (d0) x 16 (d1) x kk tile of one layer.  The gradient is read in the layout
//     the weight-gradient kernel left it in - torch [d0][d1][kk] (SIMT kernels) or the per-tap packed scratch
//     [kk][d1][d0] (tcgen05 kernels' TMA-reduce epilogue) - transposed through shared memory, Adam runs on the torch-layout
//     master weight / moments, and the updated weight goes back through shared memory into BOTH compute-dtype packs
//     [kk][d0][d1] and [kk][d1][d0] that the next step's fprop / dgrad kernels consume.
//     This replaces unpack_multi + clip_adam + pack_multi (three passes over 82 MB each) by one.
//     Full tiles of the 3x3 / 4x4 layers (all but a few hundred parameters) move m / v / p with 1-D BULK copies
//     (cp.async.bulk + mbarrier): the tile's 16 rows of each array - contiguous runs of 16*kk floats in the torch layout -
//     land in shared memory, Adam runs in place there, and the rows leave as bulk stores again while the threads write
//     the packs.  MEASURED (tools/adam_bench.py, 20.6 M parameters): register-file loads (float4 LDG, 8 in flight per
//     thread, 4 CTAs per SM) cap this kernel at 195 us = 3.4 TB/s whatever the instruction count (a version with half
//     the instructions and one with approximate div / sqrt ran at the same 195 us; reads alone, all stores removed, took
//     105-123 us) - the same ceiling bn_reduce hit before it went to bulk copies.
//   * plain jobs (d1 == 0): biases and BatchNorm affine parameters, d0 contiguous elements from `off`.
// p / g / m / v are the flat buffers; a job addresses all four at the same element offset.
// ------------------------------------------------------------------------------------------------------------------
struct AdamJob {
    long long off;      // element offset of the parameter in the flat buffers
    void* p01;          // pack [kk][d0][d1] or NULL
    void* p10;          // pack [kk][d1][d0] or NULL
    int d0, d1, kk;
    int layout;         // gradient layout: 0 torch [d0][d1][kk], 1 packed [kk][d1][d0]
    int tile0;          // first global tile of this job
    int tiles_b;        // tiles along d1 (conv jobs)
};
constexpr int AD_TA = 16, AD_TB = 16, AD_PLAIN = 2048, AD_MAX_JOBS = 512;
// shared memory of a CTA: the padded transpose tile [AD_TA][AD_TB * (kk + 1) + 1] (rounded up to 128 bytes), then the dense
// m / v / p rows of the bulk path
__host__ __device__ constexpr int ad_pad_floats(int kk) { return (AD_TA * (AD_TB * (kk + 1) + 1) + 31) / 32 * 32; }

__device__ __forceinline__ void bulk_ld_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_st_1d(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ unsigned pack_bf16x2_(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&t);
}

__device__ __forceinline__ void split_kk_(int i, int kk, int& b, int& t) {
    if (kk == 16) { b = i >> 4; t = i & 15; }
    else if (kk == 9) { b = i / 9; t = i - 9 * b; }
    else { b = i / kk; t = i - kk * b; }
}

// Full 16 x 16 x KK tile, bf16 packs.  On entry the bulk loads of the tile's m / v / p rows (dense: row a at a * 16 * KK
// floats) are in flight on `bar`.
template <int KK>
__device__ __forceinline__ void adam_tile_bulk(const AdamJob& jb, const int a0, const int b0, float* __restrict__ tile,
                                               float* __restrict__ dm, float* __restrict__ dv, float* __restrict__ dp,
                                               const uint32_t bar, float* __restrict__ p, const float* __restrict__ g,
                                               float* __restrict__ m, float* __restrict__ v, const long long row0,
                                               const long long rstride, const float coef, const float step_size,
                                               const float bc2s, const float b1, const float b2, const float eps) {
    constexpr int ROW = AD_TB * (KK + 1) + 1;
    constexpr int RUN = AD_TB * KK, RUN4 = RUN / 4;      // floats / float4 per tile row (torch layout): 144 | 256, 36 | 64
    constexpr int Q = AD_TA * RUN4;                      // float4 per array and tile: 576 | 1024
    constexpr int TRIPS = (Q + 255) / 256;               // 3 | 4
    const int d0 = jb.d0, d1 = jb.d1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool packed = jb.layout == 1;
    if (packed) {        // [kk][d1][d0] -> tile[a][b][t]: a half-warp reads the 16 consecutive d0 of one (t, b)
        const int a = lane & 15, b = 2 * warp + (lane >> 4);
        const float* gj = g + jb.off + (long long)(b0 + b) * d0 + a0 + a;
        float r[KK];
#pragma unroll
        for (int t = 0; t < KK; ++t) r[t] = __ldcs(gj + (long long)t * d1 * d0);
#pragma unroll
        for (int t = 0; t < KK; ++t) tile[a * ROW + b * (KK + 1) + t] = r[t];
    }
    float4 gi[TRIPS];
    if (!packed) {
#pragma unroll
        for (int j = 0; j < TRIPS; ++j) {
            const int q = threadIdx.x + 256 * j;
            if (q < Q) {
                const int a = q / RUN4, i4 = q - a * RUN4;
                gi[j] = __ldcs(reinterpret_cast<const float4*>(g + row0 + (long long)a * rstride + 4 * i4));
            }
        }
    }
    __syncthreads();                                     // transposed gradient tile complete
    mbar_wait(bar, 0);                                   // m / v / p rows have landed
#pragma unroll
    for (int j = 0; j < TRIPS; ++j) {
        const int q = threadIdx.x + 256 * j;
        if (q < Q) {
            const int a = q / RUN4, i4 = q - a * RUN4;
            int sl[4];
            {
                int b = (4 * i4) / KK, t = 4 * i4 - b * KK;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sl[e] = a * ROW + b * (KK + 1) + t;
                    if (++t == KK) { t = 0; ++b; }
                }
            }
            const float4 m4 = reinterpret_cast<const float4*>(dm)[q], v4 = reinterpret_cast<const float4*>(dv)[q];
            const float4 p4 = reinterpret_cast<const float4*>(dp)[q];
            float ga[4];
            if (packed) { ga[0] = tile[sl[0]]; ga[1] = tile[sl[1]]; ga[2] = tile[sl[2]]; ga[3] = tile[sl[3]]; }
            else { ga[0] = gi[j].x; ga[1] = gi[j].y; ga[2] = gi[j].z; ga[3] = gi[j].w; }
            float ma[4] = {m4.x, m4.y, m4.z, m4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w}, pa[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float gg = ga[e] * coef;
                ma[e] = b1 * ma[e] + (1.f - b1) * gg;
                va[e] = b2 * va[e] + (1.f - b2) * gg * gg;
                pa[e] = pa[e] - step_size * (ma[e] / (sqrtf(va[e]) / bc2s + eps));
                tile[sl[e]] = pa[e];
            }
            reinterpret_cast<float4*>(dm)[q] = make_float4(ma[0], ma[1], ma[2], ma[3]);
            reinterpret_cast<float4*>(dv)[q] = make_float4(va[0], va[1], va[2], va[3]);
            reinterpret_cast<float4*>(dp)[q] = make_float4(pa[0], pa[1], pa[2], pa[3]);
        }
    }
    fence_proxy_async_smem();                            // generic-proxy writes of the dense rows -> visible to the bulk stores
    __syncthreads();
    if (warp == 0) {
        if (lane < AD_TA) {
            const long long r = row0 + (long long)lane * rstride;
            bulk_st_1d(m + r, smem_u32(dm + lane * RUN), RUN * 4);
            bulk_st_1d(v + r, smem_u32(dv + lane * RUN), RUN * 4);
            bulk_st_1d(p + r, smem_u32(dp + lane * RUN), RUN * 4);
        }
        bulk_commit();
    }
    // bf16 packs from the transpose tile, 8 elements (16 bytes) per store
    const int h = threadIdx.x & 1;
    if (jb.p01) {                                        // [t][a][b]: a (t, a) row is 16 b = two 16-byte halves
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(jb.p01);
#pragma unroll
        for (int r = threadIdx.x >> 1; r < KK * AD_TA; r += 128) {
            const int t = r >> 4, a = r & 15;
            const float* src = tile + a * ROW + 8 * h * (KK + 1) + t;
            uint4 o;
            o.x = pack_bf16x2_(src[0], src[KK + 1]);
            o.y = pack_bf16x2_(src[2 * (KK + 1)], src[3 * (KK + 1)]);
            o.z = pack_bf16x2_(src[4 * (KK + 1)], src[5 * (KK + 1)]);
            o.w = pack_bf16x2_(src[6 * (KK + 1)], src[7 * (KK + 1)]);
            *reinterpret_cast<uint4*>(dst + ((long long)t * d0 + a0 + a) * d1 + b0 + 8 * h) = o;
        }
    }
    if (jb.p10) {                                        // [t][b][a]: a (t, b) row is 16 a = two 16-byte halves
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(jb.p10);
#pragma unroll
        for (int r = threadIdx.x >> 1; r < KK * AD_TB; r += 128) {
            const int t = r >> 4, b = r & 15;
            const float* src = tile + 8 * h * ROW + b * (KK + 1) + t;
            uint4 o;
            o.x = pack_bf16x2_(src[0], src[ROW]);
            o.y = pack_bf16x2_(src[2 * ROW], src[3 * ROW]);
            o.z = pack_bf16x2_(src[4 * ROW], src[5 * ROW]);
            o.w = pack_bf16x2_(src[6 * ROW], src[7 * ROW]);
            *reinterpret_cast<uint4*>(dst + ((long long)t * d1 + b0 + b) * d0 + a0 + 8 * h) = o;
        }
    }
    if (warp == 0) bulk_wait_read_0();                   // shared memory must outlive the bulk stores' reads
}

template <typename TD>
__global__ void __launch_bounds__(256, 3) adam_multi_kernel(const AdamJob* __restrict__ jobs, int njobs,
                                                          float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const double* __restrict__ sumsq, float max_norm, float grad_scale,
                                                          float lr, float b1, float b2, float eps,
                                                          const long long* __restrict__ step_ptr, const int pad_floats,
                                                          const int bulk_ok) {
    pdl_entry();
    extern __shared__ __align__(128) float tile[];
    __shared__ int s_tile0[AD_MAX_JOBS];
    __shared__ float s_coef, s_step_size, s_bc2_sqrt;
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t bar = smem_u32(&s_bar);
    for (int i = threadIdx.x; i < njobs; i += blockDim.x) s_tile0[i] = jobs[i].tile0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_tile0[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const AdamJob jb = jobs[lo];
    const int lt = blockIdx.x - jb.tile0;
    // tile geometry (conv jobs) and the bulk loads of a full tile: issued before anything else so that the clip
    // coefficient below (double-precision pow / sqrt on one thread) is computed while they fly
    int a0 = 0, b0 = 0, na = 0, nb = 0;
    long long row0 = 0, rstride = 0;
    bool fast = false;
    if (jb.d1 != 0) {
        a0 = (lt / jb.tiles_b) * AD_TA; b0 = (lt % jb.tiles_b) * AD_TB;
        na = jb.d0 - a0 < AD_TA ? jb.d0 - a0 : AD_TA; nb = jb.d1 - b0 < AD_TB ? jb.d1 - b0 : AD_TB;
        row0 = jb.off + ((long long)a0 * jb.d1 + b0) * jb.kk;       // first element of tile row a = 0 (torch layout)
        rstride = (long long)jb.d1 * jb.kk;                         // distance between tile rows
        fast = bulk_ok && sizeof(TD) == 2 && na == AD_TA && nb == AD_TB && (jb.kk == 9 || jb.kk == 16) && ((jb.d0 | jb.d1) & 7) == 0 &&
               (jb.off & 3) == 0 && (((uintptr_t)jb.p01 | (uintptr_t)jb.p10) & 15) == 0;          // CTA-uniform
    }
    const int runf = AD_TB * jb.kk;                                 // floats per row of a full tile
    float* dm = tile + pad_floats;
    float* dv = dm + AD_TA * runf;
    float* dp = dv + AD_TA * runf;
    if (fast && threadIdx.x < 32) {
        if (threadIdx.x == 0) mbar_expect_tx(bar, 3u * AD_TA * runf * 4u);
        __syncwarp();
        if (threadIdx.x < AD_TA) {
            const long long r = row0 + (long long)threadIdx.x * rstride;
            bulk_ld_1d(smem_u32(dm + threadIdx.x * runf), m + r, runf * 4u, bar);
            bulk_ld_1d(smem_u32(dv + threadIdx.x * runf), v + r, runf * 4u, bar);
            bulk_ld_1d(smem_u32(dp + threadIdx.x * runf), p + r, runf * 4u, bar);
        }
    }
    if (threadIdx.x == 32) {
        float coef = grad_scale;
        if (sumsq) {
            float total = (float)sqrt(*sumsq) * grad_scale;
            float c = max_norm / (total + 1e-6f);
            coef = grad_scale * fminf(c, 1.0f);
        }
        double t = (double)(*step_ptr);
        double bc1 = 1.0 - pow((double)b1, t);
        double bc2 = 1.0 - pow((double)b2, t);
        s_coef = coef;
        s_step_size = (float)((double)lr / bc1);
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float coef = s_coef, step_size = s_step_size, bc2s = s_bc2_sqrt;
    if (fast) {
        if (jb.kk == 9) adam_tile_bulk<9>(jb, a0, b0, tile, dm, dv, dp, bar, p, g, m, v, row0, rstride, coef, step_size, bc2s, b1, b2, eps);
        else adam_tile_bulk<16>(jb, a0, b0, tile, dm, dv, dp, bar, p, g, m, v, row0, rstride, coef, step_size, bc2s, b1, b2, eps);
        return;
    }
    auto adam = [&](long long i, float gi) -> float {
        gi *= coef;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float pn = p[i] - step_size * (mi / (sqrtf(vi) / bc2s + eps));
        p[i] = pn;
        return pn;
    };
    if (jb.d1 == 0) {                                   // plain range
        const long long i0 = jb.off + (long long)lt * AD_PLAIN;
        const long long i1 = jb.off + jb.d0 < i0 + AD_PLAIN ? jb.off + jb.d0 : i0 + AD_PLAIN;
        for (long long i = i0 + threadIdx.x; i < i1; i += 256) adam(i, g[i]);
        return;
    }
    const int kk = jb.kk, d0 = jb.d0, d1 = jb.d1;
    const int ROW = AD_TB * (kk + 1) + 1;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const float* gj = g + jb.off;
    if (jb.layout == 1) {                               // packed [kk][d1][d0]: warps read 32 consecutive d0
        if (lane < na)
            for (int t = 0; t < kk; ++t)
                for (int b = warp; b < nb; b += 8)
                    tile[lane * ROW + b * (kk + 1) + t] = gj[((long long)t * d1 + b0 + b) * d0 + a0 + lane];
        __syncthreads();
    }
    const int run = nb * kk;
    const int E = na * run;
    // torch layout: row a of the tile is a contiguous run of nb*kk floats.  When the runs are 16-byte aligned (every conv
    // layer of the models: kk = 9 | 16 with d1 % 4 == 0) p / g / m / v move as float4 - this phase is 28 of the kernel's 32
    // bytes per parameter - two independent vectors per thread and trip (8 x 16-byte loads in flight).
    if ((run & 3) == 0 && (rstride & 3) == 0 && (row0 & 3) == 0) {
        const int run4 = run >> 2;
        const int Q = na * run4;
        for (int q0 = threadIdx.x; q0 < Q; q0 += 2 * 256) {
            float4 gi[2], mi[2], vi[2], pi[2];
            long long idx[2];
            int slot[2], tt[2];
            bool on[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = q0 + u * 256;
                on[u] = qq < Q;
                slot[u] = 0; tt[u] = 0; idx[u] = row0;
                if (on[u]) {
                    const int a = qq / run4, i = (qq - a * run4) << 2;
                    int b, t;
                    split_kk_(i, kk, b, t);
                    slot[u] = a * ROW + b * (kk + 1);
                    tt[u] = t;
                    idx[u] = row0 + (long long)a * rstride + i;
                    mi[u] = *reinterpret_cast<const float4*>(m + idx[u]);
                    vi[u] = *reinterpret_cast<const float4*>(v + idx[u]);
                    pi[u] = *reinterpret_cast<const float4*>(p + idx[u]);
                    if (jb.layout != 1) gi[u] = *reinterpret_cast<const float4*>(g + idx[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!on[u]) continue;
                // shared-memory slots of the four consecutive (b, t) elements of this vector
                int sl[4];
                int sb = slot[u], st_ = tt[u];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    sl[e] = sb + st_;
                    if (++st_ == kk) { st_ = 0; sb += kk + 1; }
                }
                float ga[4];
                if (jb.layout == 1) { ga[0] = tile[sl[0]]; ga[1] = tile[sl[1]]; ga[2] = tile[sl[2]]; ga[3] = tile[sl[3]]; }
                else { ga[0] = gi[u].x; ga[1] = gi[u].y; ga[2] = gi[u].z; ga[3] = gi[u].w; }
                float ma[4] = {mi[u].x, mi[u].y, mi[u].z, mi[u].w}, va[4] = {vi[u].x, vi[u].y, vi[u].z, vi[u].w};
                float pa[4] = {pi[u].x, pi[u].y, pi[u].z, pi[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float gg = ga[e] * coef;
                    ma[e] = b1 * ma[e] + (1.f - b1) * gg;
                    va[e] = b2 * va[e] + (1.f - b2) * gg * gg;
                    pa[e] = pa[e] - step_size * (ma[e] / (sqrtf(va[e]) / bc2s + eps));
                    tile[sl[e]] = pa[e];
                }
                *reinterpret_cast<float4*>(m + idx[u]) = make_float4(ma[0], ma[1], ma[2], ma[3]);
                *reinterpret_cast<float4*>(v + idx[u]) = make_float4(va[0], va[1], va[2], va[3]);
                *reinterpret_cast<float4*>(p + idx[u]) = make_float4(pa[0], pa[1], pa[2], pa[3]);
            }
        }
    } else {
        for (int e0 = threadIdx.x; e0 < E; e0 += 4 * 256) {
            float gi[4], mi[4], vi[4], pi[4];
            long long idx[4];
            int slot[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 256;
                slot[u] = -1;
                if (e < E) {
                    const int a = e / run, i = e - a * run;
                    int b, t;
                    split_kk_(i, kk, b, t);
                    slot[u] = a * ROW + b * (kk + 1) + t;
                    idx[u] = row0 + (long long)a * rstride + i;
                    gi[u] = jb.layout == 1 ? tile[slot[u]] : g[idx[u]];
                    mi[u] = m[idx[u]]; vi[u] = v[idx[u]]; pi[u] = p[idx[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (slot[u] >= 0) {
                    const float gg = gi[u] * coef;
                    const float mn = b1 * mi[u] + (1.f - b1) * gg;
                    const float vn = b2 * vi[u] + (1.f - b2) * gg * gg;
                    const float pn = pi[u] - step_size * (mn / (sqrtf(vn) / bc2s + eps));
                    m[idx[u]] = mn; v[idx[u]] = vn; p[idx[u]] = pn;
                    tile[slot[u]] = pn;
                }
            }
        }
    }
    __syncthreads();
    TD* p01 = reinterpret_cast<TD*>(jb.p01);
    TD* p10 = reinterpret_cast<TD*>(jb.p10);
    const bool pair = sizeof(TD) == 2 && (d0 & 1) == 0 && (d1 & 1) == 0;      // bf16 packs: two elements per 4-byte store
    if (p01) {                                          // [t][a][b]: b fastest
        if (pair) {
            const int tx = threadIdx.x % (AD_TB / 2), ty = threadIdx.x / (AD_TB / 2);      // 8 lanes x 4 B = one 32-byte row
            if (2 * tx < nb)
                for (int t = 0; t < kk; ++t)
                    for (int a = ty; a < na; a += 256 / (AD_TB / 2)) {
                        const float* src = tile + a * ROW + 2 * tx * (kk + 1) + t;
                        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p01) + ((long long)t * d0 + a0 + a) * d1 + b0 + 2 * tx) =
                            __floats2bfloat162_rn(src[0], src[kk + 1]);
                    }
        } else {
            const int tx = threadIdx.x % AD_TB, ty = threadIdx.x / AD_TB;
            if (tx < nb)
                for (int t = 0; t < kk; ++t)
                    for (int a = ty; a < na; a += 256 / AD_TB)
                        p01[((long long)t * d0 + a0 + a) * d1 + b0 + tx] = Cvt<TD>::from_f(tile[a * ROW + tx * (kk + 1) + t]);
        }
    }
    if (p10) {                                          // [t][b][a]: a fastest
        if (pair) {
            const int tx = threadIdx.x % (AD_TA / 2), ty = threadIdx.x / (AD_TA / 2);      // 16 lanes x 4 B = one 64-byte row
            if (2 * tx < na)
                for (int t = 0; t < kk; ++t)
                    for (int b = ty; b < nb; b += 256 / (AD_TA / 2)) {
                        const float* src = tile + 2 * tx * ROW + b * (kk + 1) + t;
                        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p10) + ((long long)t * d1 + b0 + b) * d0 + a0 + 2 * tx) =
                            __floats2bfloat162_rn(src[0], src[ROW]);
                    }
        } else {
            if (lane < na)
                for (int t = 0; t < kk; ++t)
                    for (int b = warp; b < nb; b += 8)
                        p10[((long long)t * d1 + b0 + b) * d0 + a0 + lane] = Cvt<TD>::from_f(tile[lane * ROW + b * (kk + 1) + t]);
        }
    }
}

__global__ void step_increment_kernel(long long* s) {
    pdl_entry(); *s += 1; }

}  // namespace svrs

using namespace svrs;

extern "C" int svrs_sumsq(const float* g, int64_t n, double* acc, void* stream) {
    SVRS_CHECK_ARG(g && acc && n >= 0, "sumsq: bad args");
    if (n == 0) return 0;
    long long b = (n / 4 + 255) / 256, cap = 8LL * num_sms();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    SVRS_LAUNCH((sumsq_kernel), (unsigned)b, 256, 0, (cudaStream_t)stream, g, n, acc);
    return check_launch("sumsq");
}

extern "C" int svrs_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const double* sumsq,
                              float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps,
                              const int64_t* step_ptr, void* stream) {
    SVRS_CHECK_ARG(p && g && m && v && step_ptr && n >= 0, "clip_adam: bad args");
    if (n == 0) return 0;
    long long b = (n + 255) / 256, cap = 16LL * num_sms();
    if (b > cap) b = cap;
    SVRS_LAUNCH((clip_adam_kernel), (unsigned)b, 256, 0, (cudaStream_t)stream, p, g, m, v, n, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps,
                                                                     (const long long*)step_ptr);
    return check_launch("clip_adam");
}

extern "C" int svrs_step_increment(int64_t* step_ptr, void* stream) {
    SVRS_CHECK_ARG(step_ptr, "step_increment: null");
    SVRS_LAUNCH((step_increment_kernel), 1, 1, 0, (cudaStream_t)stream, (long long*)step_ptr);
    return check_launch("step_increment");
}

extern "C" int svrs_adam_job_bytes(void) { return (int)sizeof(svrs::AdamJob); }
extern "C" int svrs_adam_tile_rows(void) { return svrs::AD_TA; }
extern "C" int svrs_adam_tile_cols(void) { return svrs::AD_TB; }

extern "C" int svrs_adam_multi(const void* jobs, int njobs, int total_tiles, int max_kk, float* p, const float* g, float* m, float* v,
                               int pack_dtype, const double* sumsq, float max_norm, float grad_scale, float lr, float beta1,
                               float beta2, float eps, const int64_t* step_ptr, void* stream) {
    SVRS_CHECK_ARG(jobs && njobs > 0 && njobs <= svrs::AD_MAX_JOBS && total_tiles > 0 && max_kk > 0 && max_kk <= 16 && p && g && m && v && step_ptr,
                   "adam_multi: bad args (at most %d jobs)", svrs::AD_MAX_JOBS);
    // padded transpose tile + dense m / v / p rows of the bulk path (66.7 KB at kk = 16: three CTAs per SM)
    const int pad_floats = svrs::ad_pad_floats(max_kk);
    size_t smem = ((size_t)pad_floats + 3u * svrs::AD_TA * svrs::AD_TB * max_kk) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    // bulk copies need 16-byte aligned rows: the flat buffers' bases (a job's offset and row pitch are checked per tile)
    static const bool no_bulk = getenv("SVRS_ADAM_BULK") && atoi(getenv("SVRS_ADAM_BULK")) == 0;
    const int bulk_ok = !no_bulk && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(adam_multi_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(adam_multi_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(adam_multi_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(adam_multi_kernel<__nv_bfloat16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    if (pack_dtype == SVRS_F32)
        SVRS_LAUNCH((adam_multi_kernel<float>), total_tiles, 256, smem, st, (const AdamJob*)jobs, njobs, p, g, m, v, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, (const long long*)step_ptr, pad_floats, bulk_ok);
    else if (pack_dtype == SVRS_BF16)
        SVRS_LAUNCH((adam_multi_kernel<__nv_bfloat16>), total_tiles, 256, smem, st, (const AdamJob*)jobs, njobs, p, g, m, v, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, (const long long*)step_ptr, pad_floats, bulk_ok);
    else { set_error("adam_multi: bad pack dtype"); return SVRS_E_ARG; }
    return check_launch("adam_multi");
}
