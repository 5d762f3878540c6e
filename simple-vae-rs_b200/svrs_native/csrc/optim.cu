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
//     (cp.async.bulk + mbarrier): the tile's 16 rows of each array - contiguous runs of cols*kk floats in the torch layout -
//     land in shared memory, Adam runs in place there, and the rows leave as bulk stores again while the threads write
//     the packs (measurements in front of the kernel below).
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
constexpr int AD_TA = 16, AD_PLAIN = 2048, AD_MAX_JOBS = 512;
// tile width along d1: 16 columns, 8 for the 4x4 kernels (kk > 9) - a tile row (cols * kk floats, contiguous in the torch
// layout) is 576 / 512 bytes and a tile's dense m / v / p rows take 27 KB of shared memory either way
__host__ __device__ constexpr int ad_cols(int kk) { return kk > 9 ? 8 : 16; }
constexpr int AD_PAD_FLOATS = (AD_TA * (16 * 10 + 1) + 31) / 32 * 32;       // transpose tile [16][cols * (kk + 1) + 1], largest form
constexpr int AD_STAGE_FLOATS = 3 * AD_TA * 16 * 9;                          // dense m | v | p rows of one tile, largest form
static_assert(AD_TA * (8 * 17 + 1) <= AD_PAD_FLOATS && 3 * AD_TA * 8 * 16 <= AD_STAGE_FLOATS, "kk = 16 tiles must fit");
constexpr size_t AD_SMEM_BYTES = (size_t)(AD_PAD_FLOATS + AD_STAGE_FLOATS) * sizeof(float);      // 38 KB: five CTAs per SM fit

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

struct AdamConsts { float coef, step_size, bc2s, b1, b2, eps; };

// geometry of one tile of the global tile list (CTA-uniform)
struct AdamTile {
    AdamJob jb;
    int a0, b0, na, nb, lt;
    long long row0, rstride;     // first element of tile row a = 0 in the flat buffers (torch layout), distance between tile rows
    bool fast;                   // full tile of a kk = 9 | 16 layer, 16-byte aligned rows, bf16 packs: bulk-copy pipeline
};

template <typename TD>
__device__ __forceinline__ AdamTile adam_tile_of(const AdamJob* __restrict__ jobs, const int* s_tile0, int njobs, int tile, int bulk_ok) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_tile0[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    AdamTile T;
    T.jb = jobs[lo];
    T.lt = tile - T.jb.tile0;
    T.a0 = T.b0 = T.na = T.nb = 0; T.row0 = T.rstride = 0; T.fast = false;
    if (T.jb.d1 != 0) {
        const int cols = ad_cols(T.jb.kk);
        T.a0 = (T.lt / T.jb.tiles_b) * AD_TA; T.b0 = (T.lt % T.jb.tiles_b) * cols;
        T.na = T.jb.d0 - T.a0 < AD_TA ? T.jb.d0 - T.a0 : AD_TA; T.nb = T.jb.d1 - T.b0 < cols ? T.jb.d1 - T.b0 : cols;
        T.row0 = T.jb.off + ((long long)T.a0 * T.jb.d1 + T.b0) * T.jb.kk;
        T.rstride = (long long)T.jb.d1 * T.jb.kk;
        T.fast = bulk_ok && sizeof(TD) == 2 && T.na == AD_TA && T.nb == cols && (T.jb.kk == 9 || T.jb.kk == 16) &&
                 ((T.jb.d0 | T.jb.d1) & 7) == 0 && (T.jb.off & 3) == 0 && (((uintptr_t)T.jb.p01 | (uintptr_t)T.jb.p10) & 15) == 0;
    }
    return T;
}

// warp 0: bulk loads of a fast tile's m / v / p rows into a pipeline stage
__device__ __forceinline__ void adam_issue_loads(const AdamTile& T, float* stage, uint32_t bar, const float* __restrict__ p,
                                                 const float* __restrict__ m, const float* __restrict__ v) {
    const int lane = threadIdx.x & 31;
    const int runf = ad_cols(T.jb.kk) * T.jb.kk;                      // floats per tile row: 144 | 128
    if (lane == 0) mbar_expect_tx(bar, 3u * AD_TA * runf * 4u);
    __syncwarp();
    if (lane < AD_TA) {
        const long long r = T.row0 + (long long)lane * T.rstride;
        bulk_ld_1d(smem_u32(stage + lane * runf), m + r, runf * 4u, bar);
        bulk_ld_1d(smem_u32(stage + (AD_TA + lane) * runf), v + r, runf * 4u, bar);
        bulk_ld_1d(smem_u32(stage + (2 * AD_TA + lane) * runf), p + r, runf * 4u, bar);
    }
}

// Packed-layout gradient [kk][d1][d0] of a fast tile into registers (issued before the clip coefficient is computed, used
// after it): a half-warp reads the 16
// consecutive d0 of one (t, b) row, 16 rows per pass of the CTA.  (Torch-layout gradients - the few narrow layers whose
// weight gradients come from the mma.sync / CUDA-core kernels - are read inside the tile's own iteration.)
constexpr int AD_GREGS = 9;
template <int KK>
__device__ __forceinline__ void adam_prefetch_g(const AdamTile& T, const float* __restrict__ g, float (&gr)[AD_GREGS]) {
    constexpr int COLS = ad_cols(KK);
    constexpr int PASSES = KK * COLS / 16;                              // 9 | 8
    static_assert(PASSES <= AD_GREGS, "gradient prefetch registers");
    const int a = threadIdx.x & 15, slot = threadIdx.x >> 4;
    const float* gj = g + T.jb.off + (long long)T.b0 * T.jb.d0 + T.a0 + a;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
        const int rr = ps * 16 + slot, t = rr / COLS, b = rr % COLS;
        gr[ps] = __ldcs(gj + ((long long)t * T.jb.d1 + b) * T.jb.d0);
    }
}
// One fast tile: `stage` holds (or is receiving, tracked by `bar` / `parity`) the tile's dense m | v | p rows, `gr` its gradient
// when that is packed.  Adam runs in place in shared memory, the rows leave as bulk stores, the packs are written from the
// transpose tile.
template <int KK>
__device__ __forceinline__ void adam_tile_fast(const AdamTile& T, float* __restrict__ tile, float* __restrict__ stage,
                                               const uint32_t bar, const uint32_t parity, float (&gr)[AD_GREGS],
                                               float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                               float* __restrict__ v, const AdamConsts& c) {
    constexpr int COLS = ad_cols(KK);
    constexpr int ROW = COLS * (KK + 1) + 1;
    constexpr int RUN = COLS * KK, RUN4 = RUN / 4;       // floats / float4 per tile row: 144 | 128, 36 | 32
    constexpr int Q = AD_TA * RUN4;                      // float4 per array and tile: 576 | 512
    constexpr int TRIPS = (Q + 255) / 256;               // 3 | 2
    constexpr int PASSES = KK * COLS / 16;
    const int d0 = T.jb.d0, d1 = T.jb.d1, a0 = T.a0, b0 = T.b0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool packed = T.jb.layout == 1;
    float* dm = stage;
    float* dv = stage + AD_TA * RUN;
    float* dp = stage + 2 * AD_TA * RUN;
    float4 gi[TRIPS];
    if (packed) {        // [kk][d1][d0] -> tile[a][b][t]
        const int a = threadIdx.x & 15, slot = threadIdx.x >> 4;
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
            const int rr = ps * 16 + slot, t = rr / COLS, b = rr % COLS;
            tile[a * ROW + b * (KK + 1) + t] = gr[ps];
        }
        __syncthreads();                                 // transposed gradient tile complete
    } else {
#pragma unroll
        for (int j = 0; j < TRIPS; ++j) {
            const int q = threadIdx.x + 256 * j;
            if (q < Q) {
                const int a = q / RUN4, i4 = q - a * RUN4;
                gi[j] = __ldcs(reinterpret_cast<const float4*>(g + T.row0 + (long long)a * T.rstride + 4 * i4));
            }
        }
    }
    mbar_wait(bar, parity);                              // m / v / p rows have landed
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
                const float gg = ga[e] * c.coef;
                ma[e] = c.b1 * ma[e] + (1.f - c.b1) * gg;
                va[e] = c.b2 * va[e] + (1.f - c.b2) * gg * gg;
                pa[e] = pa[e] - c.step_size * (ma[e] / (sqrtf(va[e]) / c.bc2s + c.eps));
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
            const long long r = T.row0 + (long long)lane * T.rstride;
            bulk_st_1d(m + r, smem_u32(dm + lane * RUN), RUN * 4);
            bulk_st_1d(v + r, smem_u32(dv + lane * RUN), RUN * 4);
            bulk_st_1d(p + r, smem_u32(dp + lane * RUN), RUN * 4);
        }
        bulk_commit();
    }
    // bf16 packs from the transpose tile, 8 elements (16 bytes) per store
    if (T.jb.p01) {                                      // [t][a][b]: a (t, a) row is COLS b = COLS / 8 16-byte pieces
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(T.jb.p01);
        constexpr int PC = COLS / 8;
        const int h = threadIdx.x % PC;
#pragma unroll
        for (int r = threadIdx.x / PC; r < KK * AD_TA; r += 256 / PC) {
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
    if (T.jb.p10) {                                      // [t][b][a]: a (t, b) row is 16 a = two 16-byte halves
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(T.jb.p10);
        const int h = threadIdx.x & 1;
#pragma unroll
        for (int r = threadIdx.x >> 1; r < KK * COLS; r += 128) {
            const int t = r / COLS, b = r % COLS;
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

// Every other tile (ragged edges, the 4-channel layers, kk other than 9 / 16, fp32 packs, plain ranges): register-file path.
template <typename TD>
__device__ __noinline__ void adam_tile_generic(const AdamTile T, float* __restrict__ tile, float* __restrict__ p,
                                                  const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                  const AdamConsts c) {
    const AdamJob& jb = T.jb;
    const float coef = c.coef, step_size = c.step_size, bc2s = c.bc2s, b1 = c.b1, b2 = c.b2, eps = c.eps;
    if (jb.d1 == 0) {                                   // plain range
        const long long i0 = jb.off + (long long)T.lt * AD_PLAIN;
        const long long i1 = jb.off + jb.d0 < i0 + AD_PLAIN ? jb.off + jb.d0 : i0 + AD_PLAIN;
        for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
            const float gi = g[i] * coef;
            const float mi = b1 * m[i] + (1.f - b1) * gi;
            const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bc2s + eps));
        }
        return;
    }
    const int kk = jb.kk, d0 = jb.d0, d1 = jb.d1, a0 = T.a0, b0 = T.b0, na = T.na, nb = T.nb;
    const int TB = ad_cols(kk);
    const int ROW = TB * (kk + 1) + 1;
    const long long row0 = T.row0, rstride = T.rstride;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const float* gj = g + jb.off;
    if (jb.layout == 1) {                               // packed [kk][d1][d0]: lanes read consecutive d0
        if (lane < na)
            for (int t = 0; t < kk; ++t)
                for (int b = warp; b < nb; b += 8)
                    tile[lane * ROW + b * (kk + 1) + t] = gj[((long long)t * d1 + b0 + b) * d0 + a0 + lane];
        __syncthreads();
    }
    const int run = nb * kk;
    const int E = na * run;
    // torch layout: row a of the tile is a contiguous run of nb*kk floats; float4 when the runs are 16-byte aligned
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
        for (int e0 = threadIdx.x; e0 < E; e0 += 256) {
            const int a = e0 / run, i = e0 - a * run;
            int b, t;
            split_kk_(i, kk, b, t);
            const int slot = a * ROW + b * (kk + 1) + t;
            const long long idx = row0 + (long long)a * rstride + i;
            const float gg = (jb.layout == 1 ? tile[slot] : g[idx]) * coef;
            const float mn = b1 * m[idx] + (1.f - b1) * gg;
            const float vn = b2 * v[idx] + (1.f - b2) * gg * gg;
            const float pn = p[idx] - step_size * (mn / (sqrtf(vn) / bc2s + eps));
            m[idx] = mn; v[idx] = vn; p[idx] = pn;
            tile[slot] = pn;
        }
    }
    __syncthreads();
    TD* p01 = reinterpret_cast<TD*>(jb.p01);
    TD* p10 = reinterpret_cast<TD*>(jb.p10);
    const bool pair = sizeof(TD) == 2 && (d0 & 1) == 0 && (d1 & 1) == 0;      // bf16 packs: two elements per 4-byte store
    if (p01) {                                          // [t][a][b]: b fastest
        if (pair) {
            const int tx = threadIdx.x % (TB / 2), ty = threadIdx.x / (TB / 2);
            if (2 * tx < nb)
                for (int t = 0; t < kk; ++t)
                    for (int a = ty; a < na; a += 256 / (TB / 2)) {
                        const float* src = tile + a * ROW + 2 * tx * (kk + 1) + t;
                        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p01) + ((long long)t * d0 + a0 + a) * d1 + b0 + 2 * tx) =
                            __floats2bfloat162_rn(src[0], src[kk + 1]);
                    }
        } else {
            const int tx = threadIdx.x % TB, ty = threadIdx.x / TB;
            if (tx < nb)
                for (int t = 0; t < kk; ++t)
                    for (int a = ty; a < na; a += 256 / TB)
                        p01[((long long)t * d0 + a0 + a) * d1 + b0 + tx] = Cvt<TD>::from_f(tile[a * ROW + tx * (kk + 1) + t]);
        }
    }
    if (p10) {                                          // [t][b][a]: a fastest
        if (pair) {
            const int tx = threadIdx.x % (AD_TA / 2), ty = threadIdx.x / (AD_TA / 2);
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
    __syncthreads();                                    // the transpose tile is reused by the CTA's next tile
}

// One CTA per tile, five CTAs per SM (48 registers, 38 KB).  MEASURED (tools/adam_bench.py, 20.6 M parameters; the register-file
// version with 32 x 16 tiles and 4 CTAs per SM took 195 us = 3.4 TB/s):
//   * the same tiles through bulk copies at 3 CTAs per SM (67 KB): 195 us - and 195 us again with half the instructions or
//     with approximate div / sqrt: neither the load path nor the instruction count was the limiter, occupancy was;
//   * persistent CTAs (two per SM, 93 KB) walking the tile list through a 3-stage bulk-copy pipeline with the next tile's
//     gradient prefetched into registers: 306 us (544 us with one CTA per SM).  A tile's chain - transpose store, barrier,
//     ~850 dependent instructions of IEEE div / sqrt per thread, proxy fence, barrier, bulk stores, packs - is ~10 us long
//     and only thread-level parallelism hides it;
//   * 16 x 16 / 16 x 8 tiles, one per CTA: 4 CTAs per SM 170 us, 5 per SM 165 us = 4.0 TB/s (approximate div / sqrt would
//     give 157 us; not taken - the update stays the IEEE arithmetic of svrs_clip_adam).
template <typename TD, int OCC>
__global__ void __launch_bounds__(256, OCC) adam_multi_kernel(const AdamJob* __restrict__ jobs, int njobs,
                                                          float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const double* __restrict__ sumsq, float max_norm, float grad_scale,
                                                          float lr, float b1, float b2, float eps,
                                                          const long long* __restrict__ step_ptr, const int bulk_ok) {
    pdl_entry();
    extern __shared__ __align__(128) float smem_f[];
    float* tile = smem_f;
    float* stage = smem_f + AD_PAD_FLOATS;
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
    const AdamTile T = adam_tile_of<TD>(jobs, s_tile0, njobs, blockIdx.x, bulk_ok);
    // the tile's loads first: the clip coefficient (double-precision pow / sqrt on one thread) is computed while they fly
    if (T.fast && threadIdx.x < 32) adam_issue_loads(T, stage, bar, p, m, v);
    float gr[AD_GREGS];
    if (T.fast && T.jb.layout == 1) { if (T.jb.kk == 9) adam_prefetch_g<9>(T, g, gr); else adam_prefetch_g<16>(T, g, gr); }
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
    AdamConsts c;
    c.coef = s_coef; c.step_size = s_step_size; c.bc2s = s_bc2_sqrt; c.b1 = b1; c.b2 = b2; c.eps = eps;
    if (T.fast) {
        if (T.jb.kk == 9) adam_tile_fast<9>(T, tile, stage, bar, 0u, gr, p, g, m, v, c);
        else adam_tile_fast<16>(T, tile, stage, bar, 0u, gr, p, g, m, v, c);
    } else {
        adam_tile_generic<TD>(T, tile, p, g, m, v, c);
    }
}

__global__ void step_increment_kernel(long long* s) {
    pdl_entry(); *s += 1; }

__global__ void step_begin_kernel(long long* s, double* acc) {
    pdl_entry(); *s += 1; *acc = 0.0; }

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

extern "C" int svrs_step_begin(int64_t* step_ptr, double* norm_acc, void* stream) {
    SVRS_CHECK_ARG(step_ptr && norm_acc, "step_begin: null");
    SVRS_LAUNCH((step_begin_kernel), 1, 1, 0, (cudaStream_t)stream, (long long*)step_ptr, norm_acc);
    return check_launch("step_begin");
}

extern "C" int svrs_adam_job_bytes(void) { return (int)sizeof(svrs::AdamJob); }
extern "C" int svrs_adam_tile_rows(void) { return svrs::AD_TA; }
extern "C" int svrs_adam_tile_cols(int kk) { return svrs::ad_cols(kk); }

extern "C" int svrs_adam_multi(const void* jobs, int njobs, int total_tiles, int max_kk, float* p, const float* g, float* m, float* v,
                               int pack_dtype, const double* sumsq, float max_norm, float grad_scale, float lr, float beta1,
                               float beta2, float eps, const int64_t* step_ptr, void* stream) {
    SVRS_CHECK_ARG(jobs && njobs > 0 && njobs <= svrs::AD_MAX_JOBS && total_tiles > 0 && max_kk > 0 && max_kk <= 16 && p && g && m && v && step_ptr,
                   "adam_multi: bad args (at most %d jobs)", svrs::AD_MAX_JOBS);
    cudaStream_t st = (cudaStream_t)stream;
    // bulk copies need 16-byte aligned rows: the flat buffers' bases (a job's offset and row pitch are checked per tile)
    static const bool no_bulk = getenv("SVRS_ADAM_BULK") && atoi(getenv("SVRS_ADAM_BULK")) == 0;
    const int bulk_ok = !no_bulk && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    static const int occ = getenv("SVRS_ADAM_OCC") ? atoi(getenv("SVRS_ADAM_OCC")) : 5;      // CTAs per SM the bf16 kernel is compiled for
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(adam_multi_kernel<float, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(adam_multi_kernel<__nv_bfloat16, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(adam_multi_kernel<__nv_bfloat16, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    if (pack_dtype == SVRS_F32)
        SVRS_LAUNCH((adam_multi_kernel<float, 4>), total_tiles, 256, svrs::AD_SMEM_BYTES, st, (const AdamJob*)jobs, njobs, p, g, m, v, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, (const long long*)step_ptr, bulk_ok);
    else if (pack_dtype == SVRS_BF16 && occ == 5)
        SVRS_LAUNCH((adam_multi_kernel<__nv_bfloat16, 5>), total_tiles, 256, svrs::AD_SMEM_BYTES, st, (const AdamJob*)jobs, njobs, p, g, m, v, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, (const long long*)step_ptr, bulk_ok);
    else if (pack_dtype == SVRS_BF16)
        SVRS_LAUNCH((adam_multi_kernel<__nv_bfloat16, 4>), total_tiles, 256, svrs::AD_SMEM_BYTES, st, (const AdamJob*)jobs, njobs, p, g, m, v, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, (const long long*)step_ptr, bulk_ok);
    else { set_error("adam_multi: bad pack dtype"); return SVRS_E_ARG; }
    return check_launch("adam_multi");
}
