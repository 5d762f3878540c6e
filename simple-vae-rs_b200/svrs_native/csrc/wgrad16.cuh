// wgrad16.cuh - weight (and bias) gradients of the 3x3 stride-1 convolutions that PRODUCE 16 channels from 16 or 64
// (decoder tails 64 -> 16 -> 16 at full resolution, cond_vae.py:76-78,139-141, vae.py:81-83; the first conv of every second
// down_block, layers.py:231-233), bf16.
//
// wgrad_tc.cu takes these layers too, but is bound by the TMA ROW RATE on them (in-kernel clock stamps, SVRS_WG_PROF): a box
// of 16-channel pixels is 128 rows of 32 bytes, TMA delivers ~1 box row per 4 cycles per SM whatever its width, and every one
// of the nine taps is its own box -> 4.6 k cycles of row traffic per 128-pixel step against ~1 k cycles of MMA issue:
// 155 us (16->16 at 64x64, 128 patches) and 145 us (64->16) on half the GPU, 25 TFLOP/s.  Here the operands are staged in
// shared memory with plain 16-byte loads - R + 2 input rows and R gradient rows of ONE image per step, so the nine taps are
// nine shifted views of the same staged rows (each input element read once per step) - and fed to mma.sync.m16n8k16 through
// ldmatrix.x4.trans, which performs the pixel-major -> k-major transpose both operands need (k = pixel) for free
// (movmatrix, used by wgrad_narrow.cuh where pixels are only 8 bytes, issues ~16x slower).
//   D[(tap, x-channel) rows, g-channel cols] += X_tap^T * G over the pixels of the CTA's row range
//   * CB = 16: the four warps split the 16-pixel steps; CB = 64: the four warps split the four 16-channel chunks of X.
//   * accumulators (9 taps x 2 n-tiles x 4 = 72 registers per thread) live in registers over the CTA's whole range;
//     fold in shared memory, cluster (DSMEM) reduction, one fp32 atomic per weight and cluster (see wgrad_narrow.cuh).
// Row pitches are padded by 16 bytes (48 / 144 bytes per pixel) so that the eight 16-byte rows of an ldmatrix 8x8 block
// fall in distinct bank groups.
#pragma once
#include <cooperative_groups.h>

namespace svrs {

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

struct Wg16Args {
    const __nv_bfloat16* x;     // [N][H][W][CB]   layer input
    const __nv_bfloat16* g;     // [N][H][W][16]   gradient wrt the layer output
    float* dw;                  // torch layout [16][CB][3][3], fp32, atomically accumulated
    float* db;                  // [16] or null
    int N, H, W, R;             // R = output rows staged per step (H % R == 0)
    int units, units_per_cta;   // unit = (image, row block)
};

constexpr int WG16_THREADS = 128;

template <int CB>
__global__ void __launch_bounds__(WG16_THREADS) wgrad16_mma_kernel(const __grid_constant__ Wg16Args a) {
    pdl_entry();
    constexpr int XP = CB * 2 + 16;          // bytes per staged X pixel (padded)
    constexpr int GP = 16 * 2 + 16;          // bytes per staged G pixel
    constexpr int CHUNKS = CB / 16;
    extern __shared__ __align__(16) uint8_t smem[];
    const int W = a.W, R = a.R;
    const int xrow_bytes = (W + 2) * XP;
    uint8_t* xs = smem;                                      // [(R + 2) rows][(W + 2) pixels][XP]
    uint8_t* gs = smem + (size_t)(R + 2) * xrow_bytes;       // [R rows][W pixels][GP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, q = lane & 3;
    // zero the padding columns (x = -1 and x = W) of every staged row once: nothing else ever writes them
    for (int i = threadIdx.x; i < (R + 2) * 2 * (XP / 16); i += WG16_THREADS) {
        const int row = i / (2 * (XP / 16)), rem = i % (2 * (XP / 16));
        const int side = rem / (XP / 16), c16 = rem % (XP / 16);
        *reinterpret_cast<uint4*>(xs + (size_t)row * xrow_bytes + (size_t)(side ? W + 1 : 0) * XP + c16 * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    float acc[9][2][4], accb[2][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[t][j][e] = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) accb[j][e] = 0.f;
    const uint32_t xs_u = smem_u32(xs), gs_u = smem_u32(gs);
    const int blocks_per_img = a.H / R;
    const int mt_per_row = W / 16, mtiles = R * mt_per_row;
    const int u0 = blockIdx.x * a.units_per_cta;
    const int u1 = u0 + a.units_per_cta < a.units ? u0 + a.units_per_cta : a.units;
    // ldmatrix row addressing of this lane: matrix mi = lane / 8 (pixel half = mi / 2, channel half = mi % 2), row = lane % 8
    const int l_pix = ((lane >> 4) & 1) * 8 + (lane & 7), l_ch8 = (lane >> 3) & 1;
    // G: matrices (pix 0-7, ch 0-7), (pix 8-15, ch 0-7), (pix 0-7, ch 8-15), (pix 8-15, ch 8-15)
    const int g_pix = ((lane >> 3) & 1) * 8 + (lane & 7), g_ch8 = (lane >> 4) & 1;
    const uint32_t ONES = 0x3F803F80u;
    for (int u = u0; u < u1; ++u) {
        const int n = u / blocks_per_img, y0 = (u % blocks_per_img) * R;
        __syncthreads();                                     // previous step's ldmatrix reads are done
        // ---- stage X rows y0-1 .. y0+R and G rows y0 .. y0+R-1 (16-byte chunks, zero rows outside the image).  A thread keeps
        //      its 16-byte column of the pixel (c16) and walks pixels with a fixed stride: no index arithmetic per chunk (the
        //      first version spent 3x more instructions computing chunk indices than on ldmatrix + MMA)
        {
            constexpr int CPP = CB / 8;                      // 16-byte chunks per pixel
            constexpr int PPP = WG16_THREADS / CPP;          // pixels per pass of the CTA
            const int c16 = threadIdx.x % CPP, pxl = threadIdx.x / CPP;
            for (int row = 0; row < R + 2; ++row) {
                const int yy = y0 - 1 + row;
                const bool ok = yy >= 0 && yy < a.H;
                const __nv_bfloat16* src = a.x + (((long long)n * a.H + (ok ? yy : 0)) * W) * CB + c16 * 8;
                uint8_t* dst = xs + (size_t)row * xrow_bytes + XP + c16 * 16;
#pragma unroll 4
                for (int px = pxl; px < W; px += PPP) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (ok) v = __ldg(reinterpret_cast<const uint4*>(src + (long long)px * CB));
                    *reinterpret_cast<uint4*>(dst + (size_t)px * XP) = v;
                }
            }
            const int gc = threadIdx.x & 1, gpx = threadIdx.x >> 1;
            for (int row = 0; row < R; ++row) {
                const __nv_bfloat16* src = a.g + (((long long)n * a.H + y0 + row) * W) * 16 + gc * 8;
                uint8_t* dst = gs + (size_t)row * W * GP + gc * 16;
#pragma unroll 2
                for (int px = gpx; px < W; px += WG16_THREADS / 2)
                    *reinterpret_cast<uint4*>(dst + (size_t)px * GP) = __ldg(reinterpret_cast<const uint4*>(src + (long long)px * 16));
            }
        }
        __syncthreads();
        // ---- MMAs: CB = 16 -> warps split the 16-pixel steps; CB = 64 -> warps split the channel chunks
        for (int mt = (CHUNKS == 1 ? warp : 0); mt < mtiles; mt += (CHUNKS == 1 ? 4 : 1)) {
            const int rr = mt / mt_per_row, x0 = (mt - rr * mt_per_row) * 16;
            const int chunk = CHUNKS == 1 ? 0 : warp;
            uint32_t bf[4];
            ldsm_x4_trans(bf, gs_u + (uint32_t)(((rr * W) + x0 + g_pix) * GP + g_ch8 * 16));
            // bf[0], bf[1] = n-tile 0 (k 0-7, 8-15); bf[2], bf[3] = n-tile 1
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int dy = t / 3 - 1, dx = t % 3 - 1;
                uint32_t af[4];
                // A fragment registers: (ch 0-7, pix 0-7), (ch 8-15, pix 0-7), (ch 0-7, pix 8-15), (ch 8-15, pix 8-15): matrix
                // order seen by ldmatrix = lane / 8 -> (pixel half, channel half) = (mi / 2, mi % 2)
                ldsm_x4_trans(af, xs_u + (uint32_t)((rr + 1 + dy) * xrow_bytes + (x0 + dx + 1 + l_pix) * XP + chunk * 32 + l_ch8 * 16));
                mma_bf16_16816(acc[t][0], af, bf[0], bf[1]);
                mma_bf16_16816(acc[t][1], af, bf[2], bf[3]);
            }
            if (a.db && (CHUNKS == 1 || warp == 0)) {
                const uint32_t one4[4] = {ONES, ONES, ONES, ONES};
                mma_bf16_16816(accb[0], one4, bf[0], bf[1]);
                mma_bf16_16816(accb[1], one4, bf[2], bf[3]);
            }
        }
    }
    // ---- fold the CTA in shared memory: s_red[tap][x-channel CB][g-channel 16] (+ [16] bias sums)
    __syncthreads();
    float* s_red = reinterpret_cast<float*>(smem);
    constexpr int NRED = 9 * CB * 16 + 16;
    for (int i = threadIdx.x; i < NRED; i += WG16_THREADS) s_red[i] = 0.f;
    __syncthreads();
    {
        const int chunk = CHUNKS == 1 ? 0 : warp;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float* dst = s_red + (t * CB + chunk * 16) * 16 + j * 8 + 2 * q;
                atomicAdd(dst + gq * 16, acc[t][j][0]);
                atomicAdd(dst + gq * 16 + 1, acc[t][j][1]);
                atomicAdd(dst + (gq + 8) * 16, acc[t][j][2]);
                atomicAdd(dst + (gq + 8) * 16 + 1, acc[t][j][3]);
            }
        if (a.db && (CHUNKS == 1 || warp == 0) && gq == 0) {          // every row of the ones-MMA holds the column sums
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                atomicAdd(s_red + 9 * CB * 16 + j * 8 + 2 * q, accb[j][0]);
                atomicAdd(s_red + 9 * CB * 16 + j * 8 + 2 * q + 1, accb[j][1]);
            }
        }
    }
    __syncthreads();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned cs = cluster.num_blocks(), cr = cluster.block_rank();
    if (cs > 1) cluster.sync();
    for (unsigned e = cr + cs * threadIdx.x; e < (unsigned)NRED; e += cs * WG16_THREADS) {
        float v = 0.f;
        for (unsigned r = 0; r < cs; ++r) v += cs > 1 ? cluster.map_shared_rank(s_red, r)[e] : s_red[e];
        if (v == 0.f) continue;
        if (e < 9u * CB * 16u) {
            const int t = e / (CB * 16), b = (e / 16) % CB, ca = e % 16;
            atomicAdd(&a.dw[((long long)ca * CB + b) * 9 + t], v);
        } else if (a.db) {
            atomicAdd(&a.db[e - 9u * CB * 16u], v);
        }
    }
    if (cs > 1) cluster.sync();
}

static bool wgrad16_takes(int dtype, int N, int H, int W, int Cin, int Cout, int ksize) {
    if (dtype != SVRS_BF16 || ksize != 3 || Cout != 16 || !(Cin == 16 || Cin == 64) || N <= 0) return false;
    if (W % 16 != 0 || W > 64 || W < 16) return false;
    const int R = W >= 32 ? 2 : 4;
    return H % R == 0 && (long long)N * H * W < (1ll << 31);
}

static int launch_wgrad16(const void* x, const void* g, float* dw, float* db, int N, int H, int W, int Cin, cudaStream_t st) {
    Wg16Args a;
    a.x = reinterpret_cast<const __nv_bfloat16*>(x); a.g = reinterpret_cast<const __nv_bfloat16*>(g);
    a.dw = dw; a.db = db; a.N = N; a.H = H; a.W = W; a.R = W >= 32 ? 2 : 4;
    a.units = N * (H / a.R);
    // 2 CTAs per SM.  MEASURED (SVRS_WG16_CTAS_PER_SM, tools/narrow_bench.py): 4 and 8 per SM are SLOWER (16->16 at 64x64:
    // 45 / 63 / 63 us; whole step 2.79 / 2.87 / 2.98 ms) although the kernel stalls on its staging loads at 13 % warp
    // occupancy - every extra CTA pays the fold + cluster reduction + atomics of 9 * CB * 16 outputs again
    static const int per_sm = getenv("SVRS_WG16_CTAS_PER_SM") ? atoi(getenv("SVRS_WG16_CTAS_PER_SM")) : 2;
    long long ctas = (long long)per_sm * num_sms();
    if (ctas > a.units) ctas = a.units;
    a.units_per_cta = (int)((a.units + ctas - 1) / ctas);
    ctas = (a.units + a.units_per_cta - 1) / a.units_per_cta;
    const int cs = ctas >= 8 ? 8 : 1;
    ctas = (ctas + cs - 1) / cs * cs;
    const int XP = Cin * 2 + 16, GP = 48;
    size_t stage = (size_t)(a.R + 2) * (W + 2) * XP + (size_t)a.R * W * GP;
    size_t red = (size_t)(9 * Cin * 16 + 16) * sizeof(float);
    size_t smem = stage > red ? stage : red;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(wgrad16_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(wgrad16_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    if (smem > 64 * 1024) { set_error("wgrad16: staging does not fit (%zu bytes)", smem); return SVRS_E_UNSUPPORTED; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(WG16_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (Cin == 16) cudaLaunchKernelEx(&cfg, wgrad16_mma_kernel<16>, a);
    else cudaLaunchKernelEx(&cfg, wgrad16_mma_kernel<64>, a);
    return check_launch("wgrad16_mma_kernel");
}

}  // namespace svrs
