// wgrad_narrow.cuh - weight (and bias) gradients of the NARROW bf16 layers (channels 4 <-> 16 and 4 <-> 4: the image-side
// first / last convolutions of every sub-network, layers.py:231-236, cond_vae.py:79,142) on warp-level tensor cores.
//
//   dW[a][b][tap] = sum over output pixels m of  G[m][a] * X[m + tap][b]        db[a] = sum_m G[m][a]
//
// is a GEMM whose reduction dimension is the PIXEL index, while both operands sit in memory pixel-major with the channels
// contiguous - the transpose of what mma.sync wants (k contiguous in a register pair).  The CUDA-core kernel
// (wgrad_narrow_kernel: one warp per tap, CA x CB outer products per pixel) spends its time converting and multiplying:
// 35-70 us per layer at 128 patches for 2-21 MB of operands (0.1-0.3 TB/s), IPC 1.6 at 17 % occupancy.  Here the operands
// are loaded exactly like the fprop fragments of conv_narrow.cuh (one 4/8-byte load per pixel row straight from global
// memory, permuted channel order) and transposed IN REGISTERS with movmatrix.sync.m8n8.trans - one instruction per 8x8
// block - so 16 pixels x one tap cost 2 loads + 4 movmatrix + 1 MMA instead of ~100 instructions.
//   MODE A  (CA = 4,  CB = 16)  rows = X channels (one tap per MMA),           cols = G channels (padded 4 -> 8)
//   MODE B  (CA = 16, CB = 4)   rows = G channels,                             cols = two taps x 4 X channels
//   MODE C  (CA = 4,  CB = 4)   rows = four taps x 4 X channels,               cols = G channels (padded 4 -> 8)
// The bias gradient is one more MMA against a matrix of ones.  Every warp keeps its accumulators in registers over its whole
// pixel range; a CTA folds its four warps in shared memory and issues one fp32 atomic per output element.
#pragma once
#include <cooperative_groups.h>

namespace svrs {

__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}

// fragment row i (0..15) of a 16-channel operand loaded as "thread q takes channels 4q..4q+3" -> channel
__device__ __forceinline__ int rowchan16(int i) { return i < 8 ? 4 * (i >> 1) + (i & 1) : 4 * ((i - 8) >> 1) + 2 + (i & 1); }

constexpr int WN_WARPS = 4;

template <int CA, int CB, int NTAPS>
__global__ void __launch_bounds__(32 * WN_WARPS) wgrad_narrow_mma_kernel(const __grid_constant__ WgradArgs a, float* __restrict__ db,
                                                                         int tiles_per_warp, int ohw_shift, int ow_shift) {
    pdl_entry();
    constexpr int MODE = (CA == 4 && CB == 16) ? 0 : (CA == 16 && CB == 4) ? 1 : 2;
    constexpr int NMMA = MODE == 0 ? NTAPS : MODE == 1 ? (NTAPS + 1) / 2 : (NTAPS + 3) / 4;
    __shared__ long long s_toff[16];
    __shared__ int s_tdy[16], s_tdx[16];
    __shared__ float s_red[(NMMA + 1) * 128];         // [mma][row 16][col 8], last block = bias MMA
    const TapGeom& g = a.g;
    const Prob& pb = g.prob[0];
    const __nv_bfloat16* __restrict__ G = reinterpret_cast<const __nv_bfloat16*>(a.gmat) + pb.out_off;
    const __nv_bfloat16* __restrict__ X = reinterpret_cast<const __nv_bfloat16*>(a.x);
    if (threadIdx.x < 16) {
        const int t = threadIdx.x;
        const bool ok = t < NTAPS;
        s_tdy[t] = ok ? pb.taps[t].dy : 0;
        s_tdx[t] = ok ? pb.taps[t].dx : 0;
        s_toff[t] = ok ? pb.taps[t].in_off + (long long)pb.taps[t].dy * g.i_sy + (long long)pb.taps[t].dx * g.i_sx : 0;
    }
    for (int i = threadIdx.x; i < (NMMA + 1) * 128; i += 32 * WN_WARPS) s_red[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, q = lane & 3;

    // this thread's tap for every MMA (X-side loads), fixed for the whole kernel
    int tdy[NMMA], tdx[NMMA];
    long long toff[NMMA];
    bool tvalid[NMMA];
#pragma unroll
    for (int j = 0; j < NMMA; ++j) {
        const int tap = MODE == 0 ? j : MODE == 1 ? 2 * j + (q >> 1) : 4 * j + q;
        const int ch = MODE == 0 ? 4 * q : MODE == 1 ? 2 * (q & 1) : 0;
        tvalid[j] = tap < NTAPS;
        const int tt = tvalid[j] ? tap : 0;
        tdy[j] = s_tdy[tt];
        tdx[j] = s_tdx[tt];
        toff[j] = s_toff[tt] + ch;
    }
    float acc[NMMA][4], accb[4];
#pragma unroll
    for (int j = 0; j < NMMA; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) accb[e] = 0.f;
    const uint32_t ONES = 0x3F803F80u;               // bf16 (1.0, 1.0)

    const unsigned M = (unsigned)g.N * (unsigned)g.OH * (unsigned)g.OW;
    const unsigned ohw = (unsigned)g.OH * (unsigned)g.OW;
    const unsigned mtiles = (M + 15u) / 16u;
    const unsigned wt0 = (blockIdx.x * WN_WARPS + warp) * (unsigned)tiles_per_warp;
    for (int it = 0; it < tiles_per_warp; ++it) {
        const unsigned mt = wt0 + it;
        if (mt >= mtiles) break;
        long long xbase[2], gbase[2];
        int oy[2], ox[2];
        bool live[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned m = mt * 16u + gq + 8u * r;
            live[r] = m < M;
            const unsigned mm = live[r] ? m : 0u;
            unsigned n, rem;
            if (ohw_shift >= 0) { n = mm >> ohw_shift; rem = mm & (ohw - 1u); }
            else { n = mm / ohw; rem = mm - n * ohw; }
            if (ow_shift >= 0) { oy[r] = (int)(rem >> ow_shift); ox[r] = (int)(rem & ((unsigned)g.OW - 1u)); }
            else { oy[r] = (int)(rem / (unsigned)g.OW); ox[r] = (int)rem - oy[r] * g.OW; }
            xbase[r] = (long long)n * g.i_sn + (long long)oy[r] * g.i_sy + (long long)ox[r] * g.i_sx;
            gbase[r] = (long long)n * g.o_sn + (long long)oy[r] * g.o_sy + (long long)ox[r] * g.o_sx;
        }
        // ---- all loads of the step first (predicated)
        uint2 gv[2], xv[NMMA][2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            gv[r] = make_uint2(0u, 0u);
            if (CA == 16) { if (live[r]) gv[r] = __ldg(reinterpret_cast<const uint2*>(G + gbase[r] + 4 * q)); }
            else { if (live[r] && q < 2) gv[r].x = __ldg(reinterpret_cast<const unsigned*>(G + gbase[r] + 2 * q)); }
        }
#pragma unroll
        for (int j = 0; j < NMMA; ++j) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int iy = oy[r] + tdy[j], ix = ox[r] + tdx[j];
                const bool ok = live[r] && tvalid[j] && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
                xv[j][r] = make_uint2(0u, 0u);
                if (ok) {
                    if (MODE == 1) xv[j][r].x = __ldg(reinterpret_cast<const unsigned*>(X + xbase[r] + toff[j]));
                    else xv[j][r] = __ldg(reinterpret_cast<const uint2*>(X + xbase[r] + toff[j]));
                }
            }
        }
        // ---- transposes + MMAs
        if (MODE == 1) {
            // A = G^T (16 channels x 16 pixels), shared by all tap pairs
            const uint32_t af[4] = {movm_trans(gv[0].x), movm_trans(gv[0].y), movm_trans(gv[1].x), movm_trans(gv[1].y)};
#pragma unroll
            for (int j = 0; j < NMMA; ++j) mma_bf16_16816(acc[j], af, movm_trans(xv[j][0].x), movm_trans(xv[j][1].x));
            mma_bf16_16816(accb, af, ONES, ONES);
        } else {
            // B = G (16 pixels x 4 channels, padded to 8 columns), shared by all taps
            const uint32_t b0 = movm_trans(gv[0].x), b1 = movm_trans(gv[1].x);
#pragma unroll
            for (int j = 0; j < NMMA; ++j) {
                const uint32_t af[4] = {movm_trans(xv[j][0].x), movm_trans(xv[j][0].y), movm_trans(xv[j][1].x), movm_trans(xv[j][1].y)};
                mma_bf16_16816(acc[j], af, b0, b1);
            }
            const uint32_t one4[4] = {ONES, ONES, ONES, ONES};
            mma_bf16_16816(accb, one4, b0, b1);
        }
    }
    // ---- fold the CTA's warps in shared memory: fragment (row gq | gq+8, cols 2q, 2q+1)
#pragma unroll
    for (int j = 0; j <= NMMA; ++j) {
        const float* c = j < NMMA ? acc[j] : accb;
        float* dst = s_red + j * 128;
        atomicAdd(dst + gq * 8 + 2 * q, c[0]);
        atomicAdd(dst + gq * 8 + 2 * q + 1, c[1]);
        atomicAdd(dst + (gq + 8) * 8 + 2 * q, c[2]);
        atomicAdd(dst + (gq + 8) * 8 + 2 * q + 1, c[3]);
    }
    __syncthreads();
    // ---- fold the CLUSTER through distributed shared memory, then one fp32 atomic per output element per cluster.
    // MEASURED: with one atomic per element per CTA the kernel was bound by the atomics, not by its loads - ~400 CTAs x 576
    // outputs land on 18 cache lines and same-line fp32 atomics retire at ~3 ns each (43 us for a 5 MB layer).  Rank r of
    // the cluster sums elements e = r (mod cluster size) over all ranks' shared memory and issues the atomics for them.
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned cs = cluster.num_blocks(), cr = cluster.block_rank();
    if (cs > 1) cluster.sync();
    for (unsigned e = cr + cs * threadIdx.x; e < (NMMA + 1) * 128; e += cs * 32 * WN_WARPS) {
        float v = 0.f;
        for (unsigned r = 0; r < cs; ++r) v += cs > 1 ? cluster.map_shared_rank(s_red, r)[e] : s_red[e];
        if (v == 0.f) continue;
        const int j = e >> 7, i = (e >> 3) & 15, n = e & 7;
        if (j < NMMA) {
            int ca, cb, tap;
            if (MODE == 0) { ca = n; cb = rowchan16(i); tap = j; }
            else if (MODE == 1) { ca = rowchan16(i); cb = n & 3; tap = 2 * j + (n >> 2); }
            else { ca = n; cb = i < 8 ? (i & 1) : 2 + (i & 1); tap = 4 * j + (i < 8 ? (i >> 1) : ((i - 8) >> 1)); }
            if (ca < CA && tap < NTAPS) atomicAdd(&a.dw[((long long)ca * CB + cb) * a.KK + tap], v);
        } else if (db) {
            // bias MMA: MODE A / C hold sum_m G[m][n] in every row (take row 0), MODE B sum_m G[m][rowchan16(i)] in every column
            if (MODE == 1) { if (n == 0) atomicAdd(&db[rowchan16(i)], v); }
            else if (i == 0 && n < CA) atomicAdd(&db[n], v);
        }
    }
    if (cs > 1) cluster.sync();          // nobody leaves while a peer may still read its shared memory
}

template <int CA, int CB, int NTAPS>
static void launch_wgrad_narrow_mma_t(const WgradArgs& a, float* db, cudaStream_t st) {
    const TapGeom& g = a.g;
    const long long M = (long long)g.N * g.OH * g.OW;
    const long long mtiles = (M + 15) / 16;
    // ~3 CTAs per SM in clusters of 8 (one atomic per output element per cluster)
    long long ctas = 3LL * num_sms();
    long long tpw = (mtiles + ctas * WN_WARPS - 1) / (ctas * WN_WARPS);
    if (tpw < 1) tpw = 1;
    ctas = (mtiles + WN_WARPS * tpw - 1) / (WN_WARPS * tpw);
    int cs = ctas >= 8 ? 8 : 1;
    ctas = (ctas + cs - 1) / cs * cs;
    auto log2_or_neg = [](long long v) { int s = 0; while ((1ll << s) < v) ++s; return (1ll << s) == v ? s : -1; };
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(32 * WN_WARPS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, wgrad_narrow_mma_kernel<CA, CB, NTAPS>, a, db, (int)tpw, log2_or_neg((long long)g.OH * g.OW), log2_or_neg(g.OW));
}

// bf16, one problem of 9 or 16 taps, channels (4,16) / (16,4) / (4,4), 8-byte aligned views
static bool wgrad_narrow_mma_takes(const TapGeom& g, int dtype) {
    if (dtype != SVRS_BF16 || g.nprob != 1) return false;
    const int Ca = g.Nc, Cb = g.K, nt = g.prob[0].ntaps;
    if (!((Ca == 4 && Cb == 16) || (Ca == 16 && Cb == 4) || (Ca == 4 && Cb == 4))) return false;
    if (nt != 9 && nt != 16) return false;
    if ((long long)g.N * g.OH * g.OW >= (1ll << 31) - 16) return false;
    if (g.i_sn % 4 || g.i_sy % 4 || g.i_sx % 4 || g.o_sn % 4 || g.o_sy % 4 || g.o_sx % 4 || g.prob[0].out_off % 4) return false;
    for (int t = 0; t < nt; ++t)
        if (g.prob[0].taps[t].in_off % 4) return false;
    return true;
}

static int launch_wgrad_narrow_mma(const WgradArgs& a, float* db, cudaStream_t st) {
    const int Ca = a.g.Nc, Cb = a.g.K, nt = a.g.prob[0].ntaps;
#define WN_GO(CA_, CB_) do { if (nt == 9) launch_wgrad_narrow_mma_t<CA_, CB_, 9>(a, db, st); else launch_wgrad_narrow_mma_t<CA_, CB_, 16>(a, db, st); } while (0)
    if (Ca == 4 && Cb == 16) WN_GO(4, 16);
    else if (Ca == 16 && Cb == 4) WN_GO(16, 4);
    else WN_GO(4, 4);
#undef WN_GO
    return check_launch("wgrad_narrow_mma_kernel");
}

}  // namespace svrs
