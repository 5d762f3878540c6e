// conv_narrow.cuh - fprop / dgrad of the NARROW layers in bf16 (reduction channels K and output channels Nc in {4, 16},
// not both 16): the image-side first and last convolutions of every sub-network at full resolution
// (encoder_*.0.conv / .0.downsample, y_to_z.0.*, decoder_*.{6,7} and their dgrads; layers.py:231-236, cond_vae.py:79,142).
//
// These layers move a few hundred FLOP per pixel: they are HBM-bound streaming ops.  A 4-channel bf16 pixel is 8 bytes -
// below the 16-byte granularity of TMA boxes and UMMA shared-memory descriptors - so tcgen05 cannot take them, and the
// CUDA-core version (conv_pixel_kernel: one thread per pixel, weights broadcast from shared memory) is bound by LDS /
// FFMA issue, not by memory (16->4 at 64x64, 128 patches: 38.7 us for 25 MB = 0.65 TB/s; its 4->16 dgrad 80 us).
// Here the multi-tap GEMM runs on the warp-level tensor-core path instead (mma.sync m16n8k16, bf16 x bf16 -> fp32):
//   * one warp owns 16 consecutive output pixels per step; the A fragment is built STRAIGHT FROM GLOBAL MEMORY: the
//     k index of the MMA is permuted so that the four k values a thread owns are four consecutive channels of ONE tap
//     of ONE pixel = a single 8-byte load (K = 16: one tap per k-step, thread q takes channels 4q..4q+3;
//     K = 4: four taps per k-step, thread q takes tap q).  Neighbouring taps re-hit L1; no shared-memory staging.
//   * B fragments (the whole filter: <= 32 registers) are loaded once per warp with the same k permutation;
//   * the output columns are permuted the same way, so a thread stores its 2 (Nc = 4) or 4 (Nc = 16) channels of a pixel
//     with one 4/8-byte (bf16) or 8/16-byte (fp32) store and a pixel's channels form full 32-byte sectors;
//   * optional BatchNorm statistics (sum, sum of squares per channel) from the fp32 accumulators (svrs_conv2d_fprop_ex).
// It is legacy mma.sync on purpose: the tensor pipe is nowhere near the limiter here, the point is to get instruction
// issue out of the way of the memory system.
#pragma once

namespace svrs {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename TO> struct NarrowStore;
template <> struct NarrowStore<__nv_bfloat16> {
    __device__ __forceinline__ static void st2(__nv_bfloat16* p, float a, float b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        *reinterpret_cast<__nv_bfloat162*>(p) = h;
    }
    __device__ __forceinline__ static void st4(__nv_bfloat16* p, float a, float b, float c, float d) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b), h1 = __floats2bfloat162_rn(c, d);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t*>(&h0);
        r.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(p) = r;
    }
};
template <> struct NarrowStore<float> {
    __device__ __forceinline__ static void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
    __device__ __forceinline__ static void st4(float* p, float a, float b, float c, float d) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    }
};

constexpr int NARROW_WARPS = 4;

// K: reduction channels (4 | 16); NC: output channels (4 | 16); NTAPS: taps per problem (4 | 9 | 16), a compile-time
// constant so that ALL loads of a 16-pixel step are issued before the first MMA (a run-time tap loop serialises one
// global-load latency per k-step: measured 40-85 us per launch, the same as the CUDA-core kernel it was meant to beat).
template <int K, int NC, int NTAPS, typename TO>
__global__ void __launch_bounds__(32 * NARROW_WARPS) conv_narrow_mma_kernel(const __grid_constant__ ConvArgs a, int tiles_per_warp,
                                                                            int ohw_shift, int ow_shift) {
    pdl_entry();
    constexpr int NT = (NC + 7) / 8;                 // n-tiles of 8 columns
    constexpr int TPK = 16 / K;                      // taps per k-step
    constexpr int NKS = (NTAPS + TPK - 1) / TPK;     // k-steps of 16
    __shared__ long long s_toff[16];                 // per tap: element offset of the tap relative to the output pixel's twin
    __shared__ int s_tdy[16], s_tdx[16];
    __shared__ float s_bn[2 * 16];
    __shared__ __nv_bfloat16 s_w[NTAPS * K * NC];    // the whole filter, [tap][k][n]
    const TapGeom& g = a.g;
    const Prob& pb = g.prob[blockIdx.z];
    const __nv_bfloat16* __restrict__ in = reinterpret_cast<const __nv_bfloat16*>(a.in);
    const __nv_bfloat16* __restrict__ w = reinterpret_cast<const __nv_bfloat16*>(a.w);
    TO* __restrict__ out = reinterpret_cast<TO*>(a.out) + pb.out_off;
    if (threadIdx.x < 16) {
        const int t = threadIdx.x;
        const bool ok = t < NTAPS;
        s_tdy[t] = ok ? pb.taps[t].dy : 0;
        s_tdx[t] = ok ? pb.taps[t].dx : 0;
        s_toff[t] = ok ? pb.taps[t].in_off + (long long)pb.taps[t].dy * g.i_sy + (long long)pb.taps[t].dx * g.i_sx : 0;
    }
    if (threadIdx.x < 2 * 16) s_bn[threadIdx.x] = 0.f;
    for (int i = threadIdx.x; i < NTAPS * K * NC; i += 32 * NARROW_WARPS)
        s_w[i] = w[pb.taps[i / (K * NC)].w_off + i % (K * NC)];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, q = lane & 3;           // fragment row group / thread-in-group

    // ---- B fragments.  Fragment column c of n-tile j is output channel NT*2*(c/2) + 2*j + (c & 1) (so thread q' = c/2 of
    //      the C fragment owns NT*2 consecutive channels); k' pairs (2q, 2q+1) / (2q+8, 2q+9) are the four consecutive
    //      channels of the (tap, channel) slot described in the header
    uint32_t breg[NKS][NT][2];
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = (NT * 2) * (gq >> 1) + 2 * j + (gq & 1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int tap = K == 16 ? ks : ks * 4 + q;
                const int ch = K == 16 ? 4 * q + 2 * h : 2 * h;
                uint32_t r = 0u;
                if (tap < NTAPS && co < NC) {
                    const __nv_bfloat16* wp = s_w + (tap * K + ch) * NC + co;
                    r = (uint32_t)__bfloat16_as_ushort(wp[0]) | ((uint32_t)__bfloat16_as_ushort(wp[NC]) << 16);
                }
                breg[ks][j][h] = r;
            }
        }
    }
    // this thread's taps (one per k-step), fixed for the whole kernel
    int tdy[NKS], tdx[NKS];
    long long toff[NKS];
    bool tvalid[NKS];
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
        const int tap = K == 16 ? ks : ks * 4 + q;
        tvalid[ks] = tap < NTAPS;
        const int tt = tvalid[ks] ? tap : 0;
        tdy[ks] = s_tdy[tt];
        tdx[ks] = s_tdx[tt];
        toff[ks] = s_toff[tt] + (K == 16 ? 4 * q : 0);
    }
    // bias for this thread's output channels: NT*2 consecutive channels starting at NT*2*q
    float bz[NT * 2];
#pragma unroll
    for (int e = 0; e < NT * 2; ++e) {
        const int co = NT * 2 * q + e;
        bz[e] = (a.bias && co < NC) ? a.bias[co] : 0.f;
    }
    float bs1[NT * 2], bs2[NT * 2];
#pragma unroll
    for (int e = 0; e < NT * 2; ++e) { bs1[e] = 0.f; bs2[e] = 0.f; }

    // pixel index arithmetic in 32 bits (M < 2^31, checked by the launcher) with shifts for power-of-two maps: the 64-bit
    // divisions of the first version cost more instructions per 16-pixel step than its loads and MMAs together
    const unsigned M = (unsigned)g.N * (unsigned)g.OH * (unsigned)g.OW;
    const unsigned ohw = (unsigned)g.OH * (unsigned)g.OW;
    const unsigned mtiles = (M + 15u) / 16u;
    const unsigned wt0 = (blockIdx.x * NARROW_WARPS + warp) * (unsigned)tiles_per_warp;
    for (int it = 0; it < tiles_per_warp; ++it) {
        const unsigned mt = wt0 + it;
        if (mt >= mtiles) break;
        // the two fragment rows of this thread: pixels m0 = 16 mt + gq and m1 = m0 + 8
        long long ibase[2], obase[2];
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
            ibase[r] = (long long)n * g.i_sn + (long long)oy[r] * g.i_sy + (long long)ox[r] * g.i_sx;
            obase[r] = (long long)n * g.o_sn + (long long)oy[r] * g.o_sy + (long long)ox[r] * g.o_sx;
        }
        // ---- all loads of the step first (predicated, no branches in between) ...
        uint2 v[NKS][2];
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int iy = oy[r] + tdy[ks], ix = ox[r] + tdx[ks];
                const bool ok = live[r] && tvalid[ks] && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
                v[ks][r] = make_uint2(0u, 0u);
                if (ok) v[ks][r] = __ldg(reinterpret_cast<const uint2*>(in + ibase[r] + toff[ks]));
            }
        }
        // ---- ... then the MMAs
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            // rows gq / gq+8: a0a1 / a2a3 = k' 2q, 2q+1 ; a4a5 / a6a7 = k' 2q+8, 2q+9
            const uint32_t af[4] = {v[ks][0].x, v[ks][1].x, v[ks][0].y, v[ks][1].y};
#pragma unroll
            for (int j = 0; j < NT; ++j) mma_bf16_16816(acc[j], af, breg[ks][j][0], breg[ks][j][1]);
        }
        // ---- epilogue: fragment (row r, n-tile j, column pair 2q..2q+1) = channels NT*2*q + 2*j + {0, 1} of pixel row r
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float o[NT * 2];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                o[2 * j] = acc[j][2 * r] + bz[2 * j];
                o[2 * j + 1] = acc[j][2 * r + 1] + bz[2 * j + 1];
            }
            if (a.bn_sums && live[r]) {
#pragma unroll
                for (int e = 0; e < NT * 2; ++e) { bs1[e] += o[e]; bs2[e] = fmaf(o[e], o[e], bs2[e]); }
            }
            if (live[r] && NT * 2 * q < NC) {
                TO* po = out + obase[r] + NT * 2 * q;
                if (NT == 1) NarrowStore<TO>::st2(po, apply_act(o[0], a.act), apply_act(o[1], a.act));
                else NarrowStore<TO>::st4(po, apply_act(o[0], a.act), apply_act(o[1], a.act), apply_act(o[NT * 2 - 2], a.act),
                                          apply_act(o[NT * 2 - 1], a.act));
            }
        }
    }
    if (a.bn_sums) {                      // block-uniform
#pragma unroll
        for (int e = 0; e < NT * 2; ++e) {
            float s1 = bs1[e], s2 = bs2[e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {            // over the 8 row groups (lane bits 2..4)
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            const int co = NT * 2 * q + e;
            if (gq == 0 && co < NC) { atomicAdd(&s_bn[co], s1); atomicAdd(&s_bn[16 + co], s2); }
        }
        __syncthreads();
        if (threadIdx.x < 2 * NC) {
            const int c = threadIdx.x % NC, which = threadIdx.x / NC;
            const float s = s_bn[which * 16 + c];
            if (s != 0.f)
                atomicAdd(a.bn_sums + (size_t)((blockIdx.x + blockIdx.z) % SVRS_BN_REPLICAS) * 2 * NC + which * NC + c, (double)s);
        }
    }
}

template <int K, int NC, int NTAPS, typename TO>
static void launch_narrow_t(const ConvArgs& a, cudaStream_t st) {
    const TapGeom& g = a.g;
    const long long M = (long long)g.N * g.OH * g.OW;
    const long long mtiles = (M + 15) / 16;
    // enough CTAs for ~4 per SM, at most 8 tiles (128 pixels) per warp
    long long tpw = mtiles / ((long long)NARROW_WARPS * 4 * num_sms());
    if (tpw < 1) tpw = 1;
    if (tpw > 8) tpw = 8;
    const long long ctas = (mtiles + NARROW_WARPS * tpw - 1) / (NARROW_WARPS * tpw);
    dim3 grid((unsigned)ctas, 1, g.nprob);
    auto log2_or_neg = [](long long v) { int s = 0; while ((1ll << s) < v) ++s; return (1ll << s) == v ? s : -1; };
    SVRS_LAUNCH((conv_narrow_mma_kernel<K, NC, NTAPS, TO>), grid, 32 * NARROW_WARPS, 0, st, a, (int)tpw,
                log2_or_neg((long long)g.OH * g.OW), log2_or_neg(g.OW));
}

template <int K, int NC, typename TO>
static void launch_narrow_taps(const ConvArgs& a, cudaStream_t st) {
    const int nt = a.g.prob[0].ntaps;
    if (nt == 9) launch_narrow_t<K, NC, 9, TO>(a, st);
    else if (nt == 16) launch_narrow_t<K, NC, 16, TO>(a, st);
    else launch_narrow_t<K, NC, 4, TO>(a, st);
}

// bf16 operands, K / Nc in {4, 16} but not (16, 16); every view offset / stride a multiple of 4 elements (8-byte loads)
static bool narrow_mma_takes(const TapGeom& g, int dtype) {
    if (dtype != SVRS_BF16) return false;
    if (!((g.K == 4 || g.K == 16) && (g.Nc == 4 || g.Nc == 16)) || (g.K == 16 && g.Nc == 16)) return false;
    if ((long long)g.N * g.OH * g.OW >= (1ll << 31) - 16) return false;
    if (g.i_sn % 4 || g.i_sy % 4 || g.i_sx % 4 || g.o_sn % 4 || g.o_sy % 4 || g.o_sx % 4) return false;
    const int nt0 = g.prob[0].ntaps;
    if (nt0 != 4 && nt0 != 9 && nt0 != 16) return false;
    for (int z = 0; z < g.nprob; ++z) {
        if (g.prob[z].out_off % 4 || g.prob[z].ntaps != nt0) return false;
        for (int t = 0; t < g.prob[z].ntaps; ++t)
            if (g.prob[z].taps[t].in_off % 4) return false;
    }
    return true;
}

static void launch_narrow(const ConvArgs& a, int out_dtype, cudaStream_t st) {
    const int K = a.g.K, Nc = a.g.Nc;
    const bool f32 = out_dtype == SVRS_F32;
    if (K == 4 && Nc == 4) { if (f32) launch_narrow_taps<4, 4, float>(a, st); else launch_narrow_taps<4, 4, __nv_bfloat16>(a, st); }
    else if (K == 4 && Nc == 16) { if (f32) launch_narrow_taps<4, 16, float>(a, st); else launch_narrow_taps<4, 16, __nv_bfloat16>(a, st); }
    else { if (f32) launch_narrow_taps<16, 4, float>(a, st); else launch_narrow_taps<16, 4, __nv_bfloat16>(a, st); }
}

}  // namespace svrs
