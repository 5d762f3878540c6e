// patch.cu - patch gather (grid mode or given crop origins) from S x S tiles + per-patch per-channel min-max
// normalisation (dataset.py:205-247,265-274; utils.py:4-23).  Bit-exact fp32 arithmetic:
// (x - min) / ((max - min) + 1e-5f), IEEE division.
//
// patch_tma_kernel: one CTA per patch.  The patch (all C channels) is fetched by ONE 3-D TMA box
// (P columns x R rows x C channel planes of the tile tensor viewed as [T*C][S][S]) into shared memory, min/max are taken
// from shared memory (the tile is read from HBM once), and up to three layouts are emitted from the same staged data:
// NCHW fp32 (the reference's API layout), NHWC fp32 (NLL target of the fused step) and NHWC bf16 (operand of the first
// conv layer).  Patches too large for shared memory (P = 256) are streamed as strips of R rows twice (the second
// pass hits L2).  Crop origins: grid mode (index = tile * (S/P)^2 + row * (S/P) + col) or an explicit device array of
// (tile, top, left) - the reference's random crop (dataset.py:205-216) with the draws made by the caller.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <string.h>

namespace svrs {

constexpr int MAXC = 16;

template <typename TS>
__device__ __forceinline__ float load_src(const TS* p) { return (float)(*p); }

// ---- fallback (no TMA constraints): one CTA per patch, plain loads, the tile region is read twice ----------------
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) grid_patch_kernel(const TS* __restrict__ tiles, TD* __restrict__ dst, int nhwc,
                                                          int C, int S, int P) {
    pdl_entry();
    __shared__ float s_mn[MAXC], s_mx[MAXC];
    __shared__ float red_mn[8], red_mx[8];
    const int per_side = S / P;
    const int per_tile = per_side * per_side;
    const int pidx = blockIdx.x;
    const int tile = pidx / per_tile, q = pidx % per_tile;
    const int row0 = (q / per_side) * P, col0 = (q % per_side) * P;
    const TS* tbase = tiles + (long long)tile * C * S * S;
    const int npix = P * P;
    for (int c = 0; c < C; ++c) {
        float mn = INFINITY, mx = -INFINITY;
        const TS* cb = tbase + (long long)c * S * S;
        for (int i = threadIdx.x; i < npix; i += blockDim.x) {
            int r = i / P, cc = i % P;
            float v = load_src(cb + (long long)(row0 + r) * S + col0 + cc);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (threadIdx.x % 32 == 0) { red_mn[threadIdx.x / 32] = mn; red_mx[threadIdx.x / 32] = mx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)blockDim.x / 32; ++w) { mn = fminf(mn, red_mn[w]); mx = fmaxf(mx, red_mx[w]); }
            s_mn[c] = mn;
            s_mx[c] = mx;
        }
        __syncthreads();
    }
    TD* obase = dst + (long long)pidx * C * npix;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        int r = i / P, cc = i % P;
        for (int c = 0; c < C; ++c) {
            float v = load_src(tbase + (long long)c * S * S + (long long)(row0 + r) * S + col0 + cc);
            float mn = s_mn[c];
            float den = (s_mx[c] - mn) + 1e-5f;
            float o = __fdiv_rn(v - mn, den);
            if (nhwc) obase[(long long)i * C + c] = Cvt<TD>::from_f(o);
            else obase[(long long)c * npix + i] = Cvt<TD>::from_f(o);
        }
    }
}

// ---- TMA gather ---------------------------------------------------------------------------------------------------
// TMA fact (measured, tools/dbg/crop_probe.py): the box start must be 16-byte aligned in GLOBAL memory, i.e. the innermost
// coordinate times the element size must be a multiple of 16 (an odd `left` raises "illegal instruction"); outer coordinates
// are free.  Random crops therefore fetch the aligned superset box [left & ~(16/es - 1), + P + 16/es) and skip the first
// left % (16/es) columns of every shared-memory row; columns past the tile edge are zero-filled by TMA and never read.
struct alignas(64) PatchParams {
    CUtensorMap map;            // tiles as (S, S, T*C), box (PW, R, C); PW = P (grid mode) or P + 16/es (random crops)
    const int* origins;         // [npatch][3] = (tile, top, left) or NULL (grid mode)
    float* out_nchw;            // [npatch][C][P][P] or NULL
    float* out_nhwc;            // [npatch][P][P][C] or NULL
    __nv_bfloat16* out_nhwc_bf16;
    int C, S, P, R, nstrips, per_side, PW, amask;
};

template <typename TS, bool PADDED>
__global__ void __launch_bounds__(256) patch_tma_kernel(const __grid_constant__ PatchParams p) {
    pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    __shared__ float s_mn[MAXC], s_mx[MAXC];
    __shared__ float red_mn[8], red_mx[8];
    __shared__ __align__(8) uint64_t bar_storage;
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    const TS* sdata = reinterpret_cast<const TS*>(smem_raw + (sbase - smem_u32(smem_raw)));
    const uint32_t bar = smem_u32(&bar_storage);
    const int C = p.C, P = p.P, R = p.R;
    const int pidx = blockIdx.x;
    int tile, top, left;
    if (p.origins) {
        tile = p.origins[3 * pidx]; top = p.origins[3 * pidx + 1]; left = p.origins[3 * pidx + 2];
    } else {
        const int per_tile = p.per_side * p.per_side;
        tile = pidx / per_tile;
        const int q = pidx % per_tile;
        top = (q / p.per_side) * P; left = (q % p.per_side) * P;
    }
    if (threadIdx.x == 0) {
        prefetch_tmap(&p.map);
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < MAXC) { s_mn[threadIdx.x] = INFINITY; s_mx[threadIdx.x] = -INFINITY; }
    __syncthreads();
    const int PW = PADDED ? p.PW : P;                      // shared-memory row pitch (elements)
    const int left_al = PADDED ? (left & ~p.amask) : left;
    const int coff = left - left_al;                       // first wanted column inside a staged row
    const uint32_t strip_bytes = (uint32_t)C * R * PW * sizeof(TS);
    const int strip_pix = R * P;
    const int plane = R * PW;                              // staged elements per channel
    // staged index of patch pixel i (row-major inside the strip)
    auto sidx = [&](int i) { return PADDED ? (i / P) * PW + (i % P) + coff : i; };
    uint32_t phase = 0;
    auto load_strip = [&](int s) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, strip_bytes);
            tma_load_3d(sbase, &p.map, bar, left_al, top + s * R, tile * C);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
    };
    // ---- pass 1: per-channel min / max
    for (int s = 0; s < p.nstrips; ++s) {
        if (s > 0) __syncthreads();            // every thread is done reading the previous strip
        load_strip(s);
        for (int c = 0; c < C; ++c) {
            float mn = INFINITY, mx = -INFINITY;
            const TS* cb = sdata + (size_t)c * plane;
            for (int i = threadIdx.x; i < strip_pix; i += 256) {
                const float v = (float)cb[sidx(i)];
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (threadIdx.x % 32 == 0) { red_mn[threadIdx.x / 32] = mn; red_mx[threadIdx.x / 32] = mx; }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < 8; ++w) { mn = fminf(mn, red_mn[w]); mx = fmaxf(mx, red_mx[w]); }
                s_mn[c] = fminf(s_mn[c], mn);
                s_mx[c] = fmaxf(s_mx[c], mx);
            }
            __syncthreads();
        }
    }
    // ---- pass 2: normalise + emit (single-strip patches are still resident in shared memory)
    const long long npix = (long long)P * P;
    for (int s = 0; s < p.nstrips; ++s) {
        if (p.nstrips > 1) { __syncthreads(); load_strip(s); }
        const long long pix0 = (long long)s * strip_pix;       // first pixel (row-major in the patch) of this strip
        if (C == 4) {
            const float mn0 = s_mn[0], mn1 = s_mn[1], mn2 = s_mn[2], mn3 = s_mn[3];
            const float d0 = (s_mx[0] - mn0) + 1e-5f, d1 = (s_mx[1] - mn1) + 1e-5f, d2 = (s_mx[2] - mn2) + 1e-5f, d3 = (s_mx[3] - mn3) + 1e-5f;
            for (int i = threadIdx.x; i < strip_pix; i += 256) {
                float4 o;
                const int si = sidx(i);
                o.x = __fdiv_rn((float)sdata[si] - mn0, d0);
                o.y = __fdiv_rn((float)sdata[plane + si] - mn1, d1);
                o.z = __fdiv_rn((float)sdata[2 * plane + si] - mn2, d2);
                o.w = __fdiv_rn((float)sdata[3 * plane + si] - mn3, d3);
                const long long px = pix0 + i;
                if (p.out_nchw) {
                    float* q = p.out_nchw + (long long)pidx * 4 * npix + px;
                    q[0] = o.x; q[npix] = o.y; q[2 * npix] = o.z; q[3 * npix] = o.w;
                }
                if (p.out_nhwc) st4(p.out_nhwc + ((long long)pidx * npix + px) * 4, o);
                if (p.out_nhwc_bf16) st4(p.out_nhwc_bf16 + ((long long)pidx * npix + px) * 4, o);
            }
        } else {
            for (int c = 0; c < C; ++c) {
                const float mn = s_mn[c], den = (s_mx[c] - mn) + 1e-5f;
                const TS* cb = sdata + (size_t)c * plane;
                for (int i = threadIdx.x; i < strip_pix; i += 256) {
                    const float o = __fdiv_rn((float)cb[sidx(i)] - mn, den);
                    const long long px = pix0 + i;
                    if (p.out_nchw) p.out_nchw[((long long)pidx * C + c) * npix + px] = o;
                    if (p.out_nhwc) p.out_nhwc[((long long)pidx * npix + px) * C + c] = o;
                    if (p.out_nhwc_bf16) p.out_nhwc_bf16[((long long)pidx * npix + px) * C + c] = __float2bfloat16_rn(o);
                }
            }
        }
    }
}

int make_plane_map_3d(CUtensorMap* m, const void* base, int elem_bytes, int S, long long planes, int box_w, int box_h, int box_c);

static bool tma_patch_ok(int es, int C, int S, int P) {
    return (S * es) % 16 == 0 && (P * es) % 16 == 0 && P <= 256 && C <= MAXC && P >= 8;
}

}  // namespace svrs

using namespace svrs;

extern "C" int svrs_patch_gather_normalize(const void* tiles, int src_is_i16, int T, int C, int S, int P,
                                           const int32_t* origins, int npatch, float* out_nchw_f32, float* out_nhwc_f32,
                                           void* out_nhwc_bf16, void* stream) {
    SVRS_CHECK_ARG(tiles && T >= 0 && C > 0 && C <= MAXC && S > 0 && P > 0 && P <= S && npatch >= 0,
                   "patch_gather_normalize: bad args (C <= 16, P <= S)");
    SVRS_CHECK_ARG(origins || (S % P == 0 && npatch == T * (S / P) * (S / P)), "patch_gather_normalize: grid mode needs S %% P == 0 and npatch == T*(S/P)^2");
    SVRS_CHECK_ARG(out_nchw_f32 || out_nhwc_f32 || out_nhwc_bf16, "patch_gather_normalize: no output");
    const int es = src_is_i16 ? 2 : 4;
    SVRS_CHECK_ARG(tma_patch_ok(es, C, S, P) && ((uintptr_t)tiles & 15) == 0,
                   "patch_gather_normalize: needs 16-byte aligned tiles, S*elem and P*elem multiples of 16 bytes, 8 <= P <= 256 (S=%d P=%d)", S, P);
    if (npatch == 0 || T == 0) return 0;
    PatchParams p;
    memset(&p, 0, sizeof(p));
    // random crops: aligned superset box (see PatchParams); grid mode origins are multiples of P, already aligned
    const bool padded = origins != nullptr;
    const int PW = padded ? P + 16 / es : P;
    SVRS_CHECK_ARG(PW <= 256, "patch_gather_normalize: random crops need P <= %d", 256 - 16 / es);
    // strip height: the largest power of two dividing P whose C planes fit in 96 KB
    int R = P;
    while ((long long)C * R * PW * es > 96 * 1024 && R % 2 == 0) R /= 2;
    SVRS_CHECK_ARG((long long)C * R * PW * es <= 96 * 1024 && P % R == 0 && R <= 256, "patch_gather_normalize: patch does not tile into strips");
    p.origins = origins;
    p.out_nchw = out_nchw_f32; p.out_nhwc = out_nhwc_f32; p.out_nhwc_bf16 = reinterpret_cast<__nv_bfloat16*>(out_nhwc_bf16);
    p.C = C; p.S = S; p.P = P; p.R = R; p.nstrips = P / R; p.per_side = S / P; p.PW = PW; p.amask = 16 / es - 1;
    int rc = make_plane_map_3d(&p.map, tiles, es, S, (long long)T * C, PW, R, C);
    if (rc) return rc;
    const size_t smem = (size_t)C * R * PW * es + 128;
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(patch_tma_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 128);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(patch_tma_kernel<short, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 128);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(patch_tma_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 128);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(patch_tma_kernel<short, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 128);
        if (e != cudaSuccess) { set_error("patch_gather_normalize: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SVRS_E_CUDA; }
        attr_set = true;
    }
    if (padded) {
        if (src_is_i16) SVRS_LAUNCH((patch_tma_kernel<short, true>), (unsigned)npatch, 256, smem, st, p);
        else SVRS_LAUNCH((patch_tma_kernel<float, true>), (unsigned)npatch, 256, smem, st, p);
    } else {
        if (src_is_i16) SVRS_LAUNCH((patch_tma_kernel<short, false>), (unsigned)npatch, 256, smem, st, p);
        else SVRS_LAUNCH((patch_tma_kernel<float, false>), (unsigned)npatch, 256, smem, st, p);
    }
    return check_launch("patch_tma_kernel");
}

extern "C" int svrs_grid_patch_normalize(const void* tiles, int src_is_i16, void* dst, int dst_dtype, int nhwc,
                                         int T, int C, int S, int P, void* stream) {
    SVRS_CHECK_ARG(tiles && dst && T >= 0 && C > 0 && C <= MAXC && S > 0 && P > 0 && S % P == 0,
                   "grid_patch_normalize: bad args (C <= 16, S %% P == 0)");
    if (T == 0) return 0;
    const int npatch = T * (S / P) * (S / P);
    const bool tma = tma_patch_ok(src_is_i16 ? 2 : 4, C, S, P) && ((uintptr_t)tiles & 15) == 0 &&
                     !(dst_dtype == SVRS_BF16 && !nhwc) && (dst_dtype == SVRS_F32 || dst_dtype == SVRS_BF16);
    if (tma)
        return svrs_patch_gather_normalize(tiles, src_is_i16, T, C, S, P, nullptr, npatch,
                                           (dst_dtype == SVRS_F32 && !nhwc) ? (float*)dst : nullptr,
                                           (dst_dtype == SVRS_F32 && nhwc) ? (float*)dst : nullptr,
                                           dst_dtype == SVRS_BF16 ? dst : nullptr, stream);
    unsigned blocks = (unsigned)npatch;
    cudaStream_t st = (cudaStream_t)stream;
    if (src_is_i16) {
        if (dst_dtype == SVRS_F32) SVRS_LAUNCH((grid_patch_kernel<short, float>), blocks, 256, 0, st, (const short*)tiles, (float*)dst, nhwc, C, S, P);
        else if (dst_dtype == SVRS_BF16) SVRS_LAUNCH((grid_patch_kernel<short, __nv_bfloat16>), blocks, 256, 0, st, (const short*)tiles, (__nv_bfloat16*)dst, nhwc, C, S, P);
        else { set_error("grid_patch_normalize: bad dtype"); return SVRS_E_ARG; }
    } else {
        if (dst_dtype == SVRS_F32) SVRS_LAUNCH((grid_patch_kernel<float, float>), blocks, 256, 0, st, (const float*)tiles, (float*)dst, nhwc, C, S, P);
        else if (dst_dtype == SVRS_BF16) SVRS_LAUNCH((grid_patch_kernel<float, __nv_bfloat16>), blocks, 256, 0, st, (const float*)tiles, (__nv_bfloat16*)dst, nhwc, C, S, P);
        else { set_error("grid_patch_normalize: bad dtype"); return SVRS_E_ARG; }
    }
    return check_launch("grid_patch_normalize");
}
