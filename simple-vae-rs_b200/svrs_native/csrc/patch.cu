// patch.cu - grid-mode patching of S x S tiles into P x P patches + per-patch per-channel min-max
// normalisation (dataset.py:220-247,265-274; utils.py:4-23).  One CTA per patch; bit-exact fp32 arithmetic:
// (x - min) / ((max - min) + 1e-5f), IEEE division.
#include "common.cuh"

namespace svrs {

constexpr int MAXC = 16;

template <typename TS>
__device__ __forceinline__ float load_src(const TS* p) { return (float)(*p); }

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

}  // namespace svrs

using namespace svrs;

extern "C" int svrs_grid_patch_normalize(const void* tiles, int src_is_i16, void* dst, int dst_dtype, int nhwc,
                                         int T, int C, int S, int P, void* stream) {
    SVRS_CHECK_ARG(tiles && dst && T >= 0 && C > 0 && C <= MAXC && S > 0 && P > 0 && S % P == 0,
                   "grid_patch_normalize: bad args (C <= 16, S %% P == 0)");
    if (T == 0) return 0;
    unsigned blocks = (unsigned)(T * (S / P) * (S / P));
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
