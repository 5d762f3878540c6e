// bn.cu - train/eval BatchNorm2d (+ReLU) forward and backward on NHWC activations viewed as [M][C].
// Statistics are accumulated in double (per-thread double partials, double atomics) so that the
// E[x^2]-E[x]^2 form is exact to fp32 output precision (fp32 parity mode needs 1e-5).
#include "common.cuh"

namespace svrs {

// thread t owns channel quad (t % cg) and row lane (t / cg); cg = C/4 divides 256.
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                         long long M, int C, long long rows_per_block,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                                         int relu, double* __restrict__ sums) {
    __shared__ double red[256][8];
    const int cg = C / 4;
    const int q = threadIdx.x % cg, lane = threadIdx.x / cg, lanes = 256 / cg;
    const int c = q * 4;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    float sc[4], sh[4], mu[4], is[4];
    if (BWD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; mu[j] = mean[c + j]; is[j] = invstd[c + j]; }
    }
    for (long long r = r0 + lane; r < r1; r += lanes) {
        float4 xv = ld4(x + r * C + c);
        float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        if (!BWD) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { s0[j] += (double)xs[j]; s1[j] += (double)xs[j] * (double)xs[j]; }
        } else {
            float4 gv = ld4(dy + r * C + c);
            float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float g = gs[j];
                if (relu && !(fmaf(xs[j], sc[j], sh[j]) > 0.f)) g = 0.f;
                float xh = (xs[j] - mu[j]) * is[j];
                s0[j] += (double)g;
                s1[j] += (double)g * (double)xh;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[threadIdx.x][j] = s0[j]; red[threadIdx.x][4 + j] = s1[j]; }
    __syncthreads();
    if (lane == 0) {
        for (int l = 1; l < lanes; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s0[j] += red[l * cg + q][j]; s1[j] += red[l * cg + q][4 + j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sums[c + j], s0[j]);
            atomicAdd(&sums[C + c + j], s1[j]);
        }
    }
}

__global__ void bn_finalize_train_kernel(const double* __restrict__ sums, long long M, int C,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float eps, float momentum, float* __restrict__ rmean, float* __restrict__ rvar,
                                         long long* __restrict__ nbt, int n_updates,
                                         float* __restrict__ scale, float* __restrict__ shift,
                                         float* __restrict__ mean_out, float* __restrict__ invstd_out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += n_updates;
    if (c >= C) return;
    double m = sums[c] / (double)M;
    double var = sums[C + c] / (double)M - m * m;
    if (var < 0) var = 0;
    float meanf = (float)m;
    float varf = (float)var;
    float is = 1.0f / sqrtf(varf + eps);
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float sc = g * is;
    scale[c] = sc;
    shift[c] = b - meanf * sc;
    if (mean_out) mean_out[c] = meanf;
    if (invstd_out) invstd_out[c] = is;
    if (rmean && rvar) {
        float unb = (M > 1) ? (float)(var * (double)M / (double)(M - 1)) : varf;
        float rm = rmean[c], rv = rvar[c];
        for (int u = 0; u < n_updates; ++u) {
            rm = (1.f - momentum) * rm + momentum * meanf;
            rv = (1.f - momentum) * rv + momentum * unb;
        }
        rmean[c] = rm;
        rvar[c] = rv;
    }
}

__global__ void bn_finalize_eval_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                        const float* __restrict__ rmean, const float* __restrict__ rvar,
                                        float* __restrict__ scale, float* __restrict__ shift) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float is = 1.0f / sqrtf(rvar[c] + eps);
    float sc = (gamma ? gamma[c] : 1.f) * is;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - rmean[c] * sc;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec, int cg,
                                                        const float* __restrict__ scale, const float* __restrict__ shift, int relu) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % cg) * 4;
        float4 v = ld4(x + i * 4);
        float4 sc = *reinterpret_cast<const float4*>(scale + c);
        float4 sh = *reinterpret_cast<const float4*>(shift + c);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        st4(y + i * 4, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                            long long nvec, int cg, long long M, int C,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, int relu,
                                                            const double* __restrict__ sums,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dbeta) dbeta[c] += (float)sums[c];
            if (dgamma) dgamma[c] += (float)sums[C + c];
        }
    }
    const float invM = 1.0f / (float)M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % cg) * 4;
        float4 xv = ld4(x + i * 4);
        float4 gv = ld4(dy + i * 4);
        float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float gs[4] = {gv.x, gv.y, gv.z, gv.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g = gs[j];
            if (relu && !(fmaf(xs[j], scale[c + j], shift[c + j]) > 0.f)) g = 0.f;
            float xh = (xs[j] - mean[c + j]) * invstd[c + j];
            float mg = (float)sums[c + j] * invM;
            float mgx = (float)sums[C + c + j] * invM;
            float gam = gamma ? gamma[c + j] : 1.f;
            o[j] = gam * invstd[c + j] * (g - mg - xh * mgx);
        }
        st4(dx + i * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}

static bool c_ok(int C) { return C >= 4 && C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0; }

static void reduce_grid(long long M, int C, unsigned& blocks, long long& rpb) {
    int lanes = 256 / (C / 4);
    long long b = (M + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
    long long cap = 4LL * num_sms();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    rpb = (M + b - 1) / b;
    blocks = (unsigned)((M + rpb - 1) / rpb);
}

static unsigned ew_grid(long long n) {
    long long b = (n + 255) / 256, cap = 16LL * num_sms();
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace svrs

using namespace svrs;

extern "C" int svrs_bn_stats(const void* x, int dtype, int64_t M, int C, double* sums, void* stream) {
    SVRS_CHECK_ARG(x && sums && M > 0 && c_ok(C), "bn_stats: bad args (C=%d must be 4*2^k <= 1024)", C);
    unsigned blocks; long long rpb;
    reduce_grid(M, C, blocks, rpb);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32)
        bn_reduce_kernel<float, false><<<blocks, 256, 0, st>>>((const float*)x, nullptr, M, C, rpb, nullptr, nullptr, nullptr, nullptr, 0, sums);
    else if (dtype == SVRS_BF16)
        bn_reduce_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, nullptr, M, C, rpb, nullptr, nullptr, nullptr, nullptr, 0, sums);
    else { set_error("bn_stats: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_stats");
}

extern "C" int svrs_bn_finalize_train(const double* sums, int64_t M, int C, const float* gamma, const float* beta,
                                      float eps, float momentum, float* running_mean, float* running_var,
                                      int64_t* num_batches_tracked, int n_updates,
                                      float* scale, float* shift, float* mean, float* invstd, void* stream) {
    SVRS_CHECK_ARG(sums && scale && shift && M > 0 && C > 0, "bn_finalize_train: bad args");
    bn_finalize_train_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        sums, M, C, gamma, beta, eps, momentum, running_mean, running_var, (long long*)num_batches_tracked, n_updates,
        scale, shift, mean, invstd);
    return check_launch("bn_finalize_train");
}

extern "C" int svrs_bn_finalize_eval(int C, const float* gamma, const float* beta, float eps,
                                     const float* running_mean, const float* running_var,
                                     float* scale, float* shift, void* stream) {
    SVRS_CHECK_ARG(running_mean && running_var && scale && shift && C > 0, "bn_finalize_eval: bad args");
    bn_finalize_eval_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, gamma, beta, eps, running_mean, running_var, scale, shift);
    return check_launch("bn_finalize_eval");
}

extern "C" int svrs_bn_apply(const void* x, void* y, int dtype, int64_t M, int C, const float* scale,
                             const float* shift, int relu, void* stream) {
    SVRS_CHECK_ARG(x && y && scale && shift && M > 0 && C % 4 == 0, "bn_apply: bad args");
    long long nvec = M * C / 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32)
        bn_apply_kernel<float><<<ew_grid(nvec), 256, 0, st>>>((const float*)x, (float*)y, nvec, C / 4, scale, shift, relu);
    else if (dtype == SVRS_BF16)
        bn_apply_kernel<__nv_bfloat16><<<ew_grid(nvec), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, nvec, C / 4, scale, shift, relu);
    else { set_error("bn_apply: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_apply");
}

extern "C" int svrs_bn_bwd_reduce(const void* x, const void* dy, int dtype, int64_t M, int C, const float* scale,
                                  const float* shift, const float* mean, const float* invstd, int relu,
                                  double* sums, void* stream) {
    SVRS_CHECK_ARG(x && dy && sums && scale && shift && mean && invstd && M > 0 && c_ok(C), "bn_bwd_reduce: bad args");
    unsigned blocks; long long rpb;
    reduce_grid(M, C, blocks, rpb);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32)
        bn_reduce_kernel<float, true><<<blocks, 256, 0, st>>>((const float*)x, (const float*)dy, M, C, rpb, scale, shift, mean, invstd, relu, sums);
    else if (dtype == SVRS_BF16)
        bn_reduce_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, M, C, rpb, scale, shift, mean, invstd, relu, sums);
    else { set_error("bn_bwd_reduce: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_bwd_reduce");
}

extern "C" int svrs_bn_bwd_apply(const void* x, const void* dy, void* dx, int dtype, int64_t M, int C,
                                 const float* scale, const float* shift, const float* mean, const float* invstd,
                                 const float* gamma, int relu, const double* sums, float* dgamma, float* dbeta,
                                 void* stream) {
    SVRS_CHECK_ARG(x && dy && dx && sums && scale && shift && mean && invstd && M > 0 && C % 4 == 0, "bn_bwd_apply: bad args");
    long long nvec = M * C / 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SVRS_F32)
        bn_bwd_apply_kernel<float><<<ew_grid(nvec), 256, 0, st>>>((const float*)x, (const float*)dy, (float*)dx, nvec, C / 4, M, C, scale, shift, mean, invstd, gamma, relu, sums, dgamma, dbeta);
    else if (dtype == SVRS_BF16)
        bn_bwd_apply_kernel<__nv_bfloat16><<<ew_grid(nvec), 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, nvec, C / 4, M, C, scale, shift, mean, invstd, gamma, relu, sums, dgamma, dbeta);
    else { set_error("bn_bwd_apply: bad dtype"); return SVRS_E_ARG; }
    return check_launch("bn_bwd_apply");
}
