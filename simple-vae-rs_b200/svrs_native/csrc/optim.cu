// optim.cu - global-norm clip + Adam on flat fp32 buffers (models/base.py:106-107, train.py:65).
#include "common.cuh"

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
