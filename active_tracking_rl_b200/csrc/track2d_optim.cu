// track2d_optim.cu -- the learner's update on one flat fp32 parameter vector, two launches:
//   (1) squared global gradient norm (block reduce -> one atomicAdd per CTA),
//   (2) torch.nn.utils.clip_grad_norm_(params, max_norm) (player_util.py:157) fused with SharedAdam.step
//       (shared_optim.py:122-175): AMSGrad with eps added AFTER the square root and the old-style bias
//       correction folded into the step size -- NOT torch.optim.Adam:
//           m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; vmax = max(vmax, v)
//           p += -(lr * sqrt(1 - b2^t) / (1 - b1^t)) * m / (sqrt(vmax) + eps)
// The reference runs ~10 ATen ops per tensor over 32 tensors per update from Python; here it is 5 reads +
// 4 writes of 4 bytes per parameter, once.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/track2d.h"
#include "track2d_adam.cuh"

namespace {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float *__restrict__ g, int64_t n, float scale, float *__restrict__ out) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    for (; i + 3 < n; i += stride) {
        float4 v = *reinterpret_cast<const float4 *>(g + i);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (i < n) // ragged tail (at most one thread lands here with fewer than 4 left)
        for (int64_t j = i; j < n; j++) acc += (g[j] * scale) * (g[j] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        acc = part[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFu, acc, o);
        if (threadIdx.x == 0) atomicAdd(out, acc);
    }
}

// CUDA-graph-safe update counter: the count lives in device memory and the bias-corrected step size
// (shared_optim.py:169-173, python floats = double) is derived from it on the device
__global__ void adam_prep_kernel(long long *step_dev, float *neg_step_out, double lr, double beta1, double beta2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long t = *step_dev + 1;
        *step_dev = t;
        double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
        *neg_step_out = (float)(-(lr * sqrt(bc2) / bc1));
    }
}

__global__ void __launch_bounds__(256) sharedadam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                         float *__restrict__ v, float *__restrict__ vmax, int64_t n, float b1, float b2,
                                                         float eps, float neg_step, float max_norm, float grad_scale,
                                                         const float *__restrict__ sqnorm, const float *__restrict__ neg_step_dev) {
    if (neg_step_dev) neg_step = *neg_step_dev;
    // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to <= 1
    float coef = grad_scale;
    if (max_norm > 0.f) {
        float total = sqrtf(*sqnorm);
        float c = max_norm / (total + 1e-6f);
        coef *= fminf(c, 1.f);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    for (; i + 3 < n; i += stride) {
        float4 P = *reinterpret_cast<float4 *>(p + i), G = *reinterpret_cast<const float4 *>(g + i);
        float4 M = *reinterpret_cast<float4 *>(m + i), V = *reinterpret_cast<float4 *>(v + i), X = *reinterpret_cast<float4 *>(vmax + i);
        adam_one(P.x, G.x * coef, M.x, V.x, X.x, b1, b2, eps, neg_step);
        adam_one(P.y, G.y * coef, M.y, V.y, X.y, b1, b2, eps, neg_step);
        adam_one(P.z, G.z * coef, M.z, V.z, X.z, b1, b2, eps, neg_step);
        adam_one(P.w, G.w * coef, M.w, V.w, X.w, b1, b2, eps, neg_step);
        *reinterpret_cast<float4 *>(p + i) = P;
        *reinterpret_cast<float4 *>(m + i) = M;
        *reinterpret_cast<float4 *>(v + i) = V;
        *reinterpret_cast<float4 *>(vmax + i) = X;
    }
    if (i < n)
        for (int64_t j = i; j < n; j++) adam_one(p[j], g[j] * coef, m[j], v[j], vmax[j], b1, b2, eps, neg_step);
}


// Agent.optimize's backward recursion (player_util.py:127-140) for E envs x 2 agents, T steps:
//   R_t   = r_t + gamma * R_{t+1}            (R_T = bootstrap value, 0 where the episode ended)
//   gae_t = gamma * tau * gae_{t+1} + r_t + gamma * V_{t+1} - V_t
// with both recursions cut where done_t = 1 (the reference never lets a rollout span episodes,
// train.py:85-88).  rewards/values/returns/gae are [T(+1)][E][2] floats, done is [T][E] bytes.
__global__ void __launch_bounds__(256) gae_kernel(const float *__restrict__ rewards, const uint8_t *__restrict__ done,
                                                  const float *__restrict__ values, float *__restrict__ returns,
                                                  float *__restrict__ gae_out, int T, int64_t E, float gamma, float tau) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // (env, agent) pair
    if (i >= 2 * E) return;
    int64_t e = i >> 1;
    float R = values[(int64_t)T * 2 * E + i];
    float vnext = R;
    float gae = 0.f;
    for (int t = T - 1; t >= 0; t--) {
        float r = rewards[(int64_t)t * 2 * E + i];
        float v = values[(int64_t)t * 2 * E + i];
        if (done[(int64_t)t * E + e]) { R = 0.f; vnext = 0.f; gae = 0.f; }
        R = gamma * R + r;
        float delta = r + gamma * vnext - v;
        gae = gae * gamma * tau + delta;
        returns[(int64_t)t * 2 * E + i] = R;
        gae_out[(int64_t)t * 2 * E + i] = gae;
        vnext = v;
    }
}

} // namespace

void t2d_set_error(const char *fmt, ...);
void t2d_count_launches(int n);

void t2d_preload_optim_kernels() {  // see track2d_peer_create: kernels that may be launched next to a waiting kernel are loaded up front
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, adam_prep_kernel);
    cudaFuncGetAttributes(&fa, sharedadam_kernel);
    cudaFuncGetAttributes(&fa, sqnorm_kernel);
}

cudaError_t t2d_launch_adam_prep(long long *step_dev, float *neg_step_out, double lr, double beta1, double beta2, cudaStream_t s) {
    adam_prep_kernel<<<1, 32, 0, s>>>(step_dev, neg_step_out, lr, beta1, beta2);
    return cudaGetLastError();
}

extern "C" int track2d_sharedadam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq,
                                       int64_t n, int64_t step, double lr, double beta1, double beta2, double eps, double max_grad_norm,
                                       double grad_scale, float *norm_scratch, int64_t *step_dev, void *stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || !max_exp_avg_sq || n <= 0 || (!step_dev && step < 1) ||
        ((max_grad_norm > 0.0 || step_dev) && !norm_scratch)) {
        t2d_set_error("track2d_sharedadam_step: bad argument");
        return T2D_E_INVALID;
    }
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)max_exp_avg_sq) & 15u) != 0) {
        t2d_set_error("track2d_sharedadam_step: buffers must be 16-byte aligned");
        return T2D_E_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    t2d_count_launches((max_grad_norm > 0.0 ? 2 : 1) + (step_dev ? 1 : 0));
    int grid = (int)((n / 4 + 255) / 256);
    if (grid < 1) grid = 1;
    if (grid > 148 * 8) grid = 148 * 8;
    if (max_grad_norm > 0.0) {
        cudaMemsetAsync(norm_scratch, 0, sizeof(float), s);
        sqnorm_kernel<<<grid, 256, 0, s>>>(grad, n, (float)grad_scale, norm_scratch);
    }
    // shared_optim.py:169-173 (python floats = double)
    float neg_step = 0.f;
    const float *neg_step_dev = nullptr;
    if (step_dev) {
        adam_prep_kernel<<<1, 32, 0, s>>>(reinterpret_cast<long long *>(step_dev), norm_scratch + 1, lr, beta1, beta2);
        neg_step_dev = norm_scratch + 1;
    } else {
        double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
        neg_step = (float)(-(lr * sqrt(bc2) / bc1));
    }
    sharedadam_kernel<<<grid, 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n, (float)beta1, (float)beta2, (float)eps,
                                            neg_step, (float)max_grad_norm, (float)grad_scale, norm_scratch, neg_step_dev);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        t2d_set_error("track2d_sharedadam_step: %s", cudaGetErrorString(err));
        return T2D_E_CUDA;
    }
    return T2D_OK;
}

extern "C" int track2d_gae_returns(const float *rewards, const uint8_t *done, const float *values, float *returns, float *gae,
                                   int32_t T, int64_t E, double gamma, double tau, void *stream) {
    if (!rewards || !done || !values || !returns || !gae || T < 1 || E < 1) {
        t2d_set_error("track2d_gae_returns: bad argument");
        return T2D_E_INVALID;
    }
    t2d_count_launches(1);
    gae_kernel<<<(unsigned)((2 * E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, done, values, returns, gae, T, E, (float)gamma, (float)tau);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        t2d_set_error("track2d_gae_returns: %s", cudaGetErrorString(err));
        return T2D_E_CUDA;
    }
    return T2D_OK;
}
