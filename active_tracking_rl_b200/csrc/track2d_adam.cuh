// track2d_adam.cuh -- the per-element SharedAdam update (shared_optim.py:122-175), shared by the plain optimizer kernel
// (track2d_optim.cu) and the kernel that sums the peers' gradients on the fly (track2d_peer.cu).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float &vmax, float b1, float b2, float eps, float neg_step) {
    m = m * b1 + (1.f - b1) * g;
    v = v * b2 + (1.f - b2) * g * g;
    vmax = fmaxf(vmax, v);
    float denom = sqrtf(vmax) + eps;
    p = p + (neg_step * m) / denom;
}

// advances the device-resident update counter and derives the bias-corrected step size from it (defined in track2d_optim.cu)
cudaError_t t2d_launch_adam_prep(long long *step_dev, float *neg_step_out, double lr, double beta1, double beta2, cudaStream_t s);
void t2d_preload_optim_kernels();
