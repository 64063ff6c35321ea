// Measures the issue-to-completion rate of tcgen05.mma (1 CTA per SM, operands in shared memory) for
// kind::tf32 and kind::f16 (bf16) at N = 128 / 256.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)type << 61);
}
template <int KIND>  // 0 = tf32, 1 = bf16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t bar_a = smem_u32(&bar);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0;  // zeros are valid operands
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // K-major SWIZZLE_128B operands: A 128 rows x 128 bytes (16 KB), B N rows x 128 bytes right after it
    const uint32_t a_base = smem0, b_base = smem0 + 16384;
    const uint32_t fmt = KIND == 0 ? 2u : 1u;  // TF32 / BF16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of the 128-byte swizzle row (K = 8 tf32 or 16 bf16 each)
                mma<KIND>(tmem, umma_desc(a_base + 32 * k, 16, 1024, 2), umma_desc(b_base + 32 * k, 16, 1024, 2), idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// MODE bit 0: rotate over 4 accumulators; bit 1: rotate over 3 operand pairs (hi/lo tiles at different addresses);
// bit 2: SWIZZLE_64B tiles (64-byte rows, 2 MMAs of K = 8 per tile) instead of SWIZZLE_128B
template <int MODE>
__global__ void __launch_bounds__(128, 1) pattern_kernel(int iters, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t bar_a = smem_u32(&bar);
    for (int i = threadIdx.x; i < 140 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr bool SW64 = (MODE & 4) != 0;
    constexpr uint32_t TILE = SW64 ? 8192 : 16384;
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < (SW64 ? 2 : 4); ++k) {
#pragma unroll
                for (int acc = 0; acc < 4; ++acc) {
#pragma unroll
                    for (int pr = 0; pr < 3; ++pr) {
                        const uint32_t d = tmem + ((MODE & 1) ? acc * 128 : 0);
                        // tiles: [A0 hi, A0 lo, A1 hi, A1 lo, B0 hi, B0 lo, B1 hi, B1 lo]
                        const uint32_t ai = (acc >> 1) * 2 + ((MODE & 2) ? (pr == 0 ? 1 : 0) : 0), bi = 4 + (acc & 1) * 2 + ((MODE & 2) ? (pr == 1 ? 1 : 0) : 0);
                        const uint32_t a = smem0 + ((MODE & 2) ? ai * TILE : 0) + 32 * k, b = smem0 + ((MODE & 2) ? bi * TILE : TILE) + 32 * k;
                        if (SW64) mma<0>(d, umma_desc(a, 16, 512, 4), umma_desc(b, 16, 512, 4), idesc, 1u);
                        else mma<0>(d, umma_desc(a, 16, 1024, 2), umma_desc(b, 16, 1024, 2), idesc, 1u);
                    }
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE>
void run_pattern() {
    long long *out;
    cudaMalloc(&out, 8);
    const int smem = 160 * 1024;
    cudaFuncSetAttribute(pattern_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 500;
    pattern_kernel<MODE><<<148, 128, smem>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long clk = 0;
    cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
    const int per_it = ((MODE & 4) ? 2 : 4) * 12;
    printf("tf32 128x128x8 pattern: accumulators %s, operand tiles %s, %s: %7.1f clk per MMA   %s\n", (MODE & 1) ? "4 rotating" : "1", (MODE & 2) ? "hi/lo rotating" : "fixed",
           (MODE & 4) ? "SWIZZLE_64B" : "SWIZZLE_128B", (double)clk / (iters * per_it), cudaGetErrorString(e));
    cudaFree(out);
}

template <int KIND, int N>
void run(const char *name, int kdepth) {
    long long *out;
    cudaMalloc(&out, 8);
    const int smem = 64 * 1024;
    cudaFuncSetAttribute(rate_kernel<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int grid : {1, 148}) {
        rate_kernel<KIND, N><<<grid, 128, smem>>>(iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long clk = 0;
        cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
        const double per = (double)clk / (iters * 4.0);
        printf("%-22s grid %3d: %7.1f clk per MMA (128 x %d x %d)  = %6.0f FLOP/clk/SM   %s\n", name, grid, per, N, kdepth, 2.0 * 128 * N * kdepth / per, cudaGetErrorString(e));
    }
    cudaFree(out);
}

int main() {
    run<0, 128>("tf32 N=128", 8);
    run<0, 256>("tf32 N=256", 8);
    run<1, 128>("bf16 N=128", 16);
    run<1, 256>("bf16 N=256", 16);
    run_pattern<0>(); run_pattern<1>(); run_pattern<2>(); run_pattern<3>(); run_pattern<4>(); run_pattern<7>();
    return 0;
}
