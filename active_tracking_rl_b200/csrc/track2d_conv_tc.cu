// track2d_conv_tc.cu -- CNN_maze's convolution stack (perception.py:68-92) with conv2 on the 5th-generation tensor cores.
//
//   x [N][13][13] (uint8 or float32) -> conv1 3x3/s2/p1 (1->16) + ReLU -> conv2 3x3/s2/p1 (16->32) + ReLU -> y2 [N][32*4*4]
//
// conv2 is 73.7 k of the stack's 80.8 k MAC per image and is an implicit GEMM: rows = (image, output position), reduction =
// (tap, input channel) = 144, columns = 32 output channels.  A CTA (one per SM, persistent) works on tiles of 8 images = 128 rows:
//
//   warps 5..11   conv1    thread = (image, output row, channel quad): 7 x 4 outputs x 9 FMA on the CUDA cores from a zero-bordered
//                          image tile (weights in registers), written channel-last into shared memory (y1s, double-buffered) as 16-byte
//                          stores; the next tile's pixels are prefetched into registers
//   warps 12..19  gather   per tap (ki, kj): the 128 x 16 slice of the im2col matrix = one 16-byte copy per (row, 4 channels) out of y1s
//                          (zeros outside the 7x7 map), split on the fly into hi = tf32(x), lo = x - hi, stored as two K-major
//                          SWIZZLE_64B operand blocks; 4-stage ring, fence.proxy.async + mbarrier per stage
//   warp 4        MMA      per stage 2 x 3 tcgen05.mma kind::tf32 (128 x 32 x 8: A_lo B_hi, A_hi B_lo, A_hi B_hi -- fp32-accurate 3xTF32)
//                          into a TMEM accumulator (32 columns, double-buffered); the weights are split once per CTA into 9 K-major blocks
//   warps 0..3    epilogue tcgen05.ld -> + bias, ReLU -> transposed through shared memory -> the tile's 16 KB of output are contiguous
//                          in global memory and leave as full-line 16-byte stores
//
// Row order inside a tile: m = 8 * output_position + image, so the 8 rows of every UMMA core matrix are 8 images at the same
// position: the gather's quarter-warps read 8 image slices that sit 788 words apart (bank-conflict free) and write one
// swizzled 8 x 64-byte atom (conflict free by construction).  The shared-memory pipe, not the tensor pipe, bounds the kernel:
// ~2,300 wavefronts per tile against 54 MMAs x 16 cycles.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/track2d.h"
#include "track2d_tc.cuh"

void t2d_set_error(const char *fmt, ...);
void t2d_count_launches(int n);

namespace {
using namespace t2dtc;

constexpr int IMGS = 8;                        // images per tile -> 128 rows
constexpr int STAGES = 4;                      // im2col ring (one stage = one tap of one tile)
constexpr int A_TILE = 128 * 64;               // 8 KB: 128 rows x 16 fp32
constexpr int STAGE_BYTES = 2 * A_TILE;        // hi | lo
constexpr int B_BLOCK = 32 * 64;               // one tap of the weights: 32 output channels x 16 input channels
constexpr int B_BYTES = 2 * 9 * B_BLOCK;       // hi[9] | lo[9]
constexpr int Y1_IMG = 49 * 16 + 4;            // floats per image, channel-last [pos][ic]; the +4 puts image i on banks 4i
constexpr int Y1_BYTES = IMGS * Y1_IMG * 4;
constexpr int XS_IMG = 15 * 16;                // zero-bordered input image as BYTES: 15 rows of 15 (+1) pixels = 16 bytes per row
constexpr int XS_BYTES = IMGS * XS_IMG;
constexpr int OUT_IMG = 512 + 4;               // staged output row of one image
constexpr int OUT_BYTES = IMGS * OUT_IMG * 4;
constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + B_BYTES;                 // 36,864 (1024-aligned)
constexpr int OFF_Y1 = OFF_A + STAGES * STAGE_BYTES;   // 102,400
constexpr int OFF_XS = OFF_Y1 + 2 * Y1_BYTES;
constexpr int OFF_OUT = OFF_XS + 2 * XS_BYTES;
constexpr int OFF_BAR = OFF_OUT + OUT_BYTES;
constexpr int FWD_SMEM = OFF_BAR + 256 + 1024 /* alignment slack */;
// 20 warps: registers are allocated to warps in groups of 4, so 21 warps would cap a thread at 80 registers; 20 leave 96
constexpr int EPI_WARPS = 4, MMA_WARP = 4, C1_WARP0 = 5, C1_WARPS = 7, G_WARP0 = 12, G_WARPS = 8;
constexpr int FWD_THREADS = (G_WARP0 + G_WARPS) * 32;  // 640
constexpr int TMEM_COLS = 64;
static_assert(OFF_A % 1024 == 0 && OFF_Y1 % 16 == 0 && OFF_XS % 16 == 0 && OFF_OUT % 16 == 0 && OFF_BAR % 8 == 0, "shared-memory layout");
static_assert(FWD_SMEM <= 227 * 1024, "shared memory");

// The observation cells are small integers (0, 1, 2, 4): the image tile lives in shared memory as bytes, a thread fetches its
// 3 x 15 window with three 16-byte loads and expands pixels in registers (one PRMT + one FADD each, exact for 0..255).
template <typename XT>
__device__ __forceinline__ uint32_t load_px(const XT *__restrict__ x, long long xs, long long N, long long n0, int e) {
    const int img = e / 169, c = e - img * 169;
    return (e < IMGS * 169 && n0 + img < N) ? (uint32_t)x[(n0 + img) * xs + c] : 0u;
}
__device__ __forceinline__ void store_px(uint8_t *xs8, int e, uint32_t v) {
    if (e < IMGS * 169) {
        const int im = e / 169, c = e - im * 169, r = c / 13, q = c - r * 13;
        xs8[im * XS_IMG + (r + 1) * 16 + q + 1] = (uint8_t)v;
    }
}
template <int COL>
__device__ __forceinline__ float win_px(const uint4 &w) {
    const uint32_t word = COL < 4 ? w.x : (COL < 8 ? w.y : (COL < 12 ? w.z : w.w));
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | (uint32_t)(COL & 3))) - 8388608.f;  // 2^23 + b, minus 2^23
}
// y1 (conv1 + ReLU) of output row `row`, column J, channels 4 q .. 4 q + 3, from the thread's 3 x 15 pixel window
template <int J, typename W>
__device__ __forceinline__ float4 conv1_at(const uint4 (&win)[3], const W &w, const float4 &b) {
    float4 a = b;
#define T2D_TAP(KI, KJ)                                   \
    {                                                     \
        const float v = win_px<2 * J + KJ>(win[KI]);      \
        const float4 ww = w[KI * 3 + KJ];                 \
        a.x = fmaf(v, ww.x, a.x); a.y = fmaf(v, ww.y, a.y); a.z = fmaf(v, ww.z, a.z); a.w = fmaf(v, ww.w, a.w); \
    }
    T2D_TAP(0, 0) T2D_TAP(0, 1) T2D_TAP(0, 2) T2D_TAP(1, 0) T2D_TAP(1, 1) T2D_TAP(1, 2) T2D_TAP(2, 0) T2D_TAP(2, 1) T2D_TAP(2, 2)
#undef T2D_TAP
    return a;
}

template <typename XT>
__global__ void __launch_bounds__(FWD_THREADS, 1) conv_tc_fwd_kernel(const XT *__restrict__ x, long long xs, long long N, const float *__restrict__ w1,
                                                                    const float *__restrict__ b1, const float *__restrict__ w2,
                                                                    const float *__restrict__ b2, float *__restrict__ y2) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t smem0 = (raw0 + 1023u) & ~1023u;  // swizzle atoms need aligned blocks
    uint8_t *sm = smem_raw + (smem0 - raw0);
    const uint32_t bars = smem0 + OFF_BAR;
    const uint32_t bar_y1f = bars, bar_y1e = bars + 16, bar_full = bars + 32, bar_empty = bar_full + 8 * STAGES, bar_accf = bar_empty + 8 * STAGES,
                   bar_acce = bar_accf + 16, tmem_slot = bar_acce + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (N + IMGS - 1) / IMGS;

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_y1f + 8 * b, C1_WARPS * 32);
            mbar_init(bar_y1e + 8 * b, G_WARPS * 32);
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, EPI_WARPS * 32);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, G_WARPS * 32);
            mbar_init(bar_empty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    // the weights as 9 K-major operand blocks (rows = output channels, reduction = 16 input channels of one tap), split hi / lo once
    for (int i = threadIdx.x; i < 32 * 144; i += FWD_THREADS) {
        const int oc = i / 144, rem = i - oc * 144, ic = rem / 9, tap = rem - ic * 9;
        float hi, lo;
        split1(w2[i], hi, lo);
        const uint32_t off = (uint32_t)(tap * B_BLOCK) + kmajor_off(oc, ic >> 2) + (uint32_t)(ic & 3) * 4u;
        *reinterpret_cast<float *>(sm + OFF_B + off) = hi;
        *reinterpret_cast<float *>(sm + OFF_B + 9 * B_BLOCK + off) = lo;
    }
    for (int i = threadIdx.x; i < 2 * XS_BYTES / 4; i += FWD_THREADS) reinterpret_cast<uint32_t *>(sm + OFF_XS)[i] = 0u;  // borders stay zero
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + OFF_BAR + (tmem_slot - bars));

    if (warp >= G_WARP0) {
        // ===== gather: im2col slices -> UMMA operand blocks =====
        const int gt = threadIdx.x - G_WARP0 * 32;
        int ist = 0, it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(bar_y1f + 8 * buf, (uint32_t)(it >> 1) & 1u);
            const uint32_t y1 = smem0 + OFF_Y1 + buf * Y1_BYTES;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap, ++ist) {
                const int s = ist % STAGES, ki = tap / 3, kj = tap - ki * 3;
                mbar_wait(bar_empty + 8 * s, ((uint32_t)(ist / STAGES) & 1u) ^ 1u);
                const uint32_t st = smem0 + OFF_A + s * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int q = gt + 256 * j, img = q & 7, c = (q >> 3) & 3, opos = q >> 5;
                    const int r = 2 * (opos >> 2) + ki - 1, cc = 2 * (opos & 3) + kj - 1;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
                    if (r >= 0 && r < 7 && cc >= 0 && cc < 7) v = lds128(y1 + (uint32_t)(img * Y1_IMG + (r * 7 + cc) * 16 + c * 4) * 4u);
                    split4(v, hi, lo);
                    const uint32_t off = (uint32_t)(opos * 512 + img * 64 + ((c ^ (img >> 1)) << 4));
                    sts128(st + off, hi);
                    sts128(st + A_TILE + off, lo);
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                mbar_arrive(bar_full + 8 * s);
            }
            mbar_arrive(bar_y1e + 8 * buf);
        }
    } else if (warp >= C1_WARP0) {
        // ===== conv1 + ReLU on the CUDA cores =====
        // thread = (image, output row i, channel quad q): 7 positions x 4 channels per thread, weights in registers for the
        // whole kernel, the 3 x 15 pixel window in 12 registers, results leave as 16-byte channel-last stores
        const int ct = threadIdx.x - C1_WARP0 * 32;
        const int img = ct / 28, row = (ct % 28) >> 2, q = ct & 3;  // 7 warps = 8 images x 7 rows x 4 channel quads (28 = 0 mod 4)
        float4 w[9], bias;
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = make_float4(__ldg(w1 + (4 * q) * 9 + k), __ldg(w1 + (4 * q + 1) * 9 + k), __ldg(w1 + (4 * q + 2) * 9 + k), __ldg(w1 + (4 * q + 3) * 9 + k));
        bias = make_float4(__ldg(b1 + 4 * q), __ldg(b1 + 4 * q + 1), __ldg(b1 + 4 * q + 2), __ldg(b1 + 4 * q + 3));
        constexpr int PRE = (IMGS * 169 + C1_WARPS * 32 - 1) / (C1_WARPS * 32);  // 6
        uint32_t pre[PRE];
#pragma unroll
        for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, (long long)blockIdx.x * IMGS, ct + C1_WARPS * 32 * j);
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            uint8_t *xs8 = sm + OFF_XS + buf * XS_BYTES;
#pragma unroll
            for (int j = 0; j < PRE; ++j) store_px(xs8, ct + C1_WARPS * 32 * j, pre[j]);
            const long long next = tile + gridDim.x;
            if (next < n_tiles) {
#pragma unroll
                for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, next * IMGS, ct + C1_WARPS * 32 * j);
            }
            bar_sync(1, C1_WARPS * 32);
            mbar_wait(bar_y1e + 8 * buf, ((uint32_t)(it >> 1) & 1u) ^ 1u);
            if (row < 7) {
                uint4 win[3];
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) win[ki] = *reinterpret_cast<const uint4 *>(xs8 + img * XS_IMG + (2 * row + ki) * 16);
                float *yo = reinterpret_cast<float *>(sm + OFF_Y1 + buf * Y1_BYTES) + img * Y1_IMG + (row * 7) * 16 + 4 * q;
#define T2D_POS(J)                                                                                          \
    {                                                                                                       \
        const float4 a = conv1_at<J>(win, w, bias);                                                         \
        *reinterpret_cast<float4 *>(yo + J * 16) = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); \
    }
                T2D_POS(0) T2D_POS(1) T2D_POS(2) T2D_POS(3) T2D_POS(4) T2D_POS(5) T2D_POS(6)
#undef T2D_POS
            }
            mbar_arrive(bar_y1f + 8 * buf);
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = idesc_tf32(128, 32, 0, 0);
        const uint64_t dbase = kmajor_desc(0u);
        int ist = 0, it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            mbar_wait(bar_acce + 8 * abuf, ((uint32_t)(it >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)(abuf * 32);
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap, ++ist) {
                const int s = ist % STAGES;
                mbar_wait(bar_full + 8 * s, (uint32_t)(ist / STAGES) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a16 = (smem0 + OFF_A + s * STAGE_BYTES) >> 4, b16 = (smem0 + OFF_B + tap * B_BLOCK) >> 4;
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        const uint64_t a_hi = dbase + (a16 + 2 * k8), a_lo = a_hi + (A_TILE >> 4);
                        const uint64_t b_hi = dbase + (b16 + 2 * k8), b_lo = b_hi + ((9 * B_BLOCK) >> 4);
                        tc_mma_tf32(d, a_lo, b_hi, idesc, (tap > 0 || k8 > 0) ? 1u : 0u);
                        tc_mma_tf32(d, a_hi, b_lo, idesc, 1u);
                        tc_mma_tf32(d, a_hi, b_hi, idesc, 1u);
                    }
                    tc_commit(bar_empty + 8 * s);                 // the stage is free once these MMAs have read it
                    if (tap == 8) tc_commit(bar_accf + 8 * abuf);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes [32w, 32w + 32) = rows m = 8 * opos + img =====
        const int et = threadIdx.x;  // 0..127
        const int img = lane & 7, opos = warp * 4 + (lane >> 3);
        float bias[32];
#pragma unroll
        for (int oc = 0; oc < 32; ++oc) bias[oc] = __ldg(b2 + oc);
        float *outs = reinterpret_cast<float *>(sm + OFF_OUT);
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            mbar_wait_backoff(bar_accf + 8 * abuf, (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            uint32_t r[32];
            tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(abuf * 32), r);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_acce + 8 * abuf);
#pragma unroll
            for (int oc = 0; oc < 32; ++oc) outs[img * OUT_IMG + oc * 16 + opos] = fmaxf(__uint_as_float(r[oc]) + bias[oc], 0.f);
            bar_sync(2, EPI_WARPS * 32);
            const long long n0 = tile * IMGS;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i4 = et + 128 * j, im = i4 >> 7, off = (i4 & 127) * 4;
                const float4 v = *reinterpret_cast<const float4 *>(outs + im * OUT_IMG + off);
                if (n0 + im < N) *reinterpret_cast<float4 *>(y2 + (n0 + im) * 512 + off) = v;
            }
            bar_sync(2, EPI_WARPS * 32);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =====================================================================================================================
// forward, version 2: NO im2col gather.  conv1 writes y1 (split hi / lo) straight into six small arrays, one per (row parity, kj):
//     C[pr][kj][a][oj] = y1pad[2 a + pr][2 oj + kj]            (y1pad = the 7x7 map inside a zero border, 9 x 9)
// each element holding [4 channel quads][8 images][4 channels] = 512 bytes.  For tap (ki, kj) the 16 output positions then read
// elements (oi + ki/2) * 4 + oj = position + 4 * (ki / 2) of array C[ki & 1][kj]: consecutive positions, constant stride -- so the A
// operand of that tap IS a K-major, no-swizzle UMMA operand in place (8-row core matrix = 8 images x 16 bytes, SBO = 512 to the next
// position, LBO = 128 to the next channel quad), addressed by moving the descriptor's start address.  Every y1 value is stored 1.4
// times on average (x 2 for hi / lo) instead of being copied 2.9 times (x 2) out of a staging tile: a third of the shared-memory
// traffic of version 1, and the eight gather warps are gone.
//   warps 0..3   epilogue (as version 1)      warp 4   MMA      warps 5..18   conv1: warp = (output row, column half), lane = (image, channel quad)
constexpr int V2_ARR0 = 20 * 512, V2_ARR1 = 16 * 512;                  // bytes of an array with row parity 0 (5 x 4 elements) / 1 (4 x 4)
constexpr int V2_PLANE = 3 * (V2_ARR0 + V2_ARR1);                       // 55,296: one of hi / lo
constexpr int V2_OFF_B = 0, V2_OFF_A = B_BYTES, V2_OFF_XS = V2_OFF_A + 2 * V2_PLANE, V2_OFF_OUT = V2_OFF_XS + 2 * XS_BYTES,
              V2_OFF_BAR = V2_OFF_OUT + OUT_BYTES, V2_SMEM = V2_OFF_BAR + 256 + 1024;
constexpr int V2_TMEM_COLS = 128;  // two accumulator buffers of 64 columns ([A W_hi sums | A_hi W_lo])
constexpr int V2_C1_WARPS = 14, V2_THREADS = (5 + V2_C1_WARPS) * 32;    // 608: warp = (output row, column half)
static_assert(V2_OFF_A % 1024 == 0 && V2_OFF_XS % 16 == 0 && V2_OFF_OUT % 16 == 0 && V2_OFF_BAR % 8 == 0 && V2_SMEM <= 227 * 1024, "layout");
__host__ __device__ constexpr int v2_arr_base(int pr, int kj) { return pr == 0 ? kj * V2_ARR0 : 3 * V2_ARR0 + kj * V2_ARR1; }

template <typename XT>
__global__ void __launch_bounds__(V2_THREADS, 1) conv_tc_fwd2_kernel(const XT *__restrict__ x, long long xs, long long N, const float *__restrict__ w1,
                                                                    const float *__restrict__ b1, const float *__restrict__ w2,
                                                                    const float *__restrict__ b2, float *__restrict__ y2) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t smem0 = (raw0 + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (smem0 - raw0);
    const uint32_t bars = smem0 + V2_OFF_BAR;
    const uint32_t bar_af = bars, bar_ae = bars + 8, bar_accf = bars + 16, bar_acce = bars + 32, tmem_slot = bars + 48;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (N + IMGS - 1) / IMGS;

    if (threadIdx.x == 0) {
        mbar_init(bar_af, V2_C1_WARPS * 32);
        mbar_init(bar_ae, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, V2_TMEM_COLS);
    for (int i = threadIdx.x; i < 32 * 144; i += V2_THREADS) {  // the weights: 9 K-major SWIZZLE_64B blocks, hi | lo (as version 1)
        const int oc = i / 144, rem = i - oc * 144, ic = rem / 9, tap = rem - ic * 9;
        float hi, lo;
        split1(w2[i], hi, lo);
        // per tap: [hi block | lo block] = ONE 64-row operand (rows 0..31 = hi, 32..63 = lo), so that a single N = 64 MMA yields both
        // A_hi W_hi and A_hi W_lo from one read of A_hi
        const uint32_t off = (uint32_t)(2 * tap * B_BLOCK) + kmajor_off(oc, ic >> 2) + (uint32_t)(ic & 3) * 4u;
        *reinterpret_cast<float *>(sm + V2_OFF_B + off) = hi;
        *reinterpret_cast<float *>(sm + V2_OFF_B + B_BLOCK + off) = lo;
    }
    for (int i = threadIdx.x; i < (2 * V2_PLANE + 2 * XS_BYTES) / 4; i += V2_THREADS) reinterpret_cast<uint32_t *>(sm + V2_OFF_A)[i] = 0u;  // zero borders
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + V2_OFF_BAR + (tmem_slot - bars));

    if (warp >= 5) {
        // ===== conv1 + ReLU, split, direct operand stores =====
        const int ct = threadIdx.x - 5 * 32;
        const int row = ct >> 6, half = (ct >> 5) & 1, img = lane & 7, q = lane >> 3;  // a quarter-warp = 8 images of one channel quad
        float4 w[9], bias;
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = make_float4(__ldg(w1 + (4 * q) * 9 + k), __ldg(w1 + (4 * q + 1) * 9 + k), __ldg(w1 + (4 * q + 2) * 9 + k), __ldg(w1 + (4 * q + 3) * 9 + k));
        bias = make_float4(__ldg(b1 + 4 * q), __ldg(b1 + 4 * q + 1), __ldg(b1 + 4 * q + 2), __ldg(b1 + 4 * q + 3));
        constexpr int PRE = (IMGS * 169 + V2_C1_WARPS * 32 - 1) / (V2_C1_WARPS * 32);  // 4
        uint32_t pre[PRE];
#pragma unroll
        for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, (long long)blockIdx.x * IMGS, ct + V2_C1_WARPS * 32 * j);
        const int pr = (row + 1) & 1, a = (row + 1) >> 1;  // padded row r' = row + 1 = 2 a + pr
        // byte offsets of this thread's 16-byte slot inside element (a, oj = 0) of the three arrays of its row parity
        const uint32_t slot0 = smem0 + V2_OFF_A + (uint32_t)(a * 4 * 512 + q * 128 + img * 16);
        const uint32_t arr0 = slot0 + (uint32_t)(pr ? v2_arr_base(1, 0) : v2_arr_base(0, 0)), arr1 = slot0 + (uint32_t)(pr ? v2_arr_base(1, 1) : v2_arr_base(0, 1)),
                       arr2 = slot0 + (uint32_t)(pr ? v2_arr_base(1, 2) : v2_arr_base(0, 2));
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            uint8_t *xs8 = sm + V2_OFF_XS + (it & 1) * XS_BYTES;
#pragma unroll
            for (int j = 0; j < PRE; ++j) store_px(xs8, ct + V2_C1_WARPS * 32 * j, pre[j]);
            const long long next = tile + gridDim.x;
            if (next < n_tiles) {
#pragma unroll
                for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, next * IMGS, ct + V2_C1_WARPS * 32 * j);
            }
            bar_sync(1, V2_C1_WARPS * 32);
            uint4 win[3];
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) win[ki] = *reinterpret_cast<const uint4 *>(xs8 + img * XS_IMG + (2 * row + ki) * 16);
            // padded column c' = J + 1: odd -> array kj = 1, oj = J / 2; even -> kj = 0 (oj = c' / 2) and kj = 2 (oj = c' / 2 - 1)
#define T2D_POS(J)                                                                         \
    {                                                                                      \
        const float4 t = conv1_at<J>(win, w, bias);                                        \
        split4(make_float4(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f), fmaxf(t.z, 0.f), fmaxf(t.w, 0.f)), hi[J & 3], lo[J & 3]); \
    }
// (handing the six arrays over one by one, each with its own full / empty barrier, was measured and is no faster: the conv1 warps' own
// per-tile latency chain bounds the kernel, not the wait for the single operand buffer)
#define T2D_PUT(J)                                                                         \
    if ((J & 1) == 0) {                                                                    \
        sts128(arr1 + (J / 2) * 512, hi[J & 3]); sts128(arr1 + (J / 2) * 512 + V2_PLANE, lo[J & 3]); \
    } else {                                                                               \
        sts128(arr0 + ((J + 1) / 2) * 512, hi[J & 3]); sts128(arr0 + ((J + 1) / 2) * 512 + V2_PLANE, lo[J & 3]); \
        sts128(arr2 + ((J + 1) / 2 - 1) * 512, hi[J & 3]); sts128(arr2 + ((J + 1) / 2 - 1) * 512 + V2_PLANE, lo[J & 3]); \
    }
            float4 hi[4], lo[4];
            if (half == 0) {
                T2D_POS(0) T2D_POS(1) T2D_POS(2) T2D_POS(3)
                mbar_wait(bar_ae, ((uint32_t)it & 1u) ^ 1u);
                T2D_PUT(0) T2D_PUT(1) T2D_PUT(2) T2D_PUT(3)
            } else {
                T2D_POS(4) T2D_POS(5) T2D_POS(6)
                mbar_wait(bar_ae, ((uint32_t)it & 1u) ^ 1u);
                T2D_PUT(4) T2D_PUT(5) T2D_PUT(6)
            }
            fence_proxy_async();
            mbar_arrive(bar_af);
#undef T2D_POS
#undef T2D_PUT
        }
    } else if (warp == MMA_WARP) {
        // The kernel is bound by shared-memory bandwidth -- the tensor core's operand reads (4 KB of A per MMA at N = 32) plus the operand
        // stores -- so the 3xTF32 products are issued as TWO MMAs per 8-deep step instead of three: A_hi [W_hi | W_lo] (N = 64, columns
        // 0..31 and 32..63 of the accumulator) and A_lo W_hi (N = 32, columns 0..31); the epilogue adds the two halves.
        constexpr uint32_t idesc32 = idesc_tf32(128, 32, 0, 0), idesc64 = idesc_tf32(128, 64, 0, 0);
        const uint64_t abase = umma_desc(0u, 128u, 512u, 0u);  // K-major, no swizzle: LBO = next 4 channels, SBO = next output position
        const uint64_t bbase = kmajor_desc(0u);
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            mbar_wait(bar_acce + 8 * abuf, ((uint32_t)(it >> 1) & 1u) ^ 1u);
            mbar_wait(bar_af, (uint32_t)it & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)(abuf * 64);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ki = tap / 3, kj = tap - 3 * ki;
                    const uint32_t a16 = (smem0 + V2_OFF_A + v2_arr_base(ki & 1, kj) + (ki >> 1) * 4 * 512) >> 4;
                    const uint32_t b16 = (smem0 + V2_OFF_B + 2 * tap * B_BLOCK) >> 4;
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        const uint64_t a_hi = abase + (a16 + 16 * k8), a_lo = a_hi + (V2_PLANE >> 4);   // two channel quads = 256 bytes per MMA
                        const uint64_t b_hi = bbase + (b16 + 2 * k8);
                        tc_mma_tf32(d, a_hi, b_hi, idesc64, (tap | k8) ? 1u : 0u);
                        tc_mma_tf32(d, a_lo, b_hi, idesc32, 1u);
                    }
                }
                tc_commit(bar_ae);                   // the operand arrays may be refilled
                tc_commit(bar_accf + 8 * abuf);      // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue (identical to version 1) =====
        const int et = threadIdx.x;
        const int img = lane & 7, opos = warp * 4 + (lane >> 3);
        float *outs = reinterpret_cast<float *>(sm + V2_OFF_OUT);
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            mbar_wait_backoff(bar_accf + 8 * abuf, (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            uint32_t r[32], r2[32];
            tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(abuf * 64), r);
            tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(abuf * 64 + 32), r2);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_acce + 8 * abuf);
#pragma unroll
            for (int oc = 0; oc < 32; ++oc)
                outs[img * OUT_IMG + oc * 16 + opos] = fmaxf(__uint_as_float(r[oc]) + __uint_as_float(r2[oc]) + __ldg(b2 + oc), 0.f);
            bar_sync(2, EPI_WARPS * 32);
            const long long n0 = tile * IMGS;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i4 = et + 128 * j, im = i4 >> 7, off = (i4 & 127) * 4;
                const float4 v = *reinterpret_cast<const float4 *>(outs + im * OUT_IMG + off);
                if (n0 + im < N) *reinterpret_cast<float4 *>(y2 + (n0 + im) * 512 + off) = v;
            }
            bar_sync(2, EPI_WARPS * 32);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, V2_TMEM_COLS);
    }
}

// =====================================================================================================================
// backward: given gy2 = dL/dy2, the gradients of all four parameter tensors.  Same tile (8 images, row m = 8 * opos + image) and the same
// shared-memory gather as the forward; two contractions per tile on the tensor cores:
//   dCOL [128 x 144] = dz2 [128 x 32] W2 [32 x 144]          A, B K-major (reduction = output channel); then col2im -> dy1 (shared memory)
//   dW2^T [144 x 32] += COL^T [144 x 128] dz2 [128 x 32]     A, B MN-major (reduction = the tile's rows), accumulated in TMEM over ALL tiles
// conv1 is recomputed (y1 is never stored in HBM); dW1 / db1 accumulate in registers, db2 in registers.  Every CTA writes its partial
// sums to a workspace and a second kernel adds them up in a fixed order: the gradients are bit-reproducible run to run.
//   warps 0..3    epilogue   tcgen05.ld of dCOL, col2im into dy1s in four tap phases (inside a phase every cell gets at most one
//                            contribution, so plain read-modify-writes are race free and the summation order is fixed)
//   warp 4        MMA
//   warps 5..11   conv1      recompute (ReLU mask kept in a register), later dz1 = dy1 * mask -> dW1, db1
//   warps 12..19  gather     dz2 = gy2 * (y2 > 0) staged transposed, its two operand images, and the COL^T stages out of y1s
constexpr int BW2_BLOCK = 144 * 64;                 // W2 as the dCOL B operand: 144 rows (tap, ic) x 16 output channels
constexpr int BW2_BYTES = 4 * BW2_BLOCK;            // [hi | lo][2 reduction blocks]
constexpr int BSTAGES = 2;
constexpr int BA_SBO = 5 * 512;                     // COL^T: 144 rows = 4.5 atoms of 32 rows per group of 4 reduction indices
constexpr int BA_PLANE = 4 * BA_SBO;                // 16 reduction indices
constexpr int BB_SBO = 1024;                        // dz2 as MN-major B: per group of 4 reduction indices the hi atom (32 rows) and the lo atom side by
constexpr int BB_PLANE = 4 * 512;                   // side = ONE 64-row operand [dz_hi | dz_lo] (see the dW2 MMAs); BB_PLANE = bytes of one of hi / lo
constexpr int BSTAGE_BYTES = 2 * BA_PLANE + 2 * BB_PLANE;  // 24,576
constexpr int DZA_BYTES = 4 * A_TILE;               // dz2 as the dCOL A operand: [hi | lo] x 2 reduction blocks of 128 x 16
constexpr int DZT_PITCH = 36;                       // floats per row of the staged dz2 [m][oc]
constexpr int DZT_BYTES = 128 * DZT_PITCH * 4;
// dy1s [img][pos][ic], scalar accesses (a conflict-free scalar store moves 128 bytes per LSU cycle, a 128-bit one half of that -- the
// float4 layout was measured slower).  Bank of (img, pos, ic) = img offset + 18 pos + ic with the image offsets = {0, 1, 2, 3, 16, 17, 18,
// 19} mod 32: the col2im lanes (8 images x 4 positions two apart: + {0, 4, 8, 12}) and the dW1 lanes (warp = row; 8 images x 4 channel
// quads: + {0, 4, 8, 12}) both cover all 32 banks exactly once.
constexpr int DY_POS = 18, DY_IMG4 = 3600, DY_IMG1 = 897;
__host__ __device__ constexpr int dy_img(int img) { return (img >> 2) * DY_IMG4 + (img & 3) * DY_IMG1; }
static_assert(DY_IMG1 % 32 == 1 && DY_IMG4 % 32 == 16 && DY_IMG1 >= 49 * DY_POS && DY_IMG4 >= 4 * DY_IMG1, "dy1s banks");
constexpr int DY_FLOATS = 2 * DY_IMG4;
constexpr int DY_BYTES = DY_FLOATS * 4;
constexpr int BOFF_W2 = 0;
constexpr int BOFF_RING = BOFF_W2 + BW2_BYTES;                 // 36,864
constexpr int BOFF_DZA = BOFF_RING + BSTAGES * BSTAGE_BYTES;   // 110,592
constexpr int BOFF_DZT = BOFF_DZA + DZA_BYTES;
constexpr int BOFF_Y1 = BOFF_DZT + DZT_BYTES;
constexpr int BOFF_DY = BOFF_Y1 + 2 * Y1_BYTES;                // y1s and the pixel tile are double-buffered: conv1 runs one tile ahead
constexpr int BOFF_XS = BOFF_DY + (DY_BYTES + 15) / 16 * 16;
constexpr int BOFF_BAR = BOFF_XS + 2 * XS_BYTES;
constexpr int BWD_SMEM = BOFF_BAR + 256 + 640 + 1024;  // barriers, conv1 weights + biases, alignment slack
constexpr int BWD_THREADS = FWD_THREADS;
constexpr int BTMEM_COLS = 512, TM_DW2 = 0, TM_DCOL = 128, TM_DCOL_STRIDE = 160;  // dW2: 2 row blocks x [32 + 32] columns
constexpr int PART_FLOATS = 32 * 144 + 144 + 16 + 32;  // per-CTA partial sums: dw2, dw1, db1, db2
static_assert(BOFF_RING % 1024 == 0 && BOFF_DZA % 1024 == 0 && BSTAGE_BYTES % 512 == 0 && BOFF_DZT % 16 == 0 && BOFF_Y1 % 16 == 0, "shared-memory layout");
static_assert(BWD_SMEM <= 227 * 1024, "shared memory");

template <typename XT>
__global__ void __launch_bounds__(BWD_THREADS, 1) conv_tc_bwd_kernel(const XT *__restrict__ x, long long xs, const float *__restrict__ y2,
                                                                    const float *__restrict__ gy2, long long N, const float *__restrict__ w1,
                                                                    const float *__restrict__ b1, const float *__restrict__ w2, float *__restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t smem0 = (raw0 + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (smem0 - raw0);
    const uint32_t bars = smem0 + BOFF_BAR;
    const uint32_t bar_y1f = bars, bar_y1e = bars + 16, bar_dyf = bars + 32, bar_dye = bars + 40, bar_dzf = bars + 48, bar_dze = bars + 56,
                   bar_full = bars + 64, bar_empty = bar_full + 8 * BSTAGES, bar_accf = bar_empty + 8 * BSTAGES, bar_acce = bar_accf + 16,
                   bar_dw2 = bar_acce + 16, tmem_slot = bar_dw2 + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (N + IMGS - 1) / IMGS;
    float *my_part = part + (long long)blockIdx.x * PART_FLOATS;

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_y1f + 8 * b, C1_WARPS * 32);
            mbar_init(bar_y1e + 8 * b, G_WARPS * 32);
        }
        mbar_init(bar_dyf, EPI_WARPS * 32);
        mbar_init(bar_dye, C1_WARPS * 32);
        mbar_init(bar_dzf, G_WARPS * 32);
        mbar_init(bar_dze, 1);
        for (int s = 0; s < BSTAGES; ++s) {
            mbar_init(bar_full + 8 * s, G_WARPS * 32);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_accf + 8 * b, 1);
            mbar_init(bar_acce + 8 * b, EPI_WARPS * 32);
        }
        mbar_init(bar_dw2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, BTMEM_COLS);
    // W2 as the B operand of dCOL = dz2 W2: rows n = (tap, ic), reduction = output channel in two blocks of 16
    for (int i = threadIdx.x; i < 32 * 144; i += BWD_THREADS) {
        const int oc = i / 144, rem = i - oc * 144, ic = rem / 9, tap = rem - ic * 9;
        float hi, lo;
        split1(w2[i], hi, lo);
        const uint32_t off = (uint32_t)((oc >> 4) * BW2_BLOCK) + kmajor_off(tap * 16 + ic, (oc & 15) >> 2) + (uint32_t)(oc & 3) * 4u;
        *reinterpret_cast<float *>(sm + BOFF_W2 + off) = hi;
        *reinterpret_cast<float *>(sm + BOFF_W2 + 2 * BW2_BLOCK + off) = lo;
    }
    for (int i = threadIdx.x; i < 160; i += BWD_THREADS) {  // conv1 weights as [q][k][4 channels], then biases [q][4]
        float v;
        if (i < 144) { const int qq = i / 36, k = (i % 36) / 4, c = i & 3; v = w1[(4 * qq + c) * 9 + k]; }
        else v = b1[i - 144];
        reinterpret_cast<float *>(sm + BOFF_BAR + 256)[i] = v;
    }
    for (int i = threadIdx.x; i < 2 * XS_BYTES / 4; i += BWD_THREADS) reinterpret_cast<uint32_t *>(sm + BOFF_XS)[i] = 0u;
    for (int i = threadIdx.x; i < DY_FLOATS; i += BWD_THREADS) reinterpret_cast<float *>(sm + BOFF_DY)[i] = 0.f;
    for (int i = threadIdx.x; i < BSTAGES * BSTAGE_BYTES / 4; i += BWD_THREADS) reinterpret_cast<float *>(sm + BOFF_RING)[i] = 0.f;  // rows 144..159 stay 0
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + BOFF_BAR + (tmem_slot - bars));

    if (warp >= G_WARP0) {
        // ===== gather =====
        const int gt = threadIdx.x - G_WARP0 * 32;
        const int l_oc = gt & 31, l_op4 = (gt >> 5) & 3, l_img0 = gt >> 7;  // loader: float4 f = gt + 256 j -> image l_img0 + 2 j
        float db2_acc = 0.f;
        float4 pg[4], py[4];
        auto prefetch = [&](long long tile) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long n = tile * IMGS + l_img0 + 2 * j;
                if (n < N) {
                    const long long o = n * 512 + l_oc * 16 + l_op4 * 4;
                    pg[j] = __ldg(reinterpret_cast<const float4 *>(gy2 + o));
                    py[j] = __ldg(reinterpret_cast<const float4 *>(y2 + o));
                } else {
                    pg[j] = py[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        prefetch(blockIdx.x);
        float *dzt = reinterpret_cast<float *>(sm + BOFF_DZT);
        int ist = 0, it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            // (a) dz2 = gy2 * (y2 > 0), staged as [m = 8 opos + img][oc]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int img = l_img0 + 2 * j;
                const float d0 = py[j].x > 0.f ? pg[j].x : 0.f, d1 = py[j].y > 0.f ? pg[j].y : 0.f, d2 = py[j].z > 0.f ? pg[j].z : 0.f,
                            d3 = py[j].w > 0.f ? pg[j].w : 0.f;
                db2_acc += (d0 + d1) + (d2 + d3);
                float *o = dzt + ((l_op4 * 4) * 8 + img) * DZT_PITCH + l_oc;
                o[0] = d0; o[8 * DZT_PITCH] = d1; o[16 * DZT_PITCH] = d2; o[24 * DZT_PITCH] = d3;
            }
            const long long next = tile + gridDim.x;
            if (next < n_tiles) prefetch(next);
            bar_sync(3, G_WARPS * 32);
            // (b) dz2 as the A operand of dCOL (K-major, two reduction blocks of 16 output channels)
            mbar_wait(bar_dze, ((uint32_t)it & 1u) ^ 1u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = gt + 256 * j, img = q & 7, c8 = (q >> 3) & 7, opos = q >> 6, m = opos * 8 + img;
                float4 hi, lo;
                split4(lds128(smem0 + BOFF_DZT + (uint32_t)(m * DZT_PITCH + c8 * 4) * 4u), hi, lo);
                const uint32_t dst = smem0 + BOFF_DZA + (uint32_t)((c8 >> 2) * A_TILE) + kmajor_off(m, c8 & 3);
                sts128(dst, hi);
                sts128(dst + 2 * A_TILE, lo);
            }
            fence_proxy_async();
            mbar_arrive(bar_dzf);
            // (c) the COL^T stages: 16 rows of the tile (two output positions x 8 images) per stage
            const int ybuf = it & 1;
            mbar_wait(bar_y1f + 8 * ybuf, (uint32_t)(it >> 1) & 1u);
#pragma unroll 1
            for (int s8 = 0; s8 < 8; ++s8, ++ist) {
                const int s = ist % BSTAGES;
                mbar_wait(bar_empty + 8 * s, ((uint32_t)(ist / BSTAGES) & 1u) ^ 1u);
                const uint32_t st = smem0 + BOFF_RING + s * BSTAGE_BYTES;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int q = gt + 256 * j;
                    if (q < 576) {
                        const int b0 = q & 1, ih = (q >> 1) & 3, par = (q >> 3) & 1, ol = (q >> 4) & 1, cp = q >> 5;
                        const int img = ih * 2 + par, kk = ih | (par << 2) | (ol << 3), c36 = 2 * cp + b0, tap = c36 >> 2, c = c36 & 3;
                        const int opos = 2 * s8 + ol, ki = tap / 3, kj = tap - ki * 3;
                        const int r = 2 * (opos >> 2) + ki - 1, cc = 2 * (opos & 3) + kj - 1;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
                        if (r >= 0 && r < 7 && cc >= 0 && cc < 7) v = lds128(smem0 + BOFF_Y1 + ybuf * Y1_BYTES + (uint32_t)(img * Y1_IMG + (r * 7 + cc) * 16 + c * 4) * 4u);
                        split4(v, hi, lo);
                        const uint32_t off = mnmajor_off(kk, c36, BA_SBO);
                        sts128(st + off, hi);
                        sts128(st + BA_PLANE + off, lo);
                    }
                }
                if (gt < 128) {  // dz2 as the MN-major B operand of this stage
                    const int b0 = gt & 1, ih = (gt >> 1) & 3, par = (gt >> 3) & 1, ol = (gt >> 4) & 1, cp = gt >> 5;
                    const int img = ih * 2 + par, kk = ih | (par << 2) | (ol << 3), c8 = 2 * cp + b0, m = (2 * s8 + ol) * 8 + img;
                    float4 hi, lo;
                    split4(lds128(smem0 + BOFF_DZT + (uint32_t)(m * DZT_PITCH + c8 * 4) * 4u), hi, lo);
                    const uint32_t off = 2 * BA_PLANE + mnmajor_off(kk, c8, BB_SBO);
                    sts128(st + off, hi);
                    sts128(st + off + 512, lo);
                }
                fence_proxy_async();
                mbar_arrive(bar_full + 8 * s);
            }
            mbar_arrive(bar_y1e + 8 * ybuf);
            bar_sync(3, G_WARPS * 32);  // dzt is rewritten by the next tile
        }
        // db2: the 8 loader threads of an output channel, added in a fixed order (the ring is scratch once the last MMA has read it)
        __syncthreads();
        mbar_wait_backoff(bar_dw2, 0u);
        float *red = reinterpret_cast<float *>(sm + BOFF_RING);
        red[gt] = db2_acc;
        bar_sync(3, G_WARPS * 32);
        if (gt < 32) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k * 32 + gt];
            my_part[32 * 144 + 144 + 16 + gt] = t;
        }
    } else if (warp >= C1_WARP0) {
        // ===== conv1 recompute (one tile ahead), then dz1 -> dW1 / db1 =====
        const int ct = threadIdx.x - C1_WARP0 * 32;
        const int row = ct >> 5, img = lane & 7, q = lane >> 3;  // warp = output row, lane = (image, channel quad): see the bank layout of dy1s
        // the conv1 weights are re-read from shared memory per tile: the registers hold the 40 gradient accumulators
        float aw1[4][9], ab1[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < 9; ++k) aw1[c][k] = 0.f;
            ab1[c] = 0.f;
        }
        const float4 *wsm = reinterpret_cast<const float4 *>(sm + BOFF_BAR + 256) + 9 * q;
        const float4 *bsm = reinterpret_cast<const float4 *>(sm + BOFF_BAR + 256 + 576) + q;
        constexpr int PRE = (IMGS * 169 + C1_WARPS * 32 - 1) / (C1_WARPS * 32);
        uint32_t pre[PRE];
#pragma unroll
        for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, (long long)blockIdx.x * IMGS, ct + C1_WARPS * 32 * j);
        // stage the prefetched pixels of tile number `jt` (and prefetch the following tile), then conv1 + ReLU into y1s[jt & 1]
        auto conv1_tile = [&](int jt, long long following) -> uint32_t {
            const int buf = jt & 1;
            uint8_t *xs8 = sm + BOFF_XS + buf * XS_BYTES;
#pragma unroll
            for (int j = 0; j < PRE; ++j) store_px(xs8, ct + C1_WARPS * 32 * j, pre[j]);
            if (following < n_tiles) {
#pragma unroll
                for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, following * IMGS, ct + C1_WARPS * 32 * j);
            }
            bar_sync(1, C1_WARPS * 32);
            mbar_wait(bar_y1e + 8 * buf, ((uint32_t)(jt >> 1) & 1u) ^ 1u);
            uint32_t mask = 0u;  // bit 4 j + c: y1 > 0 at column j, channel 4 q + c
            if (row < 7) {
                uint4 win[3];
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) win[ki] = *reinterpret_cast<const uint4 *>(xs8 + img * XS_IMG + (2 * row + ki) * 16);
                const float4 *w = wsm;  // read per tap from shared memory (broadcast loads): the registers are taken by the accumulators
                const float4 bias = *bsm;
                float *yo = reinterpret_cast<float *>(sm + BOFF_Y1 + buf * Y1_BYTES) + img * Y1_IMG + (row * 7) * 16 + 4 * q;
#define T2D_POS(J)                                                                                          \
    {                                                                                                       \
        const float4 a = conv1_at<J>(win, w, bias);                                                         \
        mask |= ((a.x > 0.f ? 1u : 0u) | (a.y > 0.f ? 2u : 0u) | (a.z > 0.f ? 4u : 0u) | (a.w > 0.f ? 8u : 0u)) << (4 * J); \
        *reinterpret_cast<float4 *>(yo + J * 16) = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); \
    }
                T2D_POS(0) T2D_POS(1) T2D_POS(2) T2D_POS(3) T2D_POS(4) T2D_POS(5) T2D_POS(6)
#undef T2D_POS
            }
            mbar_arrive(bar_y1f + 8 * buf);
            return mask;
        };
        uint32_t mask_cur = conv1_tile(0, (long long)blockIdx.x + gridDim.x), mask_next = 0u;
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long next = tile + gridDim.x;
            if (next < n_tiles) {
                bar_sync(1, C1_WARPS * 32);  // every thread is done with the pixel buffer conv1_tile is about to refill (dW1 of tile it - 1)
                mask_next = conv1_tile(it + 1, next + gridDim.x);
            }
            // dy1 of this tile (col2im done by the epilogue warps) -> dz1 = dy1 * (y1 > 0) -> dW1, db1; the cells are zeroed for the next tile
            mbar_wait_backoff(bar_dyf, (uint32_t)it & 1u);
            if (row < 7) {
                const uint8_t *xs8 = sm + BOFF_XS + (it & 1) * XS_BYTES;
                uint4 win[3];
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) win[ki] = *reinterpret_cast<const uint4 *>(xs8 + img * XS_IMG + (2 * row + ki) * 16);
                float *dy = reinterpret_cast<float *>(sm + BOFF_DY) + dy_img(img) + (row * 7) * DY_POS + 4 * q;
#define T2D_TAP(J, KI, KJ)                                                                     \
    {                                                                                          \
        const float v = win_px<2 * J + KJ>(win[KI]);                                           \
        aw1[0][KI * 3 + KJ] = fmaf(dz[0], v, aw1[0][KI * 3 + KJ]); aw1[1][KI * 3 + KJ] = fmaf(dz[1], v, aw1[1][KI * 3 + KJ]); \
        aw1[2][KI * 3 + KJ] = fmaf(dz[2], v, aw1[2][KI * 3 + KJ]); aw1[3][KI * 3 + KJ] = fmaf(dz[3], v, aw1[3][KI * 3 + KJ]); \
    }
#define T2D_POS(J)                                                                             \
    {                                                                                          \
        float dz[4];                                                                           \
        _Pragma("unroll") for (int c = 0; c < 4; ++c) {                                        \
            const float d = dy[J * DY_POS + c];                                                \
            dy[J * DY_POS + c] = 0.f;                                                          \
            dz[c] = ((mask_cur >> (4 * J + c)) & 1u) ? d : 0.f;                                \
            ab1[c] += dz[c];                                                                   \
        }                                                                                      \
        T2D_TAP(J, 0, 0) T2D_TAP(J, 0, 1) T2D_TAP(J, 0, 2) T2D_TAP(J, 1, 0) T2D_TAP(J, 1, 1) T2D_TAP(J, 1, 2) \
        T2D_TAP(J, 2, 0) T2D_TAP(J, 2, 1) T2D_TAP(J, 2, 2)                                     \
    }
                T2D_POS(0) T2D_POS(1) T2D_POS(2) T2D_POS(3) T2D_POS(4) T2D_POS(5) T2D_POS(6)
#undef T2D_POS
#undef T2D_TAP
            }
            mbar_arrive(bar_dye);
            mask_cur = mask_next;
        }
        // dW1 / db1: per-thread sums -> shared memory -> added in a fixed order (56 (image, row) slots per channel entry)
        __syncthreads();
        mbar_wait_backoff(bar_dw2, 0u);
        float *red = reinterpret_cast<float *>(sm + BOFF_RING) + 256;  // [ct][40]
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < 9; ++k) red[ct * 40 + c * 10 + k] = aw1[c][k];
            red[ct * 40 + c * 10 + 9] = ab1[c];
        }
        bar_sync(1, C1_WARPS * 32);
        if (ct < 160) {  // entry (channel ch, k): k < 9 -> dw1[ch][k], k == 9 -> db1[ch]
            const int ch = ct / 10, k = ct - ch * 10, qq = ch >> 2, c = ch & 3;
            float t = 0.f;
            for (int im = 0; im < 8; ++im)
                for (int rw = 0; rw < 7; ++rw) t += red[(rw * 32 + qq * 8 + im) * 40 + c * 10 + k];
            if (k < 9) my_part[32 * 144 + ch * 9 + k] = t;
            else my_part[32 * 144 + 144 + ch] = t;
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc_col = idesc_tf32(128, 144, 0, 0), idesc_w = idesc_tf32(128, 32, 1, 1), idesc_w64 = idesc_tf32(128, 64, 1, 1);
        const uint64_t kbase = kmajor_desc(0u), abase = mnmajor_desc(0u, BA_SBO), bbase = mnmajor_desc(0u, BB_SBO);
        int ist = 0, it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            mbar_wait(bar_dzf, (uint32_t)it & 1u);
            mbar_wait(bar_acce + 8 * abuf, ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)(TM_DCOL + abuf * TM_DCOL_STRIDE);
                const uint32_t a16 = (smem0 + BOFF_DZA) >> 4, b16 = (smem0 + BOFF_W2) >> 4;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        const uint64_t a_hi = kbase + (a16 + ((kb * A_TILE) >> 4) + 2 * k8), a_lo = a_hi + ((2 * A_TILE) >> 4);
                        const uint64_t b_hi = kbase + (b16 + ((kb * BW2_BLOCK) >> 4) + 2 * k8), b_lo = b_hi + ((2 * BW2_BLOCK) >> 4);
                        tc_mma_tf32(d, a_lo, b_hi, idesc_col, (kb > 0 || k8 > 0) ? 1u : 0u);
                        tc_mma_tf32(d, a_hi, b_lo, idesc_col, 1u);
                        tc_mma_tf32(d, a_hi, b_hi, idesc_col, 1u);
                    }
                tc_commit(bar_dze);                // the dz2 operand may be rebuilt
                tc_commit(bar_accf + 8 * abuf);    // dCOL complete
            }
            __syncwarp();
#pragma unroll 1
            for (int s8 = 0; s8 < 8; ++s8, ++ist) {
                const int s = ist % BSTAGES;
                mbar_wait(bar_full + 8 * s, (uint32_t)(ist / BSTAGES) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st16 = (smem0 + BOFF_RING + s * BSTAGE_BYTES) >> 4;
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        // 3xTF32 in two MMAs: COL_hi^T [dz_hi | dz_lo] (N = 64: accumulator columns 0..31 and 32..63) and COL_lo^T dz_hi --
                        // one operand read of COL_hi instead of two (the kernel lives on shared-memory bandwidth)
                        const uint64_t b_hi = bbase + (st16 + ((2 * BA_PLANE + k8 * 2 * BB_SBO) >> 4));
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const uint64_t a_hi = abase + (st16 + ((k8 * 2 * BA_SBO + half * 2048) >> 4)), a_lo = a_hi + (BA_PLANE >> 4);
                            const uint32_t d = tmem_base + (uint32_t)(TM_DW2 + half * 64);
                            tc_mma_tf32(d, a_hi, b_hi, idesc_w64, (ist > 0 || k8 > 0) ? 1u : 0u);
                            tc_mma_tf32(d, a_lo, b_hi, idesc_w, 1u);
                        }
                    }
                    tc_commit(bar_empty + 8 * s);
                }
                __syncwarp();
            }
        }
        if (elect_one()) tc_commit(bar_dw2);
        __syncwarp();
        __syncthreads();
    } else {
        // ===== epilogue: col2im of dCOL, finally the dW2 read-out =====
        const int img = lane & 7, opos = warp * 4 + (lane >> 3), oi = opos >> 2, oj = opos & 3;
        float *dy = reinterpret_cast<float *>(sm + BOFF_DY) + dy_img(img);
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            mbar_wait_backoff(bar_accf + 8 * abuf, (uint32_t)(it >> 1) & 1u);
            mbar_wait(bar_dye, ((uint32_t)it & 1u) ^ 1u);  // conv1 warps are done with (and have zeroed) the previous tile's dy1
            tc_fence_after();
            const uint32_t tcol = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(TM_DCOL + abuf * TM_DCOL_STRIDE);
            // tap phases: inside a phase no two (position, tap) pairs reach the same dy1 cell
            const int order[9] = {0, 1, 3, 4, /**/ 6, 7, /**/ 2, 5, /**/ 8};
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
                const int tap = order[t9], ki = tap / 3, kj = tap - ki * 3;
                if (t9 == 4 || t9 == 6 || t9 == 8) bar_sync(2, EPI_WARPS * 32);
                uint32_t r[16];
                tc_ld16(tcol + (uint32_t)(tap * 16), r);
                tc_wait_ld();
                const int rr = 2 * oi + ki - 1, cc = 2 * oj + kj - 1;
                if (rr >= 0 && rr < 7 && cc >= 0 && cc < 7) {
                    float *cell = dy + (rr * 7 + cc) * DY_POS;
#pragma unroll
                    for (int ic = 0; ic < 16; ++ic) cell[ic] += __uint_as_float(r[ic]);
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acce + 8 * abuf);
            mbar_arrive(bar_dyf);
        }
        __syncthreads();
        // dW2^T [k_idx = (tap, ic)][oc]: rows 0..127 in columns 0..31 (+ the COL_hi dz_lo part in 32..63), rows 128..143 in lanes 0..15 of
        // columns 64..95 (+ 96..127)
        mbar_wait_backoff(bar_dw2, 0u);
        tc_fence_after();
        uint32_t r[32], r2[32];
        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)TM_DW2, r);
        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(TM_DW2 + 32), r2);
        tc_wait_ld();
        {
            const int kidx = warp * 32 + lane, tap = kidx >> 4, ic = kidx & 15;
#pragma unroll
            for (int oc = 0; oc < 32; ++oc) my_part[oc * 144 + ic * 9 + tap] = __uint_as_float(r[oc]) + __uint_as_float(r2[oc]);
        }
        if (warp == 0) {
            tc_ld32(tmem_base + (uint32_t)(TM_DW2 + 64), r);
            tc_ld32(tmem_base + (uint32_t)(TM_DW2 + 96), r2);
            tc_wait_ld();
            if (lane < 16) {
#pragma unroll
                for (int oc = 0; oc < 32; ++oc) my_part[oc * 144 + lane * 9 + 8] = __uint_as_float(r[oc]) + __uint_as_float(r2[oc]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BTMEM_COLS);
    }
}

// dst[i] += sum over CTAs of part[cta][i], in CTA order
__global__ void __launch_bounds__(256) conv_bwd_reduce_kernel(const float *__restrict__ part, int n_part, float *__restrict__ dw2, float *__restrict__ dw1,
                                                              float *__restrict__ db1, float *__restrict__ db2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PART_FLOATS) return;
    float t = 0.f;
    for (int p = 0; p < n_part; ++p) t += part[(long long)p * PART_FLOATS + i];
    if (i < 32 * 144) dw2[i] += t;
    else if (i < 32 * 144 + 144) dw1[i - 32 * 144] += t;
    else if (i < 32 * 144 + 144 + 16) db1[i - 32 * 144 - 144] += t;
    else db2[i - 32 * 144 - 144 - 16] += t;
}

int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

template <typename XT>
cudaError_t fwd2_launch(const XT *x, int64_t xs, int64_t n, const float *w1, const float *b1, const float *w2, const float *b2, float *y2, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_fwd2_kernel<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int64_t tiles = (n + IMGS - 1) / IMGS;
    const int sms = sm_count();
    conv_tc_fwd2_kernel<XT><<<(int)(tiles < sms ? tiles : sms), V2_THREADS, V2_SMEM, st>>>(x, xs, n, w1, b1, w2, b2, y2);
    return cudaGetLastError();
}

template <typename XT>
cudaError_t fwd_launch(const XT *x, int64_t xs, int64_t n, const float *w1, const float *b1, const float *w2, const float *b2, float *y2, cudaStream_t st) {
    static int use_v1 = -1;  // T2D_CONV_FWD=v1: the gather version (A/B measurements)
    if (use_v1 < 0) { const char *e = getenv("T2D_CONV_FWD"); use_v1 = (e && strcmp(e, "v1") == 0) ? 1 : 0; }
    if (!use_v1) return fwd2_launch(x, xs, n, w1, b1, w2, b2, y2, st);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_fwd_kernel<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int64_t tiles = (n + IMGS - 1) / IMGS;
    const int sms = sm_count();
    conv_tc_fwd_kernel<XT><<<(int)(tiles < sms ? tiles : sms), FWD_THREADS, FWD_SMEM, st>>>(x, xs, n, w1, b1, w2, b2, y2);
    return cudaGetLastError();
}

float *g_part[64] = {};  // per-device workspace of per-CTA partial sums (one stream at a time, like the env handle)

template <typename XT>
cudaError_t bwd_launch(const XT *x, int64_t xs, const float *y2, const float *gy2, int64_t n, const float *w1, const float *b1, const float *w2, float *dw1,
                       float *db1, float *dw2, float *db2, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int sms = sm_count();
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_bwd_kernel<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    if (!g_part[dev & 63]) {
        cudaError_t e = cudaMalloc(&g_part[dev & 63], (size_t)sms * PART_FLOATS * sizeof(float));
        if (e != cudaSuccess) return e;
    }
    const int64_t tiles = (n + IMGS - 1) / IMGS;
    const int grid = (int)(tiles < sms ? tiles : sms);
    conv_tc_bwd_kernel<XT><<<grid, BWD_THREADS, BWD_SMEM, st>>>(x, xs, y2, gy2, n, w1, b1, w2, g_part[dev & 63]);
    conv_bwd_reduce_kernel<<<(PART_FLOATS + 255) / 256, 256, 0, st>>>(g_part[dev & 63], grid, dw2, dw1, db1, db2);
    return cudaGetLastError();
}

}  // namespace

cudaError_t t2d_conv_tc_backward(const void *x, int x_is_u8, int64_t xs, const float *y2, const float *gy2, int64_t n, const float *w1, const float *b1,
                                 const float *w2, float *dw1, float *db1, float *dw2, float *db2, cudaStream_t st) {
    return x_is_u8 ? bwd_launch((const uint8_t *)x, xs, y2, gy2, n, w1, b1, w2, dw1, db1, dw2, db2, st)
                   : bwd_launch((const float *)x, xs, y2, gy2, n, w1, b1, w2, dw1, db1, dw2, db2, st);
}

// entry points used by track2d_policy.cu's dispatcher
cudaError_t t2d_conv_tc_forward(const void *x, int x_is_u8, int64_t xs, int64_t n, const float *w1, const float *b1, const float *w2, const float *b2, float *y2,
                                cudaStream_t st) {
    return x_is_u8 ? fwd_launch((const uint8_t *)x, xs, n, w1, b1, w2, b2, y2, st) : fwd_launch((const float *)x, xs, n, w1, b1, w2, b2, y2, st);
}
