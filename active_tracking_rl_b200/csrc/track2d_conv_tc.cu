// track2d_conv_tc.cu -- CNN_maze's convolution stack (perception.py:68-92) with conv2 on the 5th-generation tensor cores.
//
//   x [N][13][13] (uint8 or float32) -> conv1 3x3/s2/p1 (1->16) + ReLU -> conv2 3x3/s2/p1 (16->32) + ReLU -> y2 [N][32*4*4]
//
// conv2 is 73.7 k of the stack's 80.8 k MAC per image and is an implicit GEMM: rows = (image, output position), reduction =
// (tap, input channel) = 144, columns = 32 output channels.  A CTA (one per SM, persistent) works on tiles of 8 images = 128 rows:
//
//   warps 5..12   conv1    warp = image, lane = (output row, channel quad): 7 x 4 outputs x 9 FMA on the CUDA cores from a zero-bordered
//                          image tile (weights in registers), written channel-last into shared memory (y1s, double-buffered) as 16-byte
//                          stores; the next tile's pixels are prefetched into registers
//   warps 13..20  gather   per tap (ki, kj): the 128 x 16 slice of the im2col matrix = one 16-byte copy per (row, 4 channels) out of y1s
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
constexpr int XS_IMG = 15 * 15;                // zero-bordered input image
constexpr int XS_BYTES = IMGS * XS_IMG * 4;
constexpr int OUT_IMG = 512 + 4;               // staged output row of one image
constexpr int OUT_BYTES = IMGS * OUT_IMG * 4;
constexpr int OFF_B = 0;
constexpr int OFF_A = OFF_B + B_BYTES;                 // 36,864 (1024-aligned)
constexpr int OFF_Y1 = OFF_A + STAGES * STAGE_BYTES;   // 102,400
constexpr int OFF_XS = OFF_Y1 + 2 * Y1_BYTES;
constexpr int OFF_OUT = OFF_XS + 2 * XS_BYTES;
constexpr int OFF_BAR = OFF_OUT + OUT_BYTES;
constexpr int FWD_SMEM = OFF_BAR + 256 + 1024 /* alignment slack */;
constexpr int EPI_WARPS = 4, MMA_WARP = 4, C1_WARP0 = 5, C1_WARPS = 8, G_WARP0 = 13, G_WARPS = 8;
constexpr int FWD_THREADS = (G_WARP0 + G_WARPS) * 32;  // 672
constexpr int TMEM_COLS = 64;
static_assert(OFF_A % 1024 == 0 && OFF_Y1 % 16 == 0 && OFF_XS % 16 == 0 && OFF_OUT % 16 == 0 && OFF_BAR % 8 == 0, "shared-memory layout");
static_assert(FWD_SMEM <= 227 * 1024, "shared memory");

template <typename XT>
__device__ __forceinline__ float load_px(const XT *__restrict__ x, long long xs, long long N, long long n0, int e) {
    const int img = e / 169, c = e - img * 169;
    return (e < IMGS * 169 && n0 + img < N) ? (float)x[(n0 + img) * xs + c] : 0.f;
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
    for (int i = threadIdx.x; i < 2 * IMGS * XS_IMG; i += FWD_THREADS) reinterpret_cast<float *>(sm + OFF_XS)[i] = 0.f;  // borders stay zero
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
        // warp = image; lane = (output row i, channel quad q): 7 positions x 4 channels per thread, weights in registers for the
        // whole kernel, every pixel read from shared memory serves 4 FMAs, results leave as 16-byte channel-last stores
        const int ct = threadIdx.x - C1_WARP0 * 32;
        const int img = ct >> 5, row = (ct >> 2) & 7, q = ct & 3;
        float w[4][9], bias[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < 9; ++k) w[c][k] = __ldg(w1 + (4 * q + c) * 9 + k);
            bias[c] = __ldg(b1 + 4 * q + c);
        }
        constexpr int PRE = (IMGS * 169 + C1_WARPS * 32 - 1) / (C1_WARPS * 32);  // 6
        float pre[PRE];
#pragma unroll
        for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, (long long)blockIdx.x * IMGS, ct + C1_WARPS * 32 * j);
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            float *xsb = reinterpret_cast<float *>(sm + OFF_XS + buf * XS_BYTES);
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                const int e = ct + C1_WARPS * 32 * j;
                if (e < IMGS * 169) {
                    const int im = e / 169, c = e - im * 169, r = c / 13, qq = c - r * 13;
                    xsb[im * XS_IMG + (r + 1) * 15 + qq + 1] = pre[j];
                }
            }
            const long long next = tile + gridDim.x;
            if (next < n_tiles) {
#pragma unroll
                for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, next * IMGS, ct + C1_WARPS * 32 * j);
            }
            bar_sync(1, C1_WARPS * 32);
            mbar_wait(bar_y1e + 8 * buf, ((uint32_t)(it >> 1) & 1u) ^ 1u);
            if (row < 7) {
                const float *xi = xsb + img * XS_IMG + (2 * row) * 15;
                float *yo = reinterpret_cast<float *>(sm + OFF_Y1 + buf * Y1_BYTES) + img * Y1_IMG + (row * 7) * 16 + 4 * q;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    float a0 = bias[0], a1 = bias[1], a2 = bias[2], a3 = bias[3];
#pragma unroll
                    for (int ki = 0; ki < 3; ++ki)
#pragma unroll
                        for (int kj = 0; kj < 3; ++kj) {
                            const float v = xi[ki * 15 + 2 * j + kj];
                            a0 = fmaf(v, w[0][ki * 3 + kj], a0);
                            a1 = fmaf(v, w[1][ki * 3 + kj], a1);
                            a2 = fmaf(v, w[2][ki * 3 + kj], a2);
                            a3 = fmaf(v, w[3][ki * 3 + kj], a3);
                        }
                    *reinterpret_cast<float4 *>(yo + j * 16) = make_float4(fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f));
                }
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
// backward: given gy2 = dL/dy2, the gradients of all four parameter tensors.  Same tile (8 images, row m = 8 * opos + image) and the same
// shared-memory gather as the forward; two contractions per tile on the tensor cores:
//   dCOL [128 x 144] = dz2 [128 x 32] W2 [32 x 144]          A, B K-major (reduction = output channel); then col2im -> dy1 (shared memory)
//   dW2^T [144 x 32] += COL^T [144 x 128] dz2 [128 x 32]     A, B MN-major (reduction = the tile's rows), accumulated in TMEM over ALL tiles
// conv1 is recomputed (y1 is never stored in HBM); dW1 / db1 accumulate in registers, db2 in registers.  Every CTA writes its partial
// sums to a workspace and a second kernel adds them up in a fixed order: the gradients are bit-reproducible run to run.
//   warps 0..3    epilogue   tcgen05.ld of dCOL, col2im into dy1s in four tap phases (inside a phase every cell gets at most one
//                            contribution, so plain read-modify-writes are race free and the summation order is fixed)
//   warp 4        MMA
//   warps 5..12   conv1      recompute (ReLU mask kept in a register), later dz1 = dy1 * mask -> dW1, db1
//   warps 13..20  gather     dz2 = gy2 * (y2 > 0) staged transposed, its two operand images, and the COL^T stages out of y1s
constexpr int BW2_BLOCK = 144 * 64;                 // W2 as the dCOL B operand: 144 rows (tap, ic) x 16 output channels
constexpr int BW2_BYTES = 4 * BW2_BLOCK;            // [hi | lo][2 reduction blocks]
constexpr int BSTAGES = 3;
constexpr int BA_SBO = 5 * 512;                     // COL^T: 144 rows = 4.5 atoms of 32 rows per group of 4 reduction indices
constexpr int BA_PLANE = 4 * BA_SBO;                // 16 reduction indices
constexpr int BB_SBO = 512;                         // dz2 as MN-major B: 32 rows = one atom per group
constexpr int BB_PLANE = 4 * BB_SBO;
constexpr int BSTAGE_BYTES = 2 * BA_PLANE + 2 * BB_PLANE;  // 24,576
constexpr int DZA_BYTES = 4 * A_TILE;               // dz2 as the dCOL A operand: [hi | lo] x 2 reduction blocks of 128 x 16
constexpr int DZT_PITCH = 36;                       // floats per row of the staged dz2 [m][oc]
constexpr int DZT_BYTES = 128 * DZT_PITCH * 4;
constexpr int DY_POS = 20, DY_IMG = 49 * DY_POS + 1;  // dy1s [img][pos][ic]: odd image pitch, position pitch 20 -> conflict-free col2im
constexpr int DY_BYTES = IMGS * DY_IMG * 4;
constexpr int BOFF_W2 = 0;
constexpr int BOFF_RING = BOFF_W2 + BW2_BYTES;                 // 36,864
constexpr int BOFF_DZA = BOFF_RING + BSTAGES * BSTAGE_BYTES;   // 110,592
constexpr int BOFF_DZT = BOFF_DZA + DZA_BYTES;
constexpr int BOFF_Y1 = BOFF_DZT + DZT_BYTES;
constexpr int BOFF_DY = BOFF_Y1 + Y1_BYTES;
constexpr int BOFF_XS = BOFF_DY + DY_BYTES;
constexpr int BOFF_BAR = (BOFF_XS + XS_BYTES + 15) / 16 * 16;
constexpr int BWD_SMEM = BOFF_BAR + 256 + 640 + 1024;  // barriers, conv1 weights + biases, alignment slack
constexpr int BWD_THREADS = FWD_THREADS;
constexpr int BTMEM_COLS = 512, TM_DW2 = 0, TM_DCOL = 64, TM_DCOL_STRIDE = 160;
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
    const uint32_t bar_y1f = bars, bar_y1e = bars + 8, bar_dyf = bars + 16, bar_dye = bars + 24, bar_dzf = bars + 32, bar_dze = bars + 40,
                   bar_full = bars + 48, bar_empty = bar_full + 8 * BSTAGES, bar_accf = bar_empty + 8 * BSTAGES, bar_acce = bar_accf + 16,
                   bar_dw2 = bar_acce + 16, tmem_slot = bar_dw2 + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (N + IMGS - 1) / IMGS;
    float *my_part = part + (long long)blockIdx.x * PART_FLOATS;

    if (threadIdx.x == 0) {
        mbar_init(bar_y1f, C1_WARPS * 32);
        mbar_init(bar_y1e, G_WARPS * 32);
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
    for (int i = threadIdx.x; i < 160; i += BWD_THREADS) reinterpret_cast<float *>(sm + BOFF_BAR + 256)[i] = i < 144 ? w1[i] : b1[i - 144];
    for (int i = threadIdx.x; i < IMGS * XS_IMG; i += BWD_THREADS) reinterpret_cast<float *>(sm + BOFF_XS)[i] = 0.f;
    for (int i = threadIdx.x; i < IMGS * DY_IMG; i += BWD_THREADS) reinterpret_cast<float *>(sm + BOFF_DY)[i] = 0.f;
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
            mbar_wait(bar_y1f, (uint32_t)it & 1u);
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
                        if (r >= 0 && r < 7 && cc >= 0 && cc < 7) v = lds128(smem0 + BOFF_Y1 + (uint32_t)(img * Y1_IMG + (r * 7 + cc) * 16 + c * 4) * 4u);
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
                    sts128(st + BB_PLANE + off, lo);
                }
                fence_proxy_async();
                mbar_arrive(bar_full + 8 * s);
            }
            mbar_arrive(bar_y1e);
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
        // ===== conv1 recompute, then dz1 -> dW1 / db1 =====
        const int ct = threadIdx.x - C1_WARP0 * 32;
        const int img = ct >> 5, row = (ct >> 2) & 7, q = ct & 3;
        // (the conv1 weights are re-read from shared memory per tile here: the registers hold the 40 gradient accumulators instead)
        float aw1[4][9], ab1[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < 9; ++k) aw1[c][k] = 0.f;
            ab1[c] = 0.f;
        }
        const float *wsm = reinterpret_cast<const float *>(sm + BOFF_BAR + 256) + 4 * q * 9;  // [16][9] weights, then [16] biases
        constexpr int PRE = (IMGS * 169 + C1_WARPS * 32 - 1) / (C1_WARPS * 32);
        float pre[PRE];
#pragma unroll
        for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, (long long)blockIdx.x * IMGS, ct + C1_WARPS * 32 * j);
        float *xsb = reinterpret_cast<float *>(sm + BOFF_XS);
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            bar_sync(1, C1_WARPS * 32);  // every thread is done reading the previous tile's pixels
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                const int e = ct + C1_WARPS * 32 * j;
                if (e < IMGS * 169) {
                    const int im = e / 169, c = e - im * 169, r = c / 13, qq = c - r * 13;
                    xsb[im * XS_IMG + (r + 1) * 15 + qq + 1] = pre[j];
                }
            }
            const long long next = tile + gridDim.x;
            if (next < n_tiles) {
#pragma unroll
                for (int j = 0; j < PRE; ++j) pre[j] = load_px(x, xs, N, next * IMGS, ct + C1_WARPS * 32 * j);
            }
            bar_sync(1, C1_WARPS * 32);
            mbar_wait(bar_y1e, ((uint32_t)it & 1u) ^ 1u);
            uint32_t mask = 0u;  // bit 4 j + c: y1 > 0 at column j, channel 4 q + c
            const float *xi = xsb + img * XS_IMG + (2 * row) * 15;
            if (row < 7) {
                float *yo = reinterpret_cast<float *>(sm + BOFF_Y1) + img * Y1_IMG + (row * 7) * 16 + 4 * q;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const float *bsm = reinterpret_cast<const float *>(sm + BOFF_BAR + 256) + 144 + 4 * q;
                    float a0 = bsm[0], a1 = bsm[1], a2 = bsm[2], a3 = bsm[3];
#pragma unroll
                    for (int ki = 0; ki < 3; ++ki)
#pragma unroll
                        for (int kj = 0; kj < 3; ++kj) {
                            const float v = xi[ki * 15 + 2 * j + kj];
                            a0 = fmaf(v, wsm[ki * 3 + kj], a0);
                            a1 = fmaf(v, wsm[9 + ki * 3 + kj], a1);
                            a2 = fmaf(v, wsm[18 + ki * 3 + kj], a2);
                            a3 = fmaf(v, wsm[27 + ki * 3 + kj], a3);
                        }
                    mask |= ((a0 > 0.f ? 1u : 0u) | (a1 > 0.f ? 2u : 0u) | (a2 > 0.f ? 4u : 0u) | (a3 > 0.f ? 8u : 0u)) << (4 * j);
                    *reinterpret_cast<float4 *>(yo + j * 16) = make_float4(fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f));
                }
            }
            mbar_arrive(bar_y1f);
            // dy1 of this tile (col2im done by the epilogue warps) -> dz1 = dy1 * (y1 > 0) -> dW1, db1; the cells are zeroed for the next tile
            mbar_wait_backoff(bar_dyf, (uint32_t)it & 1u);
            if (row < 7) {
                float *dy = reinterpret_cast<float *>(sm + BOFF_DY) + img * DY_IMG + (row * 7) * DY_POS + 4 * q;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    float dz[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float d = dy[j * DY_POS + c];
                        dy[j * DY_POS + c] = 0.f;
                        dz[c] = ((mask >> (4 * j + c)) & 1u) ? d : 0.f;
                        ab1[c] += dz[c];
                    }
#pragma unroll
                    for (int ki = 0; ki < 3; ++ki)
#pragma unroll
                        for (int kj = 0; kj < 3; ++kj) {
                            const float v = xi[ki * 15 + 2 * j + kj];
#pragma unroll
                            for (int c = 0; c < 4; ++c) aw1[c][ki * 3 + kj] = fmaf(dz[c], v, aw1[c][ki * 3 + kj]);
                        }
                }
            }
            mbar_arrive(bar_dye);
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
                for (int rw = 0; rw < 7; ++rw) t += red[(im * 32 + rw * 4 + qq) * 40 + c * 10 + k];
            if (k < 9) my_part[32 * 144 + ch * 9 + k] = t;
            else my_part[32 * 144 + 144 + ch] = t;
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc_col = idesc_tf32(128, 144, 0, 0), idesc_w = idesc_tf32(128, 32, 1, 1);
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
                        const uint64_t b_hi = bbase + (st16 + ((2 * BA_PLANE + k8 * 2 * BB_SBO) >> 4)), b_lo = b_hi + (BB_PLANE >> 4);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const uint64_t a_hi = abase + (st16 + ((k8 * 2 * BA_SBO + half * 2048) >> 4)), a_lo = a_hi + (BA_PLANE >> 4);
                            const uint32_t d = tmem_base + (uint32_t)(TM_DW2 + half * 32);
                            tc_mma_tf32(d, a_lo, b_hi, idesc_w, (ist > 0 || k8 > 0) ? 1u : 0u);
                            tc_mma_tf32(d, a_hi, b_lo, idesc_w, 1u);
                            tc_mma_tf32(d, a_hi, b_hi, idesc_w, 1u);
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
        float *dy = reinterpret_cast<float *>(sm + BOFF_DY) + img * DY_IMG;
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
        // dW2^T [k_idx = (tap, ic)][oc]: rows 0..127 in columns 0..31, rows 128..143 in lanes 0..15 of columns 32..63
        mbar_wait_backoff(bar_dw2, 0u);
        tc_fence_after();
        uint32_t r[32];
        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)TM_DW2, r);
        tc_wait_ld();
        {
            const int kidx = warp * 32 + lane, tap = kidx >> 4, ic = kidx & 15;
#pragma unroll
            for (int oc = 0; oc < 32; ++oc) my_part[oc * 144 + ic * 9 + tap] = __uint_as_float(r[oc]);
        }
        if (warp == 0) {
            tc_ld32(tmem_base + (uint32_t)(TM_DW2 + 32), r);
            tc_wait_ld();
            if (lane < 16) {
#pragma unroll
                for (int oc = 0; oc < 32; ++oc) my_part[oc * 144 + lane * 9 + 8] = __uint_as_float(r[oc]);
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
cudaError_t fwd_launch(const XT *x, int64_t xs, int64_t n, const float *w1, const float *b1, const float *w2, const float *b2, float *y2, cudaStream_t st) {
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
