// track2d_policy.cu -- the CNN_maze convolution stack (perception.py:68-92) as two fused fp32 kernels.
//
//   forward :  x [N][13][13] -> conv1 3x3/s2/p1 (1->16) -> ReLU -> conv2 3x3/s2/p1 (16->32) -> ReLU -> y2 [N][32*4*4]
//   backward:  given dL/dy2, accumulates dW1, db1, dW2, db2 (the observation needs no gradient); the conv1
//              activations are RECOMPUTED from x instead of being stored (7 k MAC per image vs 3 KB of HBM
//              traffic each way), so nothing but x and y2 lives between forward and backward.
//
// N is (envs x frames): 65,536 tracker images + 131,072 target (TAT) images per env-step at the benchmark size,
// i.e. a very tall, very thin problem (73.7 k MAC and ~2.7 KB per image).  cuDNN's fp32 kernels for these shapes run
// at a few percent of the FP32 pipe (profiles/); here every CTA keeps both weight tensors and a group of 8 images
// (zero-bordered, so no bounds tests) in shared memory and each thread owns a register tile:
//   conv2 fwd   thread = (image, output position), 32 output-channel accumulators, weights via broadcast LDS.128
//   dW2         thread = (4 output channels, 1 input channel), 36 accumulators that live across the whole grid-stride
//               loop; one atomicAdd per weight per CTA at the end
//   dy1         thread = (image, input channel), 49 accumulators with compile-time scatter indices
// All arithmetic is fp32 FFMA, same operation set as the reference's float32 convs; summation order differs
// (tolerance-level, tests/test_gpu_learner.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/track2d.h"

void t2d_set_error(const char *fmt, ...);
void t2d_count_launches(int n);
// track2d_conv_tc.cu: the same stack with conv2 on the tensor cores
cudaError_t t2d_conv_tc_forward(const void *x, int x_is_u8, int64_t xs, int64_t n, const float *w1, const float *b1, const float *w2, const float *b2, float *y2,
                                cudaStream_t st);
cudaError_t t2d_conv_tc_backward(const void *x, int x_is_u8, int64_t xs, const float *y2, const float *gy2, int64_t n, const float *w1, const float *b1,
                                 const float *w2, float *dw1, float *db1, float *dw2, float *db2, cudaStream_t st);

#include <stdlib.h>
#include <string.h>

namespace {

// T2D_CONV_IMPL=simt selects the FP32 CUDA-core kernels of this file (A/B measurements); default: tcgen05 (track2d_conv_tc.cu)
bool conv_use_tc() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("T2D_CONV_IMPL");
        v = (e && strcmp(e, "simt") == 0) ? 0 : 1;
    }
    return v == 1;
}

// Blackwell's packed fp32 FMA (fma.rn.f32x2 -> FFMA2): two independent round-to-nearest FMAs per instruction, i.e. the same
// arithmetic as two fmaf() at half the issue slots (a scalar FFMA occupies the FMA pipe for 2 cycles per warp either way).
__device__ __forceinline__ void ffma2(float2 &acc, const float2 a, const float2 b) {
    unsigned long long c = *reinterpret_cast<unsigned long long *>(&acc);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(*reinterpret_cast<const unsigned long long *>(&a)), "l"(*reinterpret_cast<const unsigned long long *>(&b)));
    acc = *reinterpret_cast<float2 *>(&c);
}

constexpr int IMG = 8;          // images per CTA iteration
constexpr int THREADS = 128;
constexpr int XP = 15 * 15;     // zero-bordered input
constexpr int Y1P = 81;         // zero-bordered 7x7 conv1 activation (9x9)

struct ConvSmem {
    float w2t[144 * 32];        // [ic*9+k][oc]   (forward, dW2)
    float w1[16 * 9];
    float b1[16];
    float b2[32];
    float xs[IMG][XP];
    float y1[IMG][16][Y1P];
};
struct ConvBwdSmem {
    ConvSmem f;
    float w2[32 * 144];         // [oc][ic*9+k]   (dy1)
    float dz2[IMG][16][32];     // [img][pos][oc]
};

__device__ __forceinline__ void load_weights(ConvSmem &s, const float *__restrict__ w1, const float *__restrict__ b1,
                                             const float *__restrict__ w2, const float *__restrict__ b2, int tid) {
    for (int i = tid; i < 32 * 144; i += THREADS) {
        int oc = i / 144, r = i - oc * 144;
        s.w2t[r * 32 + oc] = w2[i];
    }
    for (int i = tid; i < 144; i += THREADS) s.w1[i] = w1[i];
    if (tid < 16) s.b1[tid] = b1[tid];
    if (tid < 32) s.b2[tid] = b2 ? b2[tid] : 0.f;
    // zero borders once: interiors are overwritten every iteration, borders never are
    for (int i = tid; i < IMG * XP; i += THREADS) (&s.xs[0][0])[i] = 0.f;
    for (int i = tid; i < IMG * 16 * Y1P; i += THREADS) (&s.y1[0][0][0])[i] = 0.f;
}

// stage IMG images and run conv1 + ReLU into the zero-bordered y1 tile.  thread = (image, channel).  Images are float32 or
// uint8 (the env's lossless observation encoding: cell values 0, 1, 2, 4), `xs` elements apart.
template <typename XT>
__device__ __forceinline__ void stage_and_conv1(ConvSmem &s, const XT *__restrict__ x, int64_t xs, int64_t n0, int nimg, int tid) {
    for (int i = tid; i < IMG * 169; i += THREADS) {
        int img = i / 169, c = i - img * 169;
        int r = c / 13, q = c - r * 13;
        s.xs[img][(r + 1) * 15 + q + 1] = img < nimg ? (float)x[(n0 + img) * xs + c] : 0.f;
    }
    __syncthreads();
    const int img = tid >> 4, ch = tid & 15;
    float w[9];
#pragma unroll
    for (int k = 0; k < 9; k++) w[k] = s.w1[ch * 9 + k];
    const float b = s.b1[ch];
    const float *xi = s.xs[img];
    float *yo = s.y1[img][ch];
#pragma unroll
    for (int i = 0; i < 7; i++) {
#pragma unroll
        for (int j = 0; j < 7; j++) {
            float a = b;
#pragma unroll
            for (int ki = 0; ki < 3; ki++)
#pragma unroll
                for (int kj = 0; kj < 3; kj++) a = fmaf(xi[(2 * i + ki) * 15 + 2 * j + kj], w[ki * 3 + kj], a);
            yo[(i + 1) * 9 + j + 1] = fmaxf(a, 0.f);
        }
    }
    __syncthreads();
}

template <typename XT>
__global__ void __launch_bounds__(THREADS) maze_conv_fwd_kernel(const XT *__restrict__ x, int64_t xs, int64_t N, const float *__restrict__ w1,
                                                                const float *__restrict__ b1, const float *__restrict__ w2,
                                                                const float *__restrict__ b2, float *__restrict__ y2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ConvSmem &s = *reinterpret_cast<ConvSmem *>(smem_raw);
    const int tid = threadIdx.x;
    load_weights(s, w1, b1, w2, b2, tid);
    __syncthreads();
    // thread = (image, pair of horizontally adjacent output positions, half of the output channels): 2 x 16 accumulators.
    // Per (ic, k) that is 4 LDS.128 of weights + (shared) activations for 32 FMAs -- 6 shared-memory wavefronts per
    // 32 FMA cycles, so the FMA pipe, not the LSU, is the limiter (the 1 x 32 tiling needed 10 per 32).
    const int img = tid >> 4, sub = tid & 15, pp = sub & 7, och = sub >> 3;
    const int oi = pp >> 1, oj0 = (pp & 1) * 2;
    for (int64_t n0 = (int64_t)blockIdx.x * IMG; n0 < N; n0 += (int64_t)gridDim.x * IMG) {
        const int nimg = (int)min((int64_t)IMG, N - n0);
        stage_and_conv1(s, x, xs, n0, nimg, tid);
        float2 accA[8], accB[8];  // output-channel pairs
#pragma unroll
        for (int c = 0; c < 8; c++) accA[c] = accB[c] = make_float2(s.b2[och * 16 + 2 * c], s.b2[och * 16 + 2 * c + 1]);
        const float *yi = &s.y1[img][0][(2 * oi) * 9 + 2 * oj0];
#pragma unroll 2
        for (int ic = 0; ic < 16; ic++) {
            float v[3][5];
#pragma unroll
            for (int ki = 0; ki < 3; ki++)
#pragma unroll
                for (int cj = 0; cj < 5; cj++) v[ki][cj] = yi[ic * Y1P + ki * 9 + cj];
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const float4 *wr = reinterpret_cast<const float4 *>(&s.w2t[(ic * 9 + k) * 32 + och * 16]);
                const float2 va = make_float2(v[k / 3][k % 3], v[k / 3][k % 3]), vb = make_float2(v[k / 3][k % 3 + 2], v[k / 3][k % 3 + 2]);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 wv = wr[q];
                    ffma2(accA[2 * q + 0], va, make_float2(wv.x, wv.y));
                    ffma2(accA[2 * q + 1], va, make_float2(wv.z, wv.w));
                    ffma2(accB[2 * q + 0], vb, make_float2(wv.x, wv.y));
                    ffma2(accB[2 * q + 1], vb, make_float2(wv.z, wv.w));
                }
            }
        }
        if (img < nimg) {
            float *o = y2 + (n0 + img) * 512 + (och * 16) * 16 + oi * 4 + oj0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                *reinterpret_cast<float2 *>(o + (2 * c) * 16) = make_float2(fmaxf(accA[c].x, 0.f), fmaxf(accB[c].x, 0.f));
                *reinterpret_cast<float2 *>(o + (2 * c + 1) * 16) = make_float2(fmaxf(accA[c].y, 0.f), fmaxf(accB[c].y, 0.f));
            }
        }
        __syncthreads(); // y1 / xs are rewritten by the next iteration
    }
}

template <typename XT>
__global__ void __launch_bounds__(THREADS) maze_conv_bwd_kernel(const XT *__restrict__ x, int64_t xs, const float *__restrict__ y2,
                                                                const float *__restrict__ gy2, int64_t N, const float *__restrict__ w1,
                                                                const float *__restrict__ b1, const float *__restrict__ w2,
                                                                float *__restrict__ dw1, float *__restrict__ db1, float *__restrict__ dw2,
                                                                float *__restrict__ db2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ConvBwdSmem &s = *reinterpret_cast<ConvBwdSmem *>(smem_raw);
    const int tid = threadIdx.x;
    load_weights(s.f, w1, b1, w2, nullptr, tid);
    for (int i = tid; i < 32 * 144; i += THREADS) s.w2[i] = w2[i];
    __syncthreads();

    // persistent accumulators
    const int ocg = tid & 7, icw = tid >> 3;    // dW2 tile: output channels 4*ocg..4*ocg+3, input channel icw
    float2 aw2[9][2];  // [tap][output-channel pair]
#pragma unroll
    for (int k = 0; k < 9; k++) aw2[k][0] = aw2[k][1] = make_float2(0.f, 0.f);
    const int img = tid >> 4, ic = tid & 15;    // dy1 / dW1 tile: (image, conv1 channel)
    float aw1[9], ab1 = 0.f, ab2 = 0.f;
#pragma unroll
    for (int k = 0; k < 9; k++) aw1[k] = 0.f;

    for (int64_t n0 = (int64_t)blockIdx.x * IMG; n0 < N; n0 += (int64_t)gridDim.x * IMG) {
        const int nimg = (int)min((int64_t)IMG, N - n0);
        stage_and_conv1(s.f, x, xs, n0, nimg, tid);
        // dz2 = dL/dy2 * (y2 > 0), transposed to [img][pos][oc]
        {
            const int oc = tid & 31, q4 = tid >> 5; // 4 warps: each takes positions q4*4 .. q4*4+3
            for (int im = 0; im < IMG; im++) {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    int pos = q4 * 4 + p;
                    float d = 0.f;
                    if (im < nimg) {
                        int64_t gi = (n0 + im) * 512 + oc * 16 + pos;
                        d = y2[gi] > 0.f ? gy2[gi] : 0.f;
                    }
                    s.dz2[im][pos][oc] = d;
                }
            }
        }
        __syncthreads();

        // ---- dW2[oc][ic][k] += sum_{img,pos} dz2[img][pos][oc] * y1pad[img][ic][patch(pos, k)] ----------------
        for (int im = 0; im < IMG; im++) {
            const float *yi = s.f.y1[im][icw];
#pragma unroll
            for (int pr = 0; pr < 8; pr++) { // pairs of horizontally adjacent output positions share 3 x 5 activations
                const int oi = pr >> 1, oj0 = (pr & 1) * 2;
                const float4 dA = *reinterpret_cast<const float4 *>(&s.dz2[im][oi * 4 + oj0][ocg * 4]);
                const float4 dB = *reinterpret_cast<const float4 *>(&s.dz2[im][oi * 4 + oj0 + 1][ocg * 4]);
                float v[3][5];
#pragma unroll
                for (int ki = 0; ki < 3; ki++)
#pragma unroll
                    for (int cj = 0; cj < 5; cj++) v[ki][cj] = yi[(2 * oi + ki) * 9 + 2 * oj0 + cj];
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const float2 va = make_float2(v[k / 3][k % 3], v[k / 3][k % 3]), vb = make_float2(v[k / 3][k % 3 + 2], v[k / 3][k % 3 + 2]);
                    ffma2(aw2[k][0], make_float2(dA.x, dA.y), va);
                    ffma2(aw2[k][1], make_float2(dA.z, dA.w), va);
                    ffma2(aw2[k][0], make_float2(dB.x, dB.y), vb);
                    ffma2(aw2[k][1], make_float2(dB.z, dB.w), vb);
                }
            }
        }
        if (tid < 32) {
            for (int im = 0; im < IMG; im++)
#pragma unroll
                for (int pos = 0; pos < 16; pos++) ab2 += s.dz2[im][pos][tid];
        }

        // ---- dy1[img][ic][7x7] = sum_{oc,pos,k} dz2[img][pos][oc] * w2[oc][ic][k]; then dz1 = dy1 * (y1 > 0) ------
        float2 dy[7][4];  // rows of 7 as column pairs (0,1) (2,3) (4,5) (6,-): taps kj = 1, 2 of an output position hit one pair
#pragma unroll
        for (int r = 0; r < 7; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) dy[r][c] = make_float2(0.f, 0.f);
        for (int oc = 0; oc < 32; oc++) {
            float w[9];
#pragma unroll
            for (int k = 0; k < 9; k++) w[k] = s.w2[oc * 144 + ic * 9 + k];
#pragma unroll
            for (int pos = 0; pos < 16; pos++) {
                const float d = s.dz2[img][pos][oc];
                const float2 dd = make_float2(d, d);
#pragma unroll
                for (int ki = 0; ki < 3; ki++) {
                    const int r = 2 * (pos >> 2) + ki - 1, px = pos & 3;  // compile-time after unrolling
                    if (r >= 0 && r < 7) {
                        ffma2(dy[r][px], dd, make_float2(w[ki * 3 + 1], w[ki * 3 + 2]));               // columns 2px, 2px + 1
                        if (px > 0) dy[r][px - 1].y = fmaf(d, w[ki * 3], dy[r][px - 1].y);              // column 2px - 1
                    }
                }
            }
        }
        // ---- dW1[ch][k] += sum dz1 * xpad ; db1 ----------------------------------------------------------------
        {
            const float *yi = s.f.y1[img][ic];
            const float *xi = s.f.xs[img];
#pragma unroll
            for (int i = 0; i < 7; i++)
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const float dz = yi[(i + 1) * 9 + j + 1] > 0.f ? ((j & 1) ? dy[i][j >> 1].y : dy[i][j >> 1].x) : 0.f;
                    ab1 += dz;
#pragma unroll
                    for (int ki = 0; ki < 3; ki++)
#pragma unroll
                        for (int kj = 0; kj < 3; kj++) aw1[ki * 3 + kj] = fmaf(dz, xi[(2 * i + ki) * 15 + 2 * j + kj], aw1[ki * 3 + kj]);
                }
        }
        __syncthreads();
    }

    // ---- flush the persistent accumulators ------------------------------------------------------------------
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int k = 0; k < 9; k++) atomicAdd(&dw2[(ocg * 4 + a) * 144 + icw * 9 + k], (a & 1) ? aw2[k][a >> 1].y : aw2[k][a >> 1].x);
    if (tid < 32) atomicAdd(&db2[tid], ab2);
    // dW1 / db1: 8 threads (one per image slot) share a channel: reduce through shared memory first
    float *red = reinterpret_cast<float *>(&s.dz2[0][0][0]);
    __syncthreads();
    for (int i = tid; i < 16 * 10; i += THREADS) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 9; k++) atomicAdd(&red[ic * 10 + k], aw1[k]);
    atomicAdd(&red[ic * 10 + 9], ab1);
    __syncthreads();
    for (int i = tid; i < 16 * 9; i += THREADS) atomicAdd(&dw1[i], red[(i / 9) * 10 + i % 9]);
    if (tid < 16) atomicAdd(&db1[tid], red[tid * 10 + 9]);
}

int grid_for(int64_t N, int ctas_per_sm) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t groups = (N + IMG - 1) / IMG;
    int64_t g = (int64_t)sms * ctas_per_sm;
    return (int)(groups < g ? groups : g);
}

template <typename XT>
cudaError_t conv_fwd_launch(const XT *x, int64_t xs, int64_t n, const float *w1, const float *b1, const float *w2, const float *b2, float *y2, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(maze_conv_fwd_kernel<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvSmem));
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    maze_conv_fwd_kernel<XT><<<grid_for(n, 3), THREADS, sizeof(ConvSmem), st>>>(x, xs, n, w1, b1, w2, b2, y2);
    return cudaGetLastError();
}
template <typename XT>
cudaError_t conv_bwd_launch(const XT *x, int64_t xs, const float *y2, const float *gy2, int64_t n, const float *w1, const float *b1, const float *w2,
                            float *dw1, float *db1, float *dw2, float *db2, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(maze_conv_bwd_kernel<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConvBwdSmem));
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    maze_conv_bwd_kernel<XT><<<grid_for(n, 2), THREADS, sizeof(ConvBwdSmem), st>>>(x, xs, y2, gy2, n, w1, b1, w2, dw1, db1, dw2, db2);
    return cudaGetLastError();
}

} // namespace

extern "C" int track2d_maze_conv_forward_ex(const void *x, int32_t x_is_u8, int64_t x_stride, int64_t n_images, const float *w1, const float *b1,
                                            const float *w2, const float *b2, float *y2, void *stream) {
    if (!x || !w1 || !b1 || !w2 || !b2 || !y2 || n_images < 1 || x_stride < 169) {
        t2d_set_error("track2d_maze_conv_forward: bad argument");
        return T2D_E_INVALID;
    }
    t2d_count_launches(1);
    cudaError_t err;
    if (conv_use_tc() && x_is_u8) err = t2d_conv_tc_forward(x, x_is_u8, x_stride, n_images, w1, b1, w2, b2, y2, (cudaStream_t)stream);
    else err = x_is_u8 ? conv_fwd_launch((const uint8_t *)x, x_stride, n_images, w1, b1, w2, b2, y2, (cudaStream_t)stream)
                       : conv_fwd_launch((const float *)x, x_stride, n_images, w1, b1, w2, b2, y2, (cudaStream_t)stream);
    if (err != cudaSuccess) {
        t2d_set_error("track2d_maze_conv_forward: %s", cudaGetErrorString(err));
        return T2D_E_CUDA;
    }
    return T2D_OK;
}

extern "C" int track2d_maze_conv_backward_ex(const void *x, int32_t x_is_u8, int64_t x_stride, const float *y2, const float *gy2, int64_t n_images,
                                             const float *w1, const float *b1, const float *w2, float *dw1, float *db1, float *dw2, float *db2,
                                             void *stream) {
    if (!x || !y2 || !gy2 || !w1 || !b1 || !w2 || !dw1 || !db1 || !dw2 || !db2 || n_images < 1 || x_stride < 169) {
        t2d_set_error("track2d_maze_conv_backward: bad argument");
        return T2D_E_INVALID;
    }
    cudaError_t err;
    if (conv_use_tc() && x_is_u8) {  // (float32 images -- the autograd path -- keep the CUDA-core kernels)
        t2d_count_launches(2);
        err = t2d_conv_tc_backward(x, x_is_u8, x_stride, y2, gy2, n_images, w1, b1, w2, dw1, db1, dw2, db2, (cudaStream_t)stream);
    } else {
        t2d_count_launches(1);
        err = x_is_u8 ? conv_bwd_launch((const uint8_t *)x, x_stride, y2, gy2, n_images, w1, b1, w2, dw1, db1, dw2, db2, (cudaStream_t)stream)
                      : conv_bwd_launch((const float *)x, x_stride, y2, gy2, n_images, w1, b1, w2, dw1, db1, dw2, db2, (cudaStream_t)stream);
    }
    if (err != cudaSuccess) {
        t2d_set_error("track2d_maze_conv_backward: %s", cudaGetErrorString(err));
        return T2D_E_CUDA;
    }
    return T2D_OK;
}

extern "C" int track2d_maze_conv_forward(const float *x, int64_t n_images, const float *w1, const float *b1, const float *w2, const float *b2, float *y2,
                                         void *stream) {
    return track2d_maze_conv_forward_ex(x, 0, 169, n_images, w1, b1, w2, b2, y2, stream);
}

extern "C" int track2d_maze_conv_backward(const float *x, const float *y2, const float *gy2, int64_t n_images, const float *w1, const float *b1,
                                          const float *w2, float *dw1, float *db1, float *dw2, float *db2, void *stream) {
    return track2d_maze_conv_backward_ex(x, 0, 169, y2, gy2, n_images, w1, b1, w2, dw1, db1, dw2, db2, stream);
}
