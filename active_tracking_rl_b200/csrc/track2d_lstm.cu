// track2d_lstm.cu -- the pointwise half of nn.LSTMCell (model.py:116,176 `self.lstm(feature, (hx, cx))`) for all envs, forward
// and backward, one kernel each way:
//
//   gates = igates + hgates + b_ih + b_hh = [i | f | g | o]            (igates = x W_ih^T, hgates = h W_hh^T: track2d_gemm_tf32x3)
//   i, f, o = sigmoid(.), g = tanh(.);   c' = f c + i g;   h' = o tanh(c')
//
// backward, given dL/dh' and dL/dc':  dgates (the gradient of BOTH gate GEMM outputs), dL/dc, and the bias gradient
// (= column sums of dgates) accumulated in registers while the rows stream through, so the two 65,536 x 512 reductions ATen
// launches per cell per step (b_ih and b_hh get the same sum) disappear.  Memory-bound: 7.7 KB per row forward, 6.7 KB
// backward; every access is a coalesced 16-byte vector.  The column sums are combined in a fixed order (per-CTA partials, then
// one pass over them), so gradients are bit-reproducible.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/track2d.h"

void t2d_set_error(const char *fmt, ...);
void t2d_count_launches(int n);

namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void st4_stream(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

// thread = (row, 4 consecutive hidden units); H / 4 threads cover a row, THREADS / (H / 4) rows per CTA pass
template <int H>
__global__ void __launch_bounds__(THREADS) lstm_cell_fwd_kernel(const float *__restrict__ ig, const float *__restrict__ hg, const float *__restrict__ b_ih,
                                                                const float *__restrict__ b_hh, const float *__restrict__ cx, long long cx_stride,
                                                                float *__restrict__ hy, float *__restrict__ cy, float *__restrict__ act, long long E) {
    constexpr int CPR = H / 4, RPB = THREADS / CPR;
    const int chunk = threadIdx.x % CPR, rloc = threadIdx.x / CPR;
    float4 b[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 u = ldg4(b_ih + g * H + 4 * chunk), v = ldg4(b_hh + g * H + 4 * chunk);
        b[g] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
    for (long long row = (long long)blockIdx.x * RPB + rloc; row < E; row += (long long)gridDim.x * RPB) {
        float4 z[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 u = ld4(ig + row * 4 * H + g * H + 4 * chunk), v = ld4(hg + row * 4 * H + g * H + 4 * chunk);
            z[g] = make_float4(u.x + v.x + b[g].x, u.y + v.y + b[g].y, u.z + v.z + b[g].z, u.w + v.w + b[g].w);
        }
        const float4 c = ld4(cx + row * cx_stride + 4 * chunk);
        float4 gi, gf, gg, go, c2, h2;
#define T2D_CELL(m)                                  \
    gi.m = sigmoidf_(z[0].m);                         \
    gf.m = sigmoidf_(z[1].m);                         \
    gg.m = tanhf(z[2].m);                             \
    go.m = sigmoidf_(z[3].m);                         \
    c2.m = gf.m * c.m + gi.m * gg.m;                  \
    h2.m = go.m * tanhf(c2.m);
        T2D_CELL(x) T2D_CELL(y) T2D_CELL(z) T2D_CELL(w)
#undef T2D_CELL
        st4(hy + row * H + 4 * chunk, h2);
        st4(cy + row * H + 4 * chunk, c2);
        float *a = act + row * 4 * H + 4 * chunk;  // activated gates, kept for the backward
        st4(a, gi); st4(a + H, gf); st4(a + 2 * H, gg); st4(a + 3 * H, go);
    }
}

template <int H>
__global__ void __launch_bounds__(THREADS) lstm_cell_bwd_kernel(const float *__restrict__ dhy, long long dhy_stride, const float *__restrict__ dcy, long long dcy_stride,
                                                                const float *__restrict__ cx, long long cx_stride, const float *__restrict__ cy, const float *__restrict__ act,
                                                                float *__restrict__ dgates, float *__restrict__ dcx, float *__restrict__ bias_part,
                                                                long long E) {
    constexpr int CPR = H / 4, RPB = THREADS / CPR;
    const int chunk = threadIdx.x % CPR, rloc = threadIdx.x / CPR;
    float4 sum[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) sum[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long row = (long long)blockIdx.x * RPB + rloc; row < E; row += (long long)gridDim.x * RPB) {
        const float4 dh = dhy ? ld4(dhy + row * dhy_stride + 4 * chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 dc = dcy ? ld4(dcy + row * dcy_stride + 4 * chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 c = ld4(cx + row * cx_stride + 4 * chunk), c2 = ld4(cy + row * H + 4 * chunk);
        const float *a = act + row * 4 * H + 4 * chunk;
        const float4 gi = ld4(a), gf = ld4(a + H), gg = ld4(a + 2 * H), go = ld4(a + 3 * H);
        float4 di, df, dg, dO, dcp;
#define T2D_CELL(m)                                                        \
    {                                                                      \
        const float tc = tanhf(c2.m);                                      \
        const float dct = dc.m + dh.m * go.m * (1.f - tc * tc);            \
        dO.m = dh.m * tc * go.m * (1.f - go.m);                            \
        di.m = dct * gg.m * gi.m * (1.f - gi.m);                           \
        df.m = dct * c.m * gf.m * (1.f - gf.m);                            \
        dg.m = dct * gi.m * (1.f - gg.m * gg.m);                           \
        dcp.m = dct * gf.m;                                                \
    }
        T2D_CELL(x) T2D_CELL(y) T2D_CELL(z) T2D_CELL(w)
#undef T2D_CELL
        float *o = dgates + row * 4 * H + 4 * chunk;
        st4(o, di); st4(o + H, df); st4(o + 2 * H, dg); st4(o + 3 * H, dO);
        st4(dcx + row * H + 4 * chunk, dcp);
        sum[0].x += di.x; sum[0].y += di.y; sum[0].z += di.z; sum[0].w += di.w;
        sum[1].x += df.x; sum[1].y += df.y; sum[1].z += df.z; sum[1].w += df.w;
        sum[2].x += dg.x; sum[2].y += dg.y; sum[2].z += dg.z; sum[2].w += dg.w;
        sum[3].x += dO.x; sum[3].y += dO.y; sum[3].z += dO.z; sum[3].w += dO.w;
    }
    // per-CTA column sums: the RPB row slots of a column are added in slot order
    __shared__ float4 red[THREADS][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) red[threadIdx.x][g] = sum[g];
    __syncthreads();
    if (rloc == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float4 t = red[chunk][g];
            for (int r = 1; r < RPB; ++r) {
                const float4 u = red[r * CPR + chunk][g];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            st4(bias_part + (long long)blockIdx.x * 4 * H + g * H + 4 * chunk, t);
        }
    }
}

// db[j] = sum over CTAs of bias_part[cta][j]: 32 columns per block, 8 interleaved slices of the partials per column, combined
// in a fixed order
__global__ void __launch_bounds__(256) lstm_bias_reduce_kernel(const float *__restrict__ part, int n_part, int n, float *__restrict__ db) {
    __shared__ float red[8][32];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
    float t = 0.f;
    if (col < n)
        for (int p = slice; p < n_part; p += 8) t += part[(long long)p * n + col];
    red[slice][threadIdx.x & 31] = t;
    __syncthreads();
    if (slice == 0 && col < n) {
#pragma unroll
        for (int s2 = 1; s2 < 8; ++s2) t += red[s2][threadIdx.x];
        db[col] = t;
    }
}

// out[n] = sum_m x[m][n] for a tall row-major matrix (bias gradients of the Linear layers): thread = (row slot, 4 columns),
// per-CTA partials, then lstm_bias_reduce_kernel -- fixed summation order
__global__ void __launch_bounds__(THREADS) colsum_partial_kernel(const float *__restrict__ x, long long ld, long long M, int N, float *__restrict__ part) {
    const int cpr = N >> 2, rpb = THREADS / cpr;
    const int chunk = threadIdx.x % cpr, rloc = threadIdx.x / cpr;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rloc < rpb)
        for (long long row = (long long)blockIdx.x * rpb + rloc; row < M; row += (long long)gridDim.x * rpb) {
            const float4 v = ld4(x + row * ld + 4 * chunk);
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
    __shared__ float4 red[THREADS];
    red[threadIdx.x] = sum;
    __syncthreads();
    if (rloc == 0) {
        float4 t = red[chunk];
        for (int r = 1; r < rpb; ++r) {
            const float4 u = red[r * cpr + chunk];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        st4(part + (long long)blockIdx.x * N + 4 * chunk, t);
    }
}

int colsum_grid(long long M, int N) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long rpb = THREADS / (N / 4), groups = (M + rpb - 1) / rpb;
    const long long g = (long long)sms * 4;
    return (int)(groups < g ? groups : g);
}

int lstm_grid(long long E, int H) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long rpb = THREADS / (H / 4), groups = (E + rpb - 1) / rpb;
    const long long g = (long long)sms * 4;
    return (int)(groups < g ? groups : g);
}

bool bad_ptr(const void *p) { return ((uintptr_t)p & 15) != 0; }

}  // namespace

extern "C" int64_t track2d_lstm_bias_workspace_floats(int64_t E, int32_t H) {
    if (E <= 0 || H != 128) return 0;
    return (int64_t)lstm_grid(E, H) * 4 * H;
}

extern "C" int track2d_lstm_cell_forward(const float *igates_dev, const float *hgates_dev, const float *b_ih_dev, const float *b_hh_dev,
                                         const float *cx_dev, int64_t cx_stride, float *hy_dev, float *cy_dev, float *act_dev, int64_t E, int32_t H,
                                         void *stream) {
    if (!igates_dev || !hgates_dev || !b_ih_dev || !b_hh_dev || !cx_dev || !hy_dev || !cy_dev || !act_dev || E <= 0) {
        t2d_set_error("track2d_lstm_cell_forward: bad argument");
        return T2D_E_INVALID;
    }
    if (H != 128 || cx_stride % 4 || bad_ptr(igates_dev) || bad_ptr(hgates_dev) || bad_ptr(b_ih_dev) || bad_ptr(b_hh_dev) || bad_ptr(cx_dev) ||
        bad_ptr(hy_dev) || bad_ptr(cy_dev) || bad_ptr(act_dev)) {
        t2d_set_error("track2d_lstm_cell_forward: hidden size must be 128 (rnn_out of the 2D configurations) and all buffers 16-byte aligned");
        return T2D_E_INVALID;
    }
    lstm_cell_fwd_kernel<128><<<lstm_grid(E, H), THREADS, 0, (cudaStream_t)stream>>>(igates_dev, hgates_dev, b_ih_dev, b_hh_dev, cx_dev, cx_stride, hy_dev,
                                                                                    cy_dev, act_dev, E);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        t2d_set_error("track2d_lstm_cell_forward: %s", cudaGetErrorString(e));
        return T2D_E_CUDA;
    }
    t2d_count_launches(1);
    return T2D_OK;
}

extern "C" int track2d_lstm_cell_backward(const float *dhy_dev, int64_t dhy_stride, const float *dcy_dev, int64_t dcy_stride, const float *cx_dev,
                                          int64_t cx_stride, const float *cy_dev,
                                          const float *act_dev, float *dgates_dev, float *dcx_dev, float *dbias_dev, float *workspace_dev,
                                          int64_t workspace_floats, int64_t E, int32_t H, void *stream) {
    if (!cx_dev || !cy_dev || !act_dev || !dgates_dev || !dcx_dev || !dbias_dev || !workspace_dev || E <= 0) {
        t2d_set_error("track2d_lstm_cell_backward: bad argument");
        return T2D_E_INVALID;
    }
    if (H != 128 || cx_stride % 4 || dhy_stride % 4 || dcy_stride % 4 || bad_ptr(dhy_dev) || bad_ptr(dcy_dev) || bad_ptr(cx_dev) || bad_ptr(cy_dev) || bad_ptr(act_dev) || bad_ptr(dgates_dev) ||
        bad_ptr(dcx_dev) || bad_ptr(dbias_dev) || bad_ptr(workspace_dev)) {
        t2d_set_error("track2d_lstm_cell_backward: hidden size must be 128 and all buffers 16-byte aligned");
        return T2D_E_INVALID;
    }
    const int grid = lstm_grid(E, H);
    if (workspace_floats < (int64_t)grid * 4 * H) {
        t2d_set_error("track2d_lstm_cell_backward: workspace of %lld floats needed (track2d_lstm_bias_workspace_floats)", (long long)grid * 4 * H);
        return T2D_E_INVALID;
    }
    lstm_cell_bwd_kernel<128><<<grid, THREADS, 0, (cudaStream_t)stream>>>(dhy_dev, dhy_stride, dcy_dev, dcy_stride, cx_dev, cx_stride, cy_dev, act_dev,
                                                                         dgates_dev, dcx_dev, workspace_dev, E);
    lstm_bias_reduce_kernel<<<(4 * H + 31) / 32, 256, 0, (cudaStream_t)stream>>>(workspace_dev, grid, 4 * H, dbias_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        t2d_set_error("track2d_lstm_cell_backward: %s", cudaGetErrorString(e));
        return T2D_E_CUDA;
    }
    t2d_count_launches(2);
    return T2D_OK;
}

extern "C" int64_t track2d_colsum_workspace_floats(int64_t M, int32_t N) {
    if (M <= 0 || N < 4 || N > 1024 || (N & (N - 1))) return 0;
    return (int64_t)colsum_grid(M, N) * N;
}

extern "C" int track2d_colsum(const float *x_dev, int64_t ld, int64_t M, int32_t N, float *out_dev, float *workspace_dev, int64_t workspace_floats,
                              void *stream) {
    if (!x_dev || !out_dev || !workspace_dev || M <= 0 || N < 4 || N > 1024 || (N & (N - 1)) || ld % 4 || bad_ptr(x_dev) || bad_ptr(out_dev) ||
        bad_ptr(workspace_dev)) {
        t2d_set_error("track2d_colsum: N must be a power of two in [4, 1024], buffers 16-byte aligned, ld a multiple of 4");
        return T2D_E_INVALID;
    }
    const int grid = colsum_grid(M, N);
    if (workspace_floats < (int64_t)grid * N) {
        t2d_set_error("track2d_colsum: workspace of %lld floats needed (track2d_colsum_workspace_floats)", (long long)grid * N);
        return T2D_E_INVALID;
    }
    colsum_partial_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(x_dev, ld, M, N, workspace_dev);
    lstm_bias_reduce_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(workspace_dev, grid, N, out_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        t2d_set_error("track2d_colsum: %s", cudaGetErrorString(e));
        return T2D_E_CUDA;
    }
    t2d_count_launches(2);
    return T2D_OK;
}
