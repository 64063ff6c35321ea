// track2d_a3c.cu -- the recurrent half of the policy step and the A3C loss, fused (reference: model.py:41-50 sample_action,
// :116-127 / :175-209 LSTMCell + heads, player_util.py:108-161 optimize).
//
// Forward, per agent and env-step, ONE kernel after the gate GEMM (gates = [feature | h] [W_ih | W_hh]^T, K = 384):
//     LSTM cell pointwise -> actor / critic / reward_aux heads (8 x 128 dot products per row) -> softmax, log-softmax, entropy,
//     multinomial sample (counter-based Philox) or argmax or a forced action -> action, value, log pi(a), entropy
// so the ~50 ATen launches per step of `sample_action`, the head GEMMs, cat / stack / one-hot / masking disappear.
//
// Backward: `a3c_loss_grad_kernel` runs the n-step return / GAE recursions (player_util.py:127-140, cut at episode ends) and
// turns them straight into dL/d(logits, value, reward prediction) for every (t, env, agent) plus the per-env loss statistics
// the reference logs; `lstm_heads_bwd_kernel` is the backward of the forward kernel for one step of the BPTT sweep (the head
// gradient enters as 8 x 128 FMAs per row, the recurrent gradient is masked where the episode ended).  Weight gradients are
// NOT accumulated here: they are batched over all T x E rows afterwards as long-K tensor-core GEMMs / column sums (learner.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/track2d.h"

void t2d_set_error(const char *fmt, ...);
void t2d_count_launches(int n);

namespace {

constexpr int H = 128;        // rnn_out of the 2D configurations (main.py:44)
constexpr int NOUT = 8;       // packed head outputs per row: 4 logits, value, reward prediction (TAT), 2 x padding
constexpr int WARPS = 8;      // rows per CTA pass

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))); }

// Philox4x32-10, one block: key = seed, counter = (row, stream, step_lo, step_hi)
__device__ __forceinline__ uint32_t philox_first(unsigned long long seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return x0;
}

struct FwdArgs {
    const float *gates;      // [E][4H]  x W_ih^T + h W_hh^T (no bias)
    const float *b_ih, *b_hh;
    const float *c_prev;     // [E][H]
    float *act;              // [E][4H]  activated gates (i, f, g, o), kept for the backward; may be NULL (no-grad step)
    float *c_next;           // [E][H]
    float *h_out;            // [E][H]   may be NULL
    float *h_next;           // row stride h_next_ld: the recurrent input slot of the next step; may be NULL
    long long h_next_ld;
    const float *w_head;     // [8][H]: actor (4), critic, reward_aux or 0, 0, 0
    const float *b_head;     // [8]
    float *out8;             // [E][8]   may be NULL
    int32_t *action;         // element stride 2: column `agent` of an int32 [E][2] array
    const int32_t *forced;   // same layout; NULL = sample / argmax
    float *value, *logp, *entropy;  // element stride 2 each (columns of [E][2] arrays); may be NULL
    float *logp_all;         // [E][4] log-probabilities (test mode, model.py:45-46); may be NULL
    const unsigned long long *rng_step;  // device-resident step counter (advanced by post_step)
    unsigned long long seed;
    uint32_t stream;         // Philox stream: agent | bootstrap << 1
    int greedy;
    long long E;
    long long first_row;     // index of row 0 in the whole batch: the Philox counter of row r is first_row + r (a batch processed in slices samples the same actions)
};

// one warp per row; lane = 4 consecutive hidden units
__global__ void __launch_bounds__(WARPS * 32) lstm_heads_fwd_kernel(const FwdArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 b[4], wh[NOUT];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 u = ldg4(a.b_ih + g * H + 4 * lane), v = ldg4(a.b_hh + g * H + 4 * lane);
        b[g] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
#pragma unroll
    for (int k = 0; k < NOUT; ++k) wh[k] = ldg4(a.w_head + k * H + 4 * lane);
    const float bh = lane < NOUT ? __ldg(a.b_head + lane) : 0.f;
    const unsigned long long step = a.forced || a.greedy ? 0ull : *a.rng_step;
    for (long long row = (long long)blockIdx.x * WARPS + warp; row < a.E; row += (long long)gridDim.x * WARPS) {
        float4 z[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 u = ld4(a.gates + row * 4 * H + g * H + 4 * lane);
            z[g] = make_float4(u.x + b[g].x, u.y + b[g].y, u.z + b[g].z, u.w + b[g].w);
        }
        const float4 c = ld4(a.c_prev + row * H + 4 * lane);
        float4 gi, gf, gg, go, c2, h2;
#define T2D_CELL(m)                 \
    gi.m = sigmoidf_(z[0].m);        \
    gf.m = sigmoidf_(z[1].m);        \
    gg.m = tanhf(z[2].m);            \
    go.m = sigmoidf_(z[3].m);        \
    c2.m = gf.m * c.m + gi.m * gg.m; \
    h2.m = go.m * tanhf(c2.m);
        T2D_CELL(x) T2D_CELL(y) T2D_CELL(z) T2D_CELL(w)
#undef T2D_CELL
        if (a.act) {
            float *o = a.act + row * 4 * H + 4 * lane;
            st4(o, gi); st4(o + H, gf); st4(o + 2 * H, gg); st4(o + 3 * H, go);
        }
        st4(a.c_next + row * H + 4 * lane, c2);
        if (a.h_out) st4(a.h_out + row * H + 4 * lane, h2);
        if (a.h_next) st4(a.h_next + row * a.h_next_ld + 4 * lane, h2);
        // heads: 8 dot products of length 128, butterfly-reduced (every lane ends with all 8)
        float out[NOUT];
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            float p = dot4(h2, wh[k]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xFFFFFFFFu, p, o);
            out[k] = p + __shfl_sync(0xFFFFFFFFu, bh, k);
        }
        if (a.out8 && lane < NOUT) {
            float v = out[0];
#pragma unroll
            for (int k = 1; k < NOUT; ++k) v = lane == k ? out[k] : v;
            a.out8[row * NOUT + lane] = v;
        }
        if (lane == 0) {
            // model.py:41-50: softmax, log_softmax, entropy, multinomial / argmax
            const float m = fmaxf(fmaxf(out[0], out[1]), fmaxf(out[2], out[3]));
            const float e0 = expf(out[0] - m), e1 = expf(out[1] - m), e2 = expf(out[2] - m), e3 = expf(out[3] - m);
            const float s = e0 + e1 + e2 + e3, ls = logf(s);
            const float p0 = e0 / s, p1 = e1 / s, p2 = e2 / s, p3 = e3 / s;
            const float l0 = out[0] - m - ls, l1 = out[1] - m - ls, l2 = out[2] - m - ls, l3 = out[3] - m - ls;
            int act;
            if (a.forced) {
                act = a.forced[row * 2] & 3;
            } else if (a.greedy) {
                act = 0;
                float best = p0;
                if (p1 > best) { best = p1; act = 1; }
                if (p2 > best) { best = p2; act = 2; }
                if (p3 > best) { best = p3; act = 3; }
            } else {
                const long long grow = a.first_row + row;
                const uint32_t r = philox_first(a.seed, (uint32_t)grow, a.stream | ((uint32_t)(grow >> 32) << 8), (uint32_t)step, (uint32_t)(step >> 32));
                const float u = (float)(r >> 8) * (1.0f / 16777216.0f);  // [0, 1)
                act = u < p0 ? 0 : (u < p0 + p1 ? 1 : (u < p0 + p1 + p2 ? 2 : 3));
            }
            a.action[row * 2] = act;
            if (a.value) a.value[row * 2] = out[4];
            if (a.logp) a.logp[row * 2] = act == 0 ? l0 : (act == 1 ? l1 : (act == 2 ? l2 : l3));
            if (a.entropy) a.entropy[row * 2] = -(l0 * p0 + l1 * p1 + l2 * p2 + l3 * p3);
            if (a.logp_all) st4(a.logp_all + row * 4, make_float4(l0, l1, l2, l3));
        }
    }
}

// after env.step: envs that finished start their next step from zero recurrent state (train.py:73-74 -> player.reset());
// advance the sampling counter.  One warp per env, lane = 4 hidden units.
__global__ void __launch_bounds__(256) post_step_kernel(const uint8_t *__restrict__ done, float *h0, float *h1, long long h_ld, float *c0, float *c1,
                                                        int32_t *eps_len, unsigned long long *rng_step, long long E) {
    const int lane = threadIdx.x & 31;
    const long long e = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0 && rng_step) *rng_step += 1ull;
    if (e >= E) return;
    const bool d = done && done[e] != 0;
    if (eps_len && lane == 0) eps_len[e] = d ? 0 : eps_len[e] + 1;
    if (!d) return;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h0) st4(h0 + e * h_ld + 4 * lane, zero);
    if (h1) st4(h1 + e * h_ld + 4 * lane, zero);
    if (c0) st4(c0 + e * H + 4 * lane, zero);
    if (c1) st4(c1 + e * H + 4 * lane, zero);
}

// out[e][j] = in[e][j] + W[j][a[e]] + b[j]: the TAT's tracker-action embedding (model.py:198-199 fc_action_tracker on a one-hot)
__global__ void __launch_bounds__(256) embed_add_kernel(const float *__restrict__ x, long long ld, float *__restrict__ out, long long out_ld,
                                                        const float *__restrict__ w /*[N][4]*/, const float *__restrict__ b,
                                                        const int32_t *__restrict__ a /*stride 2*/, int N, long long E) {
    const int cpr = N >> 2;
    const long long total = E * cpr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cpr;
        const int j = (int)(i - row * cpr) * 4;
        const int act = a[row * 2] & 3;
        float4 v = ld4(x + row * ld + j);
        const float4 bb = ldg4(b + j);
        v.x += __ldg(w + (j + 0) * 4 + act) + bb.x;
        v.y += __ldg(w + (j + 1) * 4 + act) + bb.y;
        v.z += __ldg(w + (j + 2) * 4 + act) + bb.z;
        v.w += __ldg(w + (j + 3) * 4 + act) + bb.w;
        st4(out + row * out_ld + j, v);
    }
}

struct LossArgs {
    const float *out8[2];    // [T+1][E][8] per agent (slot T = the bootstrap forward)
    float *dout8[2];         // [T][E][8]
    const int32_t *actions;  // [T][E][2]
    const float *rewards;    // [T][E][2]
    const uint8_t *done;     // [T][E]
    float *stats;            // [7][E]: policy_loss 0/1, value_loss 0/1, entropy 0/1, pred_loss
    float *returns, *gae;    // [T][E][2] optional outputs
    int T;
    long long E;
    float gamma, tau, w_ent[2], scale;
    int train[2];            // agent's loss is part of the objective (training_mode -1 / 0 / 1)
    int use_aux;             // reward_aux L1 term: 0 = the net has no such head, 1 = statistic only, 2 = part of the objective
};

// thread = (env, agent), t = T-1 .. 0.  player_util.py:117-145.
__global__ void __launch_bounds__(256) a3c_loss_grad_kernel(const LossArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * a.E) return;
    const long long e = i >> 1;
    const int ag = (int)(i & 1);
    const float *o8 = a.out8[ag];
    float *d8 = a.dout8[ag];
    float R = o8[((long long)a.T * a.E + e) * NOUT + 4], vnext = R, gae = 0.f;
    float pl = 0.f, vl = 0.f, ent = 0.f, prl = 0.f;
    const float w = a.w_ent[ag], sc = a.train[ag] ? a.scale : 0.f;
    const float sc_aux = (ag == 1 && a.use_aux == 2) ? a.scale : 0.f;
    for (int t = a.T - 1; t >= 0; --t) {
        const long long k = (long long)t * a.E + e;
        const float4 z = ld4(o8 + k * NOUT), vp = ld4(o8 + k * NOUT + 4);
        const float v = vp.x, pred = vp.y;
        const float r = a.rewards[k * 2 + ag];
        if (a.done[k]) { R = 0.f; vnext = 0.f; gae = 0.f; }
        R = a.gamma * R + r;
        const float delta = r + a.gamma * vnext - v;
        gae = gae * a.gamma * a.tau + delta;
        vnext = v;
        if (a.returns) { a.returns[k * 2 + ag] = R; a.gae[k * 2 + ag] = gae; }
        const float m = fmaxf(fmaxf(z.x, z.y), fmaxf(z.z, z.w));
        const float e0 = expf(z.x - m), e1 = expf(z.y - m), e2 = expf(z.z - m), e3 = expf(z.w - m);
        const float s = e0 + e1 + e2 + e3, ls = logf(s);
        const float p0 = e0 / s, p1 = e1 / s, p2 = e2 / s, p3 = e3 / s;
        const float l0 = z.x - m - ls, l1 = z.y - m - ls, l2 = z.z - m - ls, l3 = z.w - m - ls;
        const float Hn = -(l0 * p0 + l1 * p1 + l2 * p2 + l3 * p3);
        const int act = a.actions[k * 2 + ag] & 3;
        const float lp = act == 0 ? l0 : (act == 1 ? l1 : (act == 2 ? l2 : l3));
        const float adv = R - v;
        vl += 0.5f * adv * adv;          // :133
        pl += -(lp * gae) - w * Hn;      // :137-139
        ent += Hn;
        // d/dlogit_j [-lp * gae - w * H] = -gae * (1[j == a] - p_j) + w * p_j * (log p_j + H)
        float4 dz;
        dz.x = sc * (-gae * ((act == 0 ? 1.f : 0.f) - p0) + w * p0 * (l0 + Hn));
        dz.y = sc * (-gae * ((act == 1 ? 1.f : 0.f) - p1) + w * p1 * (l1 + Hn));
        dz.z = sc * (-gae * ((act == 2 ? 1.f : 0.f) - p2) + w * p2 * (l2 + Hn));
        dz.w = sc * (-gae * ((act == 3 ? 1.f : 0.f) - p3) + w * p3 * (l3 + Hn));
        float dpred = 0.f;
        if (ag == 1 && a.use_aux) {      // L1 between the target's prediction and the TRACKER's reward (:128-129)
            const float diff = pred - a.rewards[k * 2];
            prl += fabsf(diff);
            dpred = sc_aux * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
        }
        st4(d8 + k * NOUT, dz);
        st4(d8 + k * NOUT + 4, make_float4(sc * 0.5f * (v - R), dpred, 0.f, 0.f));  // loss has 0.5 * value_loss (:143)
    }
    a.stats[(0 + ag) * a.E + e] = pl;
    a.stats[(2 + ag) * a.E + e] = vl;
    a.stats[(4 + ag) * a.E + e] = ent;
    if (ag == 1) a.stats[6 * a.E + e] = prl;
}

struct BwdArgs {
    const float *dout8;      // [E][8] gradient of the packed head outputs at this step
    const float *w_head;     // [8][H]
    const float *dh_rec;     // [E][H] gradient arriving from step t+1 through h (NULL at the last step)
    float *dc;               // [E][H] in: gradient arriving from step t+1 through c (ignored when dh_rec == NULL); out: dL/dc_prev
    const uint8_t *done;     // [E] episode ended at THIS step: the recurrent gradient does not flow back through the reset
    const float *act;        // [E][4H]
    const float *c_prev;     // [E][H]
    float *dgates;           // [E][4H]
    long long E;
};

__global__ void __launch_bounds__(WARPS * 32) lstm_heads_bwd_kernel(const BwdArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 wh[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) wh[k] = ldg4(a.w_head + k * H + 4 * lane);
    for (long long row = (long long)blockIdx.x * WARPS + warp; row < a.E; row += (long long)gridDim.x * WARPS) {
        const float4 d0 = ld4(a.dout8 + row * NOUT), d1 = ld4(a.dout8 + row * NOUT + 4);
        const float dk[NOUT] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        float4 dh = make_float4(0.f, 0.f, 0.f, 0.f), dc = dh;
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            dh.x = fmaf(dk[k], wh[k].x, dh.x); dh.y = fmaf(dk[k], wh[k].y, dh.y);
            dh.z = fmaf(dk[k], wh[k].z, dh.z); dh.w = fmaf(dk[k], wh[k].w, dh.w);
        }
        if (a.dh_rec && !a.done[row]) {
            const float4 r = ld4(a.dh_rec + row * H + 4 * lane);
            dh.x += r.x; dh.y += r.y; dh.z += r.z; dh.w += r.w;
            dc = ld4(a.dc + row * H + 4 * lane);
        }
        const float4 c = ld4(a.c_prev + row * H + 4 * lane);
        const float *g = a.act + row * 4 * H + 4 * lane;
        const float4 gi = ld4(g), gf = ld4(g + H), gg = ld4(g + 2 * H), go = ld4(g + 3 * H);
        float4 di, df, dg, dO, dcp;
#define T2D_CELL(m)                                              \
    {                                                            \
        const float tc = tanhf(gf.m * c.m + gi.m * gg.m);        \
        const float dct = dc.m + dh.m * go.m * (1.f - tc * tc);  \
        dO.m = dh.m * tc * go.m * (1.f - go.m);                  \
        di.m = dct * gg.m * gi.m * (1.f - gi.m);                 \
        df.m = dct * c.m * gf.m * (1.f - gf.m);                  \
        dg.m = dct * gi.m * (1.f - gg.m * gg.m);                 \
        dcp.m = dct * gf.m;                                      \
    }
        T2D_CELL(x) T2D_CELL(y) T2D_CELL(z) T2D_CELL(w)
#undef T2D_CELL
        float *o = a.dgates + row * 4 * H + 4 * lane;
        st4(o, di); st4(o + H, df); st4(o + 2 * H, dg); st4(o + 3 * H, dO);
        st4(a.dc + row * H + 4 * lane, dcp);
    }
}

// In place dz = dy * (y > 0) over a tall [M][N] matrix (ReLU backward of the encoder's fc), with, in the same pass,
//   colsum(dz)                  -> the fc bias gradient
//   per-action colsum(dy), dy   -> the gradients of fc_action_tracker (weight [N][4], bias [N]) whose output was added AFTER the ReLU
// per-CTA partials [grid][5][N], combined in a fixed order by groupsum_reduce_kernel.
constexpr int RG_THREADS = 256;
__global__ void __launch_bounds__(RG_THREADS) relu_bwd_groupsum_kernel(float *__restrict__ dy, const float *__restrict__ y, long long y_ld,
                                                                       const int32_t *__restrict__ grp /*stride 2*/, long long M, int N, float *__restrict__ part) {
    const int cpr = N >> 2, rpb = RG_THREADS / cpr;
    const int chunk = threadIdx.x % cpr, rloc = threadIdx.x / cpr;
    float4 s[5];
#pragma unroll
    for (int g = 0; g < 5; ++g) s[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rloc < rpb)
        for (long long row = (long long)blockIdx.x * rpb + rloc; row < M; row += (long long)gridDim.x * rpb) {
            float4 v = ld4(dy + row * N + 4 * chunk);
            const float4 yy = ld4(y + row * y_ld + 4 * chunk);
            if (grp) {
                const int g = grp[row * 2] & 3;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (g == q) { s[1 + q].x += v.x; s[1 + q].y += v.y; s[1 + q].z += v.z; s[1 + q].w += v.w; }
            }
            v.x = yy.x > 0.f ? v.x : 0.f; v.y = yy.y > 0.f ? v.y : 0.f; v.z = yy.z > 0.f ? v.z : 0.f; v.w = yy.w > 0.f ? v.w : 0.f;
            st4(dy + row * N + 4 * chunk, v);
            s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
        }
    __shared__ float4 red[RG_THREADS];
    for (int g = 0; g < (grp ? 5 : 1); ++g) {
        __syncthreads();
        red[threadIdx.x] = s[g];
        __syncthreads();
        if (rloc == 0) {
            float4 t = red[chunk];
            for (int r = 1; r < rpb; ++r) {
                const float4 u = red[r * cpr + chunk];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            st4(part + ((long long)blockIdx.x * 5 + g) * N + 4 * chunk, t);
        }
    }
}

// column j: db[j] = sum_p part[p][0][j];  gw[j][a] = sum_p part[p][1 + a][j];  gb[j] = sum_a gw[j][a]
__global__ void __launch_bounds__(256) groupsum_reduce_kernel(const float *__restrict__ part, int n_part, int N, int groups, float *__restrict__ db,
                                                              float *__restrict__ gw, float *__restrict__ gb) {
    __shared__ float red[8][32];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
    float tot = 0.f;
    for (int g = 0; g < (groups ? 5 : 1); ++g) {
        float t = 0.f;
        if (col < N)
            for (int p = slice; p < n_part; p += 8) t += part[((long long)p * 5 + g) * N + col];
        __syncthreads();
        red[slice][threadIdx.x & 31] = t;
        __syncthreads();
        if (slice == 0 && col < N) {
#pragma unroll
            for (int s2 = 1; s2 < 8; ++s2) t += red[s2][threadIdx.x];
            if (g == 0) db[col] = t;
            else { gw[col * 4 + (g - 1)] = t; tot += t; }
        }
    }
    if (groups && slice == 0 && col < N) gb[col] = tot;
}

int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
int row_grid(long long rows, int rows_per_cta, int ctas_per_sm) {
    const long long groups = (rows + rows_per_cta - 1) / rows_per_cta, cap = (long long)sm_count() * ctas_per_sm;
    return (int)(groups < cap ? (groups < 1 ? 1 : groups) : cap);
}
bool mis(const void *p) { return ((uintptr_t)p & 15) != 0; }
int fail(const char *what, cudaError_t e) {
    t2d_set_error("%s: %s", what, cudaGetErrorString(e));
    return T2D_E_CUDA;
}

}  // namespace

extern "C" int track2d_lstm_heads_forward(const float *gates_dev, const float *b_ih_dev, const float *b_hh_dev, const float *c_prev_dev, float *act_dev,
                                          float *c_next_dev, float *h_out_dev, float *h_next_dev, int64_t h_next_ld, const float *w_head_dev,
                                          const float *b_head_dev, float *out8_dev, int32_t *action_dev, const int32_t *forced_dev, float *value_dev,
                                          float *logp_dev, float *entropy_dev, float *logp_all_dev, const uint64_t *rng_step_dev, uint64_t seed,
                                          uint32_t rng_stream, int32_t greedy, int64_t E, int64_t first_row, void *stream) {
    if (!gates_dev || !b_ih_dev || !b_hh_dev || !c_prev_dev || !c_next_dev || !w_head_dev || !b_head_dev || !action_dev || E < 1 || first_row < 0 ||
        (!forced_dev && !greedy && !rng_step_dev)) {
        t2d_set_error("track2d_lstm_heads_forward: bad argument");
        return T2D_E_INVALID;
    }
    if (mis(gates_dev) || mis(b_ih_dev) || mis(b_hh_dev) || mis(c_prev_dev) || mis(act_dev) || mis(c_next_dev) || mis(h_out_dev) || mis(h_next_dev) ||
        mis(w_head_dev) || mis(out8_dev) || mis(logp_all_dev) || h_next_ld % 4) {
        t2d_set_error("track2d_lstm_heads_forward: buffers must be 16-byte aligned");
        return T2D_E_INVALID;
    }
    FwdArgs a;
    a.gates = gates_dev; a.b_ih = b_ih_dev; a.b_hh = b_hh_dev; a.c_prev = c_prev_dev; a.act = act_dev; a.c_next = c_next_dev; a.h_out = h_out_dev;
    a.h_next = h_next_dev; a.h_next_ld = h_next_ld; a.w_head = w_head_dev; a.b_head = b_head_dev; a.out8 = out8_dev; a.action = action_dev;
    a.forced = forced_dev; a.value = value_dev; a.logp = logp_dev; a.entropy = entropy_dev; a.logp_all = logp_all_dev;
    a.rng_step = reinterpret_cast<const unsigned long long *>(rng_step_dev); a.seed = seed; a.stream = rng_stream; a.greedy = greedy; a.E = E; a.first_row = first_row;
    lstm_heads_fwd_kernel<<<row_grid(E, WARPS, 8), WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("track2d_lstm_heads_forward", e);
    t2d_count_launches(1);
    return T2D_OK;
}

extern "C" int track2d_policy_post_step(const uint8_t *done_dev, float *h0_dev, float *h1_dev, int64_t h_ld, float *c0_dev, float *c1_dev,
                                        int32_t *eps_len_dev, uint64_t *rng_step_dev, int64_t E, void *stream) {
    if (E < 1 || h_ld % 4 || mis(h0_dev) || mis(h1_dev) || mis(c0_dev) || mis(c1_dev)) {
        t2d_set_error("track2d_policy_post_step: bad argument");
        return T2D_E_INVALID;
    }
    post_step_kernel<<<(unsigned)((E + 7) / 8), 256, 0, (cudaStream_t)stream>>>(done_dev, h0_dev, h1_dev, h_ld, c0_dev, c1_dev, eps_len_dev,
                                                                               reinterpret_cast<unsigned long long *>(rng_step_dev), E);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("track2d_policy_post_step", e);
    t2d_count_launches(1);
    return T2D_OK;
}

extern "C" int track2d_embed_add(const float *x_dev, int64_t ld, float *out_dev, int64_t out_ld, const float *w_dev, const float *b_dev,
                                 const int32_t *action_dev, int32_t N, int64_t E, void *stream) {
    if (!x_dev || !out_dev || !w_dev || !b_dev || !action_dev || N < 4 || N % 4 || ld % 4 || out_ld % 4 || E < 1 || mis(x_dev) || mis(out_dev) || mis(b_dev)) {
        t2d_set_error("track2d_embed_add: bad argument");
        return T2D_E_INVALID;
    }
    const long long total = E * (N / 4);
    embed_add_kernel<<<row_grid(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(x_dev, ld, out_dev, out_ld, w_dev, b_dev, action_dev, N, E);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("track2d_embed_add", e);
    t2d_count_launches(1);
    return T2D_OK;
}

extern "C" int track2d_a3c_loss_grad(const float *out8_0_dev, const float *out8_1_dev, float *dout8_0_dev, float *dout8_1_dev, const int32_t *actions_dev,
                                     const float *rewards_dev, const uint8_t *done_dev, float *stats_dev, float *returns_dev, float *gae_dev, int32_t T,
                                     int64_t E, double gamma, double tau, double w_ent0, double w_ent1, double scale, int32_t train0, int32_t train1,
                                     int32_t use_aux, void *stream) {
    if (!out8_0_dev || !out8_1_dev || !dout8_0_dev || !dout8_1_dev || !actions_dev || !rewards_dev || !done_dev || !stats_dev || T < 1 || E < 1 ||
        (!returns_dev) != (!gae_dev) || mis(out8_0_dev) || mis(out8_1_dev) || mis(dout8_0_dev) || mis(dout8_1_dev)) {
        t2d_set_error("track2d_a3c_loss_grad: bad argument");
        return T2D_E_INVALID;
    }
    LossArgs a;
    a.out8[0] = out8_0_dev; a.out8[1] = out8_1_dev; a.dout8[0] = dout8_0_dev; a.dout8[1] = dout8_1_dev; a.actions = actions_dev; a.rewards = rewards_dev;
    a.done = done_dev; a.stats = stats_dev; a.returns = returns_dev; a.gae = gae_dev; a.T = T; a.E = E; a.gamma = (float)gamma; a.tau = (float)tau;
    a.w_ent[0] = (float)w_ent0; a.w_ent[1] = (float)w_ent1; a.scale = (float)scale; a.train[0] = train0; a.train[1] = train1; a.use_aux = use_aux;
    a3c_loss_grad_kernel<<<(unsigned)((2 * E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("track2d_a3c_loss_grad", e);
    t2d_count_launches(1);
    return T2D_OK;
}

extern "C" int track2d_lstm_heads_backward(const float *dout8_dev, const float *w_head_dev, const float *dh_rec_dev, float *dc_dev, const uint8_t *done_dev,
                                           const float *act_dev, const float *c_prev_dev, float *dgates_dev, int64_t E, void *stream) {
    if (!dout8_dev || !w_head_dev || !dc_dev || !act_dev || !c_prev_dev || !dgates_dev || E < 1 || (dh_rec_dev && !done_dev) || mis(dout8_dev) ||
        mis(w_head_dev) || mis(dh_rec_dev) || mis(dc_dev) || mis(act_dev) || mis(c_prev_dev) || mis(dgates_dev)) {
        t2d_set_error("track2d_lstm_heads_backward: bad argument");
        return T2D_E_INVALID;
    }
    BwdArgs a;
    a.dout8 = dout8_dev; a.w_head = w_head_dev; a.dh_rec = dh_rec_dev; a.dc = dc_dev; a.done = done_dev; a.act = act_dev; a.c_prev = c_prev_dev;
    a.dgates = dgates_dev; a.E = E;
    lstm_heads_bwd_kernel<<<row_grid(E, WARPS, 8), WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("track2d_lstm_heads_backward", e);
    t2d_count_launches(1);
    return T2D_OK;
}

extern "C" int64_t track2d_relu_backward_workspace_floats(int64_t M, int32_t N) {
    if (M < 1 || N < 4 || N > 1024 || N % 4 || RG_THREADS % (N / 4)) return 0;
    return (int64_t)row_grid(M, RG_THREADS / (N / 4), 4) * 5 * N;
}

extern "C" int track2d_relu_backward_groupsum(float *dy_dev, const float *y_dev, int64_t y_ld, const int32_t *group_dev, int64_t M, int32_t N,
                                              float *dbias_dev, float *gw_dev, float *gb_dev, float *workspace_dev, int64_t workspace_floats, void *stream) {
    if (!dy_dev || !y_dev || y_ld % 4 || !dbias_dev || !workspace_dev || M < 1 || N < 4 || N > 1024 || N % 4 || RG_THREADS % (N / 4) || (group_dev && (!gw_dev || !gb_dev)) ||
        mis(dy_dev) || mis(y_dev) || mis(workspace_dev)) {
        t2d_set_error("track2d_relu_backward_groupsum: bad argument");
        return T2D_E_INVALID;
    }
    const int grid = row_grid(M, RG_THREADS / (N / 4), 4);
    if (workspace_floats < (int64_t)grid * 5 * N) {
        t2d_set_error("track2d_relu_backward_groupsum: workspace of %lld floats needed", (long long)grid * 5 * N);
        return T2D_E_INVALID;
    }
    relu_bwd_groupsum_kernel<<<grid, RG_THREADS, 0, (cudaStream_t)stream>>>(dy_dev, y_dev, y_ld, group_dev, M, N, workspace_dev);
    groupsum_reduce_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(workspace_dev, grid, N, group_dev ? 1 : 0, dbias_dev, gw_dev, gb_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("track2d_relu_backward_groupsum", e);
    t2d_count_launches(2);
    return T2D_OK;
}
