// track2d_step.cu -- the env-step kernel: Track1v1Env.step (envs/track_1v1.py:71-127) + gym TimeLimit
// for E envs in one launch.
//
// One CTA owns a chunk of N consecutive envs, so that its slice of the observation tensor
// ([E][2][13][13] fp32, 1352 B per env) is ONE contiguous, 16-byte aligned range of HBM:
//
//   A  (one thread per env)   load pos/ctr/actions, scripted-target override (Ram: device RNG; Nav: next
//                             plan action), move both agents with a wall-bit test (track_1v1.py:271-285),
//                             integer d^2 -> IEEE-double reward (track_1v1.py:94-104), C_far / TimeLimit
//                             done (track_1v1.py:106-111), store pos/ctr/reward/done, queue auto-reset.
//   B  (one thread per 4 window rows)  fetch 13 wall bits per row from the bit-packed, wall-framed map
//                             (coalesced within the env's 156-byte window, L2-resident), expand bits to
//                             bytes and stage 52 bytes = 13 aligned words in shared memory, conflict-free
//                             (lane stride 13 words).
//   C  (one thread per env)   stamp the agent cells: own centre 2 / 4, the other agent if inside the
//                             window (track_1v1.py:295-313).
//   D  (all threads)          stream the staged bytes out as float4 (or uint4 for the u8 variant):
//                             1 LDS.32 + 4 cvt + 1 STG.128 per 4 cells, fully coalesced.
//
// HBM traffic per env-step (algorithmic): 1,755 B, of which 1,352 B is the fp32 obs write.
#include "track2d_common.cuh"

#ifndef T2D_STEP_FORCE_N
#define T2D_STEP_FORCE_N 0 /* 16 or 32 pins the envs-per-CTA choice (tuning builds) */
#endif

namespace {

template <typename ObsT>
struct ObsStore;

template <>
struct ObsStore<float> {
    // words of staged bytes [0, nwords) -> floats; dst is 16-byte aligned.  UNROLL independent LDS are issued
    // before the converts and stores so one shared-memory latency covers UNROLL stores.  Stores are
    // streaming (st.global.cs): the observation tensor is written once per step and must not push the
    // bit-packed maps out of L2.
    template <int WORDS, int THREADS>
    static __device__ __forceinline__ void run_full(const uint32_t *sobs, float *dst, int tid) {
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        constexpr int ITERS = (WORDS + THREADS - 1) / THREADS;
        uint32_t w[ITERS];
#pragma unroll
        for (int i = 0; i < ITERS; i++) {
            int q = tid + i * THREADS;
            w[i] = (q < WORDS) ? sobs[q] : 0u;
        }
#pragma unroll
        for (int i = 0; i < ITERS; i++) {
            int q = tid + i * THREADS;
            if (q < WORDS) {
                float4 v;
                v.x = (float)(w[i] & 0xFFu);
                v.y = (float)((w[i] >> 8) & 0xFFu);
                v.z = (float)((w[i] >> 16) & 0xFFu);
                v.w = (float)(w[i] >> 24);
                __stcs(d4 + q, v);
            }
        }
    }
    static __device__ __forceinline__ void run(const uint32_t *sobs, float *dst, int nwords, int tid, int nthreads) {
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int q = tid; q < nwords; q += nthreads) {
            uint32_t w = sobs[q];
            float4 v;
            v.x = (float)(w & 0xFFu);
            v.y = (float)((w >> 8) & 0xFFu);
            v.z = (float)((w >> 16) & 0xFFu);
            v.w = (float)(w >> 24);
            __stcs(d4 + q, v);
        }
    }
    static __device__ __forceinline__ void tail(const uint32_t *sobs, float *dst, int first_word, int ncells_tail, int tid) {
        // fewer than 4 cells left (odd env count at the very end of the batch)
        if (tid < ncells_tail) dst[first_word * 4 + tid] = (float)((sobs[first_word] >> (8 * tid)) & 0xFFu);
    }
};

template <>
struct ObsStore<uint8_t> {
    template <int WORDS, int THREADS>
    static __device__ __forceinline__ void run_full(const uint32_t *sobs, uint8_t *dst, int tid) {
        // dst is 16-byte aligned when the chunk starts at a multiple of 8 envs (338 * 8 = 169 * 16)
        static_assert(WORDS % 4 == 0, "uint8 chunks are whole uint4s");
        constexpr int W4 = WORDS / 4;
        constexpr int ITERS = (W4 + THREADS - 1) / THREADS;
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(sobs);
        uint4 w[ITERS];
#pragma unroll
        for (int i = 0; i < ITERS; i++) {
            int q = tid + i * THREADS;
            w[i] = (q < W4) ? s4[q] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < ITERS; i++) {
            int q = tid + i * THREADS;
            if (q < W4) __stcs(d4 + q, w[i]);
        }
    }
    static __device__ __forceinline__ void run(const uint32_t *sobs, uint8_t *dst, int nwords, int tid, int nthreads) {
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(sobs);
        int n4 = nwords >> 2;
        for (int q = tid; q < n4; q += nthreads) __stcs(d4 + q, s4[q]);
        uint32_t *d1 = reinterpret_cast<uint32_t *>(dst);
        for (int q = (n4 << 2) + tid; q < nwords; q += nthreads) d1[q] = sobs[q];
    }
    static __device__ __forceinline__ void tail(const uint32_t *sobs, uint8_t *dst, int first_word, int ncells_tail, int tid) {
        if (tid < ncells_tail) dst[first_word * 4 + tid] = (uint8_t)((sobs[first_word] >> (8 * tid)) & 0xFFu);
    }
};

// TARGET: 0 = learned target (Adv / PZR / Far), 1 = Ram, 2 = Nav / RPF
// N: envs per CTA (even; multiple of 8 for the uint8 variant)
template <int TARGET, int RNG, typename ObsT, int N, int THREADS>
__global__ void __launch_bounds__(THREADS) step_kernel(World w, const int32_t *__restrict__ actions, ObsT *__restrict__ obs,
                                                       float *__restrict__ reward, uint8_t *__restrict__ done_out) {
    __shared__ __align__(16) uint32_t sobs[(T2D_ENV_CELLS * N) / 4];
    __shared__ uint32_t spos[N];

    // Programmatic dependent launch: when the previous kernel in the stream is another step (rollouts of scripted targets, the
    // roofline loop) this grid's CTAs are scheduled while that one drains; nothing is read before the wait, which returns once
    // the previous grid has completed and flushed.  After a kernel launched the ordinary way both instructions are no-ops.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int tid = threadIdx.x;
    const int env0 = blockIdx.x * N;
    const int nenv = min(N, w.E - env0);

    // ---- phase A: transition --------------------------------------------------------------------
    if (tid < N) {
        uint32_t p = 0;
        if (tid < nenv) {
            const int e = env0 + tid;
            p = w.pos[e];
            uint32_t ctr = w.ctr[e];
            int2 act = reinterpret_cast<const int2 *>(actions)[e];
            if (((unsigned)act.x | (unsigned)act.y) > 3u) {
                if ((unsigned)act.x > 3u || (TARGET == 0 && (unsigned)act.y > 3u)) atomicOr(w.status, (uint32_t)T2D_STATUS_BAD_ACTION);
                act.x &= 3;
                act.y &= 3;
            }
            const uint32_t *m = w.maps + (size_t)e * T2D_MAP_WORDS;
            int r0 = p & 255, c0 = (p >> 8) & 255, r1 = (p >> 16) & 255, c1 = p >> 24;

            if (TARGET == 1) { // RamAgent.step (navigator.py:77-88) replaces the target's action (track_1v1.py:81-82)
                uint32_t word = w.ram[e];
                if (RNG == T2D_RNG_NUMPY) {
                    MtRng rng;
                    rng.key = w.mt_key + (size_t)e * T2D_MT_N;
                    rng.pos = w.mt_pos[e];
                    act.y = (int)ram_step(rng, word);
                    w.mt_pos[e] = rng.pos;
                } else {
                    Philox rng;
                    rng.init(w.seed, (uint32_t)e, w.episode[e], 0x10000u + (ctr >> 16));
                    act.y = (int)ram_step(rng, word);
                }
                w.ram[e] = word;
            } else if (TARGET == 2) { // Navigator.step (navigator.py:11-36); exhausted plans were re-planned by nav_replan_kernel
                uint32_t meta = w.nav_meta[e];
                uint32_t idx = (meta >> 16) & (TRACK2D_NAV_MAXPLAN - 1);  // (the plan is a ring when it is fed ahead of time)
                act.y = (w.nav_plan[(size_t)e * T2D_NAV_PLAN_BYTES + (idx >> 2)] >> (2 * (idx & 3))) & 3;
                if (w.async_nav) atomicAdd(&w.nav_meta[e], 0x10000u);  // the planner's merge adds to the length field concurrently
                else w.nav_meta[e] = meta + 0x10000u;
            }

            // _next_state (track_1v1.py:271-285): stay put when the destination is a wall; agents may overlap
            int nr0 = r0 + action_dr(act.x), nc0 = c0 + action_dc(act.x);
            int nr1 = r1 + action_dr(act.y), nc1 = c1 + action_dc(act.y);
            if (!map_is_wall(m, nr0, nc0)) { r0 = nr0; c0 = nc0; }
            if (!map_is_wall(m, nr1, nc1)) { r1 = nr1; c1 = nc1; }
            p = (uint32_t)r0 | ((uint32_t)c0 << 8) | ((uint32_t)r1 << 16) | ((uint32_t)c1 << 24);

            int dr = r1 - r0, dc = c1 - c0;
            int d2 = dr * dr + dc * dc;
            double r_track, r_target;
            dueling_reward(d2, target_w_p(w.target_mode), r_track, r_target);

            uint32_t cfar = ctr & 0xFFFFu, elapsed = ctr >> 16;
            cfar = d2 <= 36 ? 0u : min(cfar + 1u, 0xFFFFu); // distance <= max_distance  <=>  d2 <= 36
            elapsed = min(elapsed + 1u, 0xFFFFu);
            bool done = cfar > 10u;                                                    // track_1v1.py:110-111
            if (w.max_steps > 0 && elapsed >= (uint32_t)w.max_steps) done = true;      // gym 0.12.5 TimeLimit

            w.pos[e] = p;
            w.ctr[e] = cfar | (elapsed << 16);
            reinterpret_cast<float2 *>(reward)[e] = make_float2((float)r_track, (float)r_target);
            done_out[e] = done ? 1 : 0;
            if (w.rew64) reinterpret_cast<double2 *>(w.rew64)[e] = make_double2(r_track, r_target);
            if (TARGET != 0) w.tgt_act[e] = (uint8_t)act.y;
            if (done && (w.flags & T2D_FLAG_AUTO_RESET)) {
                uint32_t slot = atomicAdd(&w.work_count[0], 1u);
                w.work_list[slot] = (uint32_t)e;
            }
        }
        spos[tid] = p;
    }
    __syncthreads();

    if (w.obs_type != T2D_OBS_PARTIAL || obs == nullptr) return; // Full observations are written by full_obs_kernel

    // ---- phase B: 13-bit window rows -> bytes, 4 rows (52 B = 13 words) per thread -----------------
    // All eight map-word loads of a group are issued before any is used (no branches in between: envs past
    // the end of the batch are clamped onto the last valid env and masked afterwards), so a group costs one
    // memory latency, not four.
    constexpr int GROUPS = (26 * N) / 4;
    constexpr int GITERS = (GROUPS + THREADS - 1) / THREADS;
#pragma unroll
    for (int git = 0; git < GITERS; git++) {
        const int g = tid + git * THREADS;
        if (g >= GROUPS) break;
        uint32_t lo_w[4], hi_w[4];
        int sh[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int R = 4 * g + i;
            int el = R / 26;
            int rem = R - 26 * el;
            int a = rem >= 13;
            int wr = rem - 13 * a;
            ok[i] = el < nenv;
            el = min(el, nenv - 1);
            uint32_t p = spos[el];
            int r = a ? (p >> 16) & 255 : p & 255;
            int c = a ? p >> 24 : (p >> 8) & 255;
            const uint32_t *row = w.maps + (size_t)(env0 + el) * T2D_MAP_WORDS + (r + wr) * T2D_ROW_WORDS;
            int wi = c >> 5;
            lo_w[i] = __ldg(row + wi);
            hi_w[i] = __ldg(row + (wi < 2 ? wi + 1 : 2));
            sh[i] = c & 31;
        }
        uint32_t mrow[4];
#pragma unroll
        for (int i = 0; i < 4; i++) mrow[i] = ok[i] ? (__funnelshift_r(lo_w[i], hi_w[i], sh[i]) & 0x1FFFu) : 0u;
        uint32_t lo = mrow[0] | (mrow[1] << 13) | (mrow[2] << 26);
        uint32_t hi = (mrow[2] >> 6) | (mrow[3] << 7);
        uint32_t *dst = sobs + 13 * g;
#pragma unroll
        for (int q = 0; q < 8; q++) dst[q] = spread4((lo >> (4 * q)) & 0xFu);
#pragma unroll
        for (int q = 0; q < 5; q++) dst[8 + q] = spread4((hi >> (4 * q)) & 0xFu);
    }
    __syncthreads();

    // ---- phase C: agent cells (track_1v1.py:295-313) -----------------------------------------------
    if (tid < nenv) {
        uint32_t p = spos[tid];
        int r0 = p & 255, c0 = (p >> 8) & 255, r1 = (p >> 16) & 255, c1 = p >> 24;
        int dr = r1 - r0, dc = c1 - c0;
        uint8_t *sb = reinterpret_cast<uint8_t *>(sobs) + tid * T2D_ENV_CELLS;
        if (abs(dr) <= T2D_PAD && abs(dc) <= T2D_PAD) {
            sb[(T2D_PAD + dr) * T2D_WIN + T2D_PAD + dc] = 4;                  // the target inside the tracker's window
            sb[T2D_WIN_CELLS + (T2D_PAD - dr) * T2D_WIN + T2D_PAD - dc] = 2;  // the tracker inside the target's window
        }
        sb[84] = 2;                  // each agent re-stamps its own colour at its own centre
        sb[T2D_WIN_CELLS + 84] = 4;
    }
    __syncthreads();

    // ---- phase D: coalesced 16-byte stores ---------------------------------------------------------
    ObsT *dst = obs + (size_t)env0 * T2D_ENV_CELLS;
    if (nenv == N) {
        ObsStore<ObsT>::template run_full<(T2D_ENV_CELLS * N) / 4, THREADS>(sobs, dst, tid);
    } else {
        const int ncells = nenv * T2D_ENV_CELLS;
        ObsStore<ObsT>::run(sobs, dst, ncells >> 2, tid, THREADS);
        if (ncells & 3) ObsStore<ObsT>::tail(sobs, dst, ncells >> 2, ncells & 3, tid);
    }
}

// Full observations (obs_type 'Full', track_1v1.py:288-290): both agents get the whole map with the
// tracker cell = 2 and then the target cell = 4.  One CTA per env; plain, not a headline path.
template <typename ObsT>
__global__ void full_obs_kernel(World w, ObsT *__restrict__ obs, const uint8_t *__restrict__ mask) {
    const int e = blockIdx.x;
    if (mask && !mask[e]) return;
    const uint32_t *m = w.maps + (size_t)e * T2D_MAP_WORDS;
    uint32_t p = w.pos[e];
    int r0 = p & 255, c0 = (p >> 8) & 255, r1 = (p >> 16) & 255, c1 = p >> 24;
    const int cells = w.H * w.W;
    ObsT *o = obs + (size_t)e * 2 * cells;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        int r = i / w.W, c = i - r * w.W;
        int v = map_is_wall(m, r, c);
        if (r == r0 && c == c0) v = 2;
        if (r == r1 && c == c1) v = 4;
        o[i] = (ObsT)v;
        o[cells + i] = (ObsT)v;
    }
}

// launch with the programmatic-stream-serialization attribute (see the head of step_kernel)
template <typename Kernel, typename... Args>
cudaError_t launch_pdl(Kernel kernel, int grid, int threads, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = getenv("T2D_STEP_NO_PDL") == nullptr;  // A/B switch for profiles/
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int TARGET, int RNG, typename ObsT>
cudaError_t launch_step_t(const World &w, const int32_t *actions, ObsT *obs, float *reward, uint8_t *done, cudaStream_t s) {
    // 16 envs per CTA of 128 threads: 104 four-row groups -> at most one group per thread (a single load latency in
    // phase B), 5.4 KB of staging, 16 CTAs (2048 threads) resident per SM = 37,888 envs per wave on 148 SMs.
    // A 32-envs-per-CTA variant (65,536 envs in ONE wave instead of 1.73) was 1 % faster without programmatic dependent
    // launch and is 3.5 % slower with it (20.84 vs 20.11 us: the next step's CTAs now fill the tail of this one), so it is
    // only kept behind T2D_STEP_FORCE_N=32.  Large batches: 96 % of the HBM roofline at 1 M envs.
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const long wave16 = (long)sms * 16 * 16, wave32 = (long)sms * 16 * 32;
    static const int force = getenv("T2D_STEP_FORCE_N") ? atoi(getenv("T2D_STEP_FORCE_N")) : T2D_STEP_FORCE_N;  // tuning switch
    (void)wave16; (void)wave32;
    if (force == 32) {
        constexpr int N = 32, T = 128;
        return launch_pdl(step_kernel<TARGET, RNG, ObsT, N, T>, (w.E + N - 1) / N, T, s, w, actions, obs, reward, done);
    } else {
        constexpr int N = 16, T = 128;
        return launch_pdl(step_kernel<TARGET, RNG, ObsT, N, T>, (w.E + N - 1) / N, T, s, w, actions, obs, reward, done);
    }
}

template <typename ObsT>
cudaError_t launch_step_obs(const World &w, const int32_t *actions, ObsT *obs, float *reward, uint8_t *done, cudaStream_t s) {
    const bool numpy = w.rng_mode == T2D_RNG_NUMPY;
    switch (w.target_mode) {
        case T2D_TARGET_RAM:
            return numpy ? launch_step_t<1, T2D_RNG_NUMPY, ObsT>(w, actions, obs, reward, done, s)
                         : launch_step_t<1, T2D_RNG_PHILOX, ObsT>(w, actions, obs, reward, done, s);
        case T2D_TARGET_NAV:
        case T2D_TARGET_RPF:
            return launch_step_t<2, T2D_RNG_PHILOX, ObsT>(w, actions, obs, reward, done, s);
        default:
            // (a single-barrier variant in which every row thread recomputes the moves itself was measured at 22.3 us vs
            // 20.8 us for this kernel at 65,536 envs -- the redundant loads cost more than the two barriers)
            return launch_step_t<0, T2D_RNG_PHILOX, ObsT>(w, actions, obs, reward, done, s);
    }
}

} // namespace

cudaError_t t2d_launch_step_f32(const World &w, const int32_t *actions, float *obs, float *reward, uint8_t *done, cudaStream_t s) {
    return launch_step_obs<float>(w, actions, obs, reward, done, s);
}
cudaError_t t2d_launch_step_u8(const World &w, const int32_t *actions, uint8_t *obs, float *reward, uint8_t *done, cudaStream_t s) {
    return launch_step_obs<uint8_t>(w, actions, obs, reward, done, s);
}
cudaError_t t2d_launch_full_obs_f32(const World &w, float *obs, const uint8_t *mask, cudaStream_t s) {
    full_obs_kernel<float><<<w.E, 256, 0, s>>>(w, obs, mask);
    return cudaGetLastError();
}
cudaError_t t2d_launch_full_obs_u8(const World &w, uint8_t *obs, const uint8_t *mask, cudaStream_t s) {
    full_obs_kernel<uint8_t><<<w.E, 256, 0, s>>>(w, obs, mask);
    return cudaGetLastError();
}
