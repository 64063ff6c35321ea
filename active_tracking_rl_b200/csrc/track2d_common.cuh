// track2d_common.cuh -- world-state layout and device helpers shared by every kernel of libtrack2d.
//
// World state is a struct of arrays over E envs, resident in HBM:
//   maps   u32 [E][96][3]   bit-packed wall grid, 1 = wall.  The map is stored with a 6-cell frame of
//                           walls round it (T2D_PAD): map cell (r, c) lives at padded (r + 6, c + 6), so
//                           the 13x13 field of view of an agent at (r, c) is padded rows r..r+12, bits
//                           c..c+12 -- no bounds tests, and everything outside the map reads 1 exactly as
//                           np.pad(..., constant_values=1) does in the reference (track_1v1.py:320-321).
//                           96 rows x 12 B = 1152 B per env (9 x 128 B lines).
//   pos    u32 [E]          tracker row | tracker col << 8 | target row << 16 | target col << 24
//   ctr    u32 [E]          C_far | elapsed_steps << 16
//   ram    u32 [E]          RamAgent: 9 x 2-bit plan | len << 18 | idx << 22
//   ...                     (see struct World)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/track2d.h"

#define T2D_PAD 6
#define T2D_MAP_ROWS 96
#define T2D_ROW_WORDS 3
#define T2D_MAP_WORDS (T2D_MAP_ROWS * T2D_ROW_WORDS) /* 288 */
#define T2D_WIN 13
#define T2D_WIN_CELLS 169
#define T2D_ENV_CELLS 338 /* two agents' windows */
#define T2D_NAV_PLAN_BYTES (TRACK2D_NAV_MAXPLAN / 4)
#define T2D_MT_N 624

struct World {
    int E, H, W;
    int map_type, obs_type, target_mode, level, rng_mode, max_steps, flags;
    unsigned long long seed;
    uint32_t *maps;      // [E][288]
    uint32_t *pos;       // [E]
    uint32_t *ctr;       // [E]
    uint32_t *goals;     // [E] goal0 row | col << 8 | goal1 row << 16 | col << 24
    uint32_t *ram;       // [E]
    uint32_t *episode;   // [E] episodes started so far (Philox counter word)
    uint8_t *nav_plan;   // [E][T2D_NAV_PLAN_BYTES], 2 bits per action
    uint32_t *nav_meta;  // [E] len | idx << 16
    uint32_t *nav_goal;  // [E] row | col << 8
    uint32_t *rpf;       // [E] static-goal cursor | (4-bit "corner was a wall in the env map") << 8
    uint32_t *mt_key;    // [E][624]  (T2D_RNG_NUMPY only)
    int32_t *mt_pos;     // [E]
    double *rew64;       // [E][2] or NULL
    uint8_t *tgt_act;    // [E]
    uint32_t *work_list; // [E] env indices queued for reset / replan
    uint32_t *work_count;// [4]: [0] resets queued, [1] replans queued, [2] reset-kernel CTA ticket
    uint32_t *status;    // [1] OR of T2D_STATUS_*
    unsigned long long *stats; // [2] episodes finished, env-steps done
    uint8_t *astar_ws;   // A* workspace (Nav/RPF only)
    int astar_slots;
    // ---- asynchronous planning (Philox + Nav + auto-reset; see track2d_reset.cu "plan-ahead") -------------------------------
    int async_nav;       // 1: nav_plan is an append-only ring (len and idx of nav_meta only grow), fed ahead of time
    uint32_t *nav_end;   // [E] cell the target stands on after the last planned action (row | col << 8)
    uint4 *ext_info;     // [E] hand-over slot of the side-stream planner: x = state (0 idle, 2 planning, 1 ready) | episode << 8,
                         //     y = base | length << 16, z = goal, w = end cell
    uint8_t *ext_plan;   // [E][T2D_NAV_PLAN_BYTES] the planned segment, slots 0 .. length-1
};

// ---- map bit helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int map_word_index(int r, int c) { return (r + T2D_PAD) * T2D_ROW_WORDS + ((c + T2D_PAD) >> 5); }
__device__ __forceinline__ uint32_t map_bit_of(int c) { return 1u << ((c + T2D_PAD) & 31); }
__device__ __forceinline__ int map_is_wall(const uint32_t *__restrict__ m, int r, int c) {
    return (__ldg(m + map_word_index(r, c)) >> ((c + T2D_PAD) & 31)) & 1u;
}
// 13 wall bits of padded row `pr`, starting at padded column `pc` (0..81)
__device__ __forceinline__ uint32_t map_row13(const uint32_t *__restrict__ m, int pr, int pc) {
    const uint32_t *row = m + pr * T2D_ROW_WORDS;
    int wi = pc >> 5;
    uint32_t lo = row[wi];
    uint32_t hi = row[wi < 2 ? wi + 1 : 2];
    return __funnelshift_r(lo, hi, pc & 31) & 0x1FFFu;
}
// 4 bits -> 4 bytes (bit i -> byte i): one IMAD + one LOP
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ int action_dr(int a) { return a == 0 ? -1 : (a == 1 ? 1 : 0); }  // track_1v1.py:276
__device__ __forceinline__ int action_dc(int a) { return a == 2 ? -1 : (a == 3 ? 1 : 0); }

// ---- Philox4x32-10 (Salmon et al. 2011), written out so the host tests can restate it ------------
struct Philox {
    uint32_t k0, k1;
    uint32_t c0, c1, c2; // c3 = block counter
    uint32_t blk;
    uint32_t out[4];
    int have;
    __device__ __forceinline__ void init(unsigned long long seed, uint32_t env, uint32_t episode, uint32_t stream) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        c0 = env; c1 = episode; c2 = stream; blk = 0; have = 0;
    }
    __device__ __forceinline__ void block() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = blk++;
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
        have = 4;
    }
    __device__ __forceinline__ uint32_t u32() {
        if (have == 0) block();
        // select without dynamic indexing (keeps out[] in registers)
        uint32_t v = have == 4 ? out[0] : (have == 3 ? out[1] : (have == 2 ? out[2] : out[3]));
        have--;
        return v;
    }
    // uniform integer in [0, n): 64-bit multiply-shift, bias < n / 2^64
    __device__ __forceinline__ uint32_t below(uint32_t n) {
        unsigned long long x = ((unsigned long long)u32() << 32) | u32();
        return (uint32_t)__umul64hi(x, (unsigned long long)n);
    }
    // same construction as numpy's legacy random_sample so int(r * 6400) has the reference's law
    __device__ __forceinline__ double dbl() {
        uint32_t a = u32() >> 5, b = u32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    __device__ __forceinline__ uint32_t interval(uint32_t max) { return max == 0 ? 0u : below(max + 1u); }
};

// ---- numpy legacy RandomState: MT19937 state in (shared or global) memory, driven by ONE thread ---
// Restates numpy/random/src/mt19937 + legacy-distributions exactly as the reference's numpy==1.14
// pin consumes them: genrand_int32, random_sample, rk_interval masked rejection.
struct MtRng {
    uint32_t *key; // 624 words
    int pos;
    __device__ __forceinline__ void regenerate() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
        int i;
        uint32_t y;
        for (i = 0; i < 624 - 397; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
        }
        for (; i < 623; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
        }
        y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
        pos = 0;
    }
    __device__ __forceinline__ uint32_t u32() {
        if (pos >= 624) regenerate();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    __device__ __forceinline__ double dbl() {
        uint32_t a = u32() >> 5, b = u32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    __device__ __forceinline__ uint32_t interval(uint32_t max) {
        if (max == 0) return 0;
        uint32_t mask = max, v;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        while ((v = (u32() & mask)) > max) {
        }
        return v;
    }
    __device__ __forceinline__ uint32_t below(uint32_t n) { return interval(n - 1u); } // randint(0, n)
};

// ---- RamAgent state word --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ram_pack(uint32_t plan_bits, uint32_t len, uint32_t idx) { return (plan_bits & 0x3FFFFu) | (len << 18) | (idx << 22); }

// navigator.py:90-93 RamAgent.reset and :77-88 step, templated on the RNG (draw order: randint first)
template <typename Rng>
__device__ __forceinline__ uint32_t ram_new_plan(Rng &rng, bool constant, uint32_t action) {
    uint32_t len = 1u + rng.interval(8u); // np.random.randint(1, 10)
    uint32_t bits = 0;
    for (uint32_t i = 0; i < len; i++) bits |= (constant ? action : rng.interval(3u)) << (2 * i);
    return ram_pack(bits, len, 0);
}
template <typename Rng>
__device__ __forceinline__ uint32_t ram_step(Rng &rng, uint32_t &word) {
    uint32_t len = (word >> 18) & 15u, idx = (word >> 22) & 15u;
    uint32_t action = (word >> (2 * idx)) & 3u;
    idx++;
    if (idx >= len) {
        if (rng.interval(1u) == 0) {          // np.random.choice([0, 1], 1) == 0
            action = rng.interval(3u);        // np.random.choice(all_actions, 1): returned as this step's action
            word = ram_new_plan(rng, true, action);
        } else {
            word = ram_new_plan(rng, false, 0);
        }
    } else {
        word = (word & ~(15u << 22)) | (idx << 22);
    }
    return action;
}

// Reward arithmetic of track_1v1.py:94-111 in IEEE double, op for op (intrinsics so nothing is
// contracted or re-associated).  d2 = squared integer distance.
__device__ __forceinline__ void dueling_reward(int d2, double w_p, double &r_track, double &r_target) {
    const double maxd = 6.0;
    double dist = __dsqrt_rn((double)d2);                                  // np.linalg.norm
    r_track = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(2.0, dist), maxd));       // 1 - 2*distance/max_distance
    r_track = r_track > -1.0 ? r_track : -1.0;                             // max(r_track, -1)
    double over = __dsub_rn(dist, maxd);
    over = over > 0.0 ? over : 0.0;                                        // max(distance - max_distance, 0)
    r_target = __dsub_rn(-r_track, __ddiv_rn(__dmul_rn(w_p, over), maxd)); // -r_track - w_p*over/max_distance
    r_target = r_target > -1.0 ? r_target : -1.0;
}

__host__ __device__ __forceinline__ double target_w_p(int target_mode) { // track_1v1.py:147-152
    return target_mode == T2D_TARGET_PZR ? 1.0 : (target_mode == T2D_TARGET_FAR ? -0.5 : 0.0);
}
