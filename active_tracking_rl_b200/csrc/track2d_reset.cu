// track2d_reset.cu -- Track1v1Env.reset (envs/track_1v1.py:134-168): a fresh random map per episode
// (init_maze, :218-240), spawn sampling (generators.py:21-94), scripted-target reset, counters, and the
// first observation.  One WARP per env being reset; the map is built as a bit grid in shared memory
// and written to HBM with coalesced 4-byte stores.
//
// Two RNG back-ends (track2d.h):
//   T2D_RNG_PHILOX  warp-parallel sampling with the reference's distributions:
//                   k = int(0.15*u*6400) distinct obstacles by parallel draw-until-new with warp match /
//                   ballot de-duplication (== sequential sampling without replacement), spawn = j-th free
//                   cell by popcount prefix scan, target = uniform free cell of the 2x2 block.
//   T2D_RNG_NUMPY   lane 0 replays numpy's legacy RandomState calls in the reference's exact order
//                   (random_sample, permutation(6400)[:k], permutation(n_free)[:2], ...), so env e equals
//                   the reference after np.random.seed(seed + e) bit for bit.  Slow by construction (four
//                   sequential Fisher-Yates shuffles of ~6400 per reset); it is the parity mode.
#include "track2d_common.cuh"
#include "track2d_nav.cuh"

namespace {

// ---- shared helpers -------------------------------------------------------------------------------

// bit grid with the 6-cell wall frame; interior (map rows/cols 1..H-2) cleared
__device__ void bm_init(uint32_t *bm, int H, int W, int lane) {
    const int lastc = W - 2 + T2D_PAD; // last interior padded column
    for (int i = lane; i < T2D_MAP_WORDS; i += 32) {
        int pr = i / 3, wi = i - 3 * pr;
        uint32_t v = 0xFFFFFFFFu;
        if (pr >= T2D_PAD + 1 && pr <= H - 2 + T2D_PAD) {
            int lo = max(T2D_PAD + 1, 32 * wi), hi = min(lastc, 32 * wi + 31); // zero [lo, hi]
            if (lo <= hi) {
                uint32_t span = hi - lo + 1;
                uint32_t zmask = (span == 32 ? 0xFFFFFFFFu : ((1u << span) - 1u)) << (lo - 32 * wi);
                v = ~zmask;
            }
        }
        bm[i] = v;
    }
    __syncwarp();
}

__device__ __forceinline__ int bm_wall(const uint32_t *bm, int r, int c) { return (bm[map_word_index(r, c)] >> ((c + T2D_PAD) & 31)) & 1u; }
__device__ __forceinline__ void bm_set(uint32_t *bm, int r, int c) { atomicOr(&bm[map_word_index(r, c)], map_bit_of(c)); }
__device__ __forceinline__ void bm_clear(uint32_t *bm, int r, int c) { atomicAnd(&bm[map_word_index(r, c)], ~map_bit_of(c)); }

// generators.py:115-145 RandomMazeGenerator._generate_maze, sequential (one thread), any RNG
template <typename Rng>
__device__ void maze_walk(uint32_t *bm, Rng &rng, double r) {
    const int sh = 81;
    int complexity = (int)(r * (double)(5 * (sh + sh)));
    int density = (int)(r * (double)((sh / 2) * (sh / 2)));
    for (int i = 0; i < density; i++) {
        int x = (int)rng.interval(sh / 2) * 2; // randint(0, 41) * 2, x before y
        int y = (int)rng.interval(sh / 2) * 2;
        bm[map_word_index(y, x)] |= map_bit_of(x);
        for (int j = 0; j < complexity; j++) {
            int ny[4], nx[4], nn = 0;
            if (x > 1) { ny[nn] = y; nx[nn] = x - 2; nn++; }
            if (x < sh - 2) { ny[nn] = y; nx[nn] = x + 2; nn++; }
            if (y > 1) { ny[nn] = y - 2; nx[nn] = x; nn++; }
            if (y < sh - 2) { ny[nn] = y + 2; nx[nn] = x; nn++; }
            int pick = (int)rng.interval((uint32_t)(nn - 1));
            int y_ = ny[0], x_ = nx[0];
#pragma unroll
            for (int q = 1; q < 4; q++)
                if (pick == q) { y_ = ny[q]; x_ = nx[q]; }
            if (!bm_wall(bm, y_, x_)) {
                bm[map_word_index(y_, x_)] |= map_bit_of(x_);
                int my = y_ + (y - y_) / 2, mx = x_ + (x - x_) / 2;
                bm[map_word_index(my, mx)] |= map_bit_of(mx);
                x = x_;
                y = y_;
            }
        }
    }
}

// Philox words generated ahead by the whole warp (see reset_env_philox), consumed by the sequential maze walk: keeps the
// ten-round Philox out of lane 0's critical path.  interval(max) = floor(u32 * (max + 1) / 2^32) (bias <= 41 / 2^32).
#define T2D_WALK_WORDS 1232 /* >= 47 seeds x (2 + 24 steps) */
struct WordStream {
    const uint32_t *w;
    int i;
    __device__ __forceinline__ uint32_t interval(uint32_t max) { return __umulhi(w[i++], max + 1u); }
};

// number of free cells, and the j-th free cell in row-major order (np.where(maze == 0) order).
// Each lane owns 9 consecutive words of the grid.
__device__ int bm_count_free(const uint32_t *bm, int lane) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) n += __popc(~bm[9 * lane + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
    return n;
}
__device__ uint32_t bm_select_free(const uint32_t *bm, int j, int lane) { // returns row | col << 8
    int mine = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) mine += __popc(~bm[9 * lane + i]);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    int excl = incl - mine;
    uint32_t found = 0;
    if (j >= excl && j < incl) {
        int rem = j - excl;
        for (int i = 0; i < 9; i++) {
            uint32_t z = ~bm[9 * lane + i];
            int pc = __popc(z);
            if (rem < pc) {
                int bit = __fns(z, 0, rem + 1);
                int word = 9 * lane + i;
                int pr = word / 3, pcol = 32 * (word - 3 * pr) + bit;
                found = (uint32_t)(pr - T2D_PAD) | ((uint32_t)(pcol - T2D_PAD) << 8);
                break;
            }
            rem -= pc;
        }
    }
    uint32_t owner = __ballot_sync(0xFFFFFFFFu, j >= excl && j < incl);
    return __shfl_sync(0xFFFFFFFFu, found, __ffs(owner) - 1);
}

// generators.py:82-94 get_around(state, 1): free cells of {r-1, r} x {c-1, c} in row-major order
__device__ __forceinline__ int around_free(const uint32_t *bm, int r, int c, int cr[4], int cc[4]) {
    int n = 0;
    for (int rr = max(0, r - 1); rr < r + 1; rr++)
        for (int q = max(0, c - 1); q < c + 1; q++)
            if (!bm_wall(bm, rr, q)) { cr[n] = rr; cc[n] = q; n++; }
    return n;
}

template <typename ObsT>
__device__ void write_partial_obs(const uint32_t *bm, uint32_t p, ObsT *__restrict__ o, int lane) {
    int r0 = p & 255, c0 = (p >> 8) & 255, r1 = (p >> 16) & 255, c1 = p >> 24;
    for (int i = lane; i < T2D_ENV_CELLS; i += 32) {
        int a = i >= T2D_WIN_CELLS;
        int cell = i - T2D_WIN_CELLS * a;
        int wr = cell / T2D_WIN, wc = cell - T2D_WIN * wr;
        int r = (a ? r1 : r0) - T2D_PAD + wr, c = (a ? c1 : c0) - T2D_PAD + wc;
        int v = (bm[(r + T2D_PAD) * T2D_ROW_WORDS + ((c + T2D_PAD) >> 5)] >> ((c + T2D_PAD) & 31)) & 1u;
        if (r == r0 && c == c0) v = 2;
        if (r == r1 && c == c1) v = 4;
        if (cell == 84) v = 2 + 2 * a;
        o[i] = (ObsT)v;
    }
}

// after a list-driven auto-reset the LAST CTA to finish accounts the episodes and clears the queue
// (every CTA has read work_count[0] by then), so no extra launch is needed
__device__ __forceinline__ void finish_reset(const World &w) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        uint32_t t = atomicAdd(&w.work_count[2], 1u);
        if (t == gridDim.x - 1) {
            w.stats[0] += w.work_count[0];
            w.work_count[0] = 0;
            w.work_count[2] = 0;
            __threadfence();
        }
    }
}


// ---- numpy legacy sampling helpers (T2D_RNG_NUMPY) ------------------------------------------------------
// perm / freel (numpy sampling) and the A* maps are never live at the same time: one union.
struct NumpyScratch {
    uint32_t key[T2D_MT_N];
    uint32_t bm[T2D_MAP_WORDS];
    union {
        struct {
            uint16_t perm[82 * 82];
            uint16_t freel[82 * 82];
        };
        AStarScratch astar;
    };
};
struct PhiloxNavScratch {
    uint32_t bm[T2D_MAP_WORDS];
    AStarScratch astar;
};

// np.random.permutation(n) into s.perm: arange by the warp, legacy shuffle by lane 0
__device__ void np_permutation(NumpyScratch &s, MtRng &rng, int n, int lane) {
    for (int i = lane; i < n; i += 32) s.perm[i] = (uint16_t)i;
    __syncwarp();
    if (lane == 0) {
        for (int i = n - 1; i >= 1; i--) {
            int j = (int)rng.interval((uint32_t)i);
            uint16_t t = s.perm[i];
            s.perm[i] = s.perm[j];
            s.perm[j] = t;
        }
    }
    __syncwarp();
}

// np.where(maze == 0) row-major over the H x W map -> s.freel (cell = r * W + c); returns the count
__device__ int np_free_list(NumpyScratch &s, int H, int W, int lane) {
    int base = 0;
    const int cells = H * W;
    for (int i0 = 0; i0 < cells; i0 += 32) {
        int i = i0 + lane;
        bool fr = false;
        if (i < cells) {
            int r = i / W, c = i - r * W;
            fr = !bm_wall(s.bm, r, c);
        }
        uint32_t msk = __ballot_sync(0xFFFFFFFFu, fr);
        if (fr) s.freel[base + __popc(msk & ((1u << lane) - 1u))] = (uint16_t)i;
        base += __popc(msk);
    }
    __syncwarp();
    return base;
}

// RPF: advance the static-goal cursor (generators.py:47-50) and return that corner
__device__ __forceinline__ uint32_t rpf_next_goal(const World &w, int e, int lane) {
    uint32_t word = w.rpf[e];
    uint32_t v = ((word & 255u) + 1u) & 3u;
    __syncwarp();
    if (lane == 0) w.rpf[e] = (word & ~255u) | v;
    __syncwarp();
    return rpf_corner(w.H, w.W, (int)v);
}

// generators.py:38-51 sample_goal(num <= 2) on the generator maze (non-static) or the RPF corner cursor
__device__ void np_sample_goal(const World &w, NumpyScratch &s, MtRng &rng, int e, int num, uint32_t g[2], int lane) {
    if (w.target_mode != T2D_TARGET_RPF) {
        int n = np_free_list(s, w.H, w.W, lane);
        np_permutation(s, rng, n, lane);
        for (int i = 0; i < num; i++) {
            int cell = s.freel[s.perm[i]];
            g[i] = (uint32_t)(cell / w.W) | ((uint32_t)(cell % w.W) << 8);
        }
        __syncwarp();
    } else {
        uint32_t corner = rpf_next_goal(w, e, lane);
        for (int i = 0; i < num; i++) g[i] = corner;
    }
}

// static_goals (generators.py:12-19) on a generator-maze copy: remember which corners were walls in the
// ENV map (Track1v1Env.maze is copied before static_goals runs, track_1v1.py:234-236), clear them.
__device__ __forceinline__ uint32_t rpf_clear_corners(uint32_t *bm, int H, int W, int lane) {
    uint32_t wallbits = 0;
    for (int q = 0; q < 4; q++) {
        uint32_t cn = rpf_corner(H, W, q);
        wallbits |= (uint32_t)bm_wall(bm, cn & 255, cn >> 8) << q;
    }
    __syncwarp();
    if (lane == 0)
        for (int q = 0; q < 4; q++) {
            uint32_t cn = rpf_corner(H, W, q);
            bm[map_word_index(cn & 255, cn >> 8)] &= ~map_bit_of(cn >> 8);
        }
    __syncwarp();
    return wallbits;
}

// ---- Navigator glue (envs/navigator.py) -------------------------------------------------------------
struct NumpyNavPolicy {
    const World &w; int e; NumpyScratch &s; MtRng &rng;
    __device__ const uint32_t *bm() const { return s.bm; }
    __device__ AStarScratch &astar() { return s.astar; }
    __device__ uint32_t sample_goal1(int lane) { // maze_generator.sample_goal(1)[0]
        uint32_t g[2];
        np_sample_goal(w, s, rng, e, 1, g, lane);
        return g[0];
    }
    __device__ void plan_b(uint32_t acts[10], int lane) { // np.random.choice(all_actions, 10)
        if (lane == 0)
            for (int i = 0; i < 10; i++) acts[i] = rng.interval(3u);
    }
};
struct PhiloxNavPolicy {
    const World &w; int e; PhiloxNavScratch &s; Philox &rng;
    __device__ const uint32_t *bm() const { return s.bm; }
    __device__ AStarScratch &astar() { return s.astar; }
    __device__ uint32_t sample_goal1(int lane) {
        if (w.target_mode == T2D_TARGET_RPF) return rpf_next_goal(w, e, lane);
        int nfree = bm_count_free(s.bm, lane);
        return bm_select_free(s.bm, (int)rng.below((uint32_t)nfree), lane);
    }
    __device__ void plan_b(uint32_t acts[10], int) {
        for (int i = 0; i < 10; i++) acts[i] = rng.interval(3u);
    }
};

// Shared tail of Navigator.reset (navigator.py:38-63) and of the replan inside Navigator.step (:15-36):
// A* to `goal`; while unsolvable or empty: up to 5 fresh goals; then plan B = 10 random actions.
template <typename Policy>
__device__ int nav_plan_into(const World &w, uint8_t *plan, int base, Policy &P, int slot, int sr, int sc, uint32_t &goal, int lane) {
    int count_res = 0;
    bool planb = false;
    int len = astar_plan(w, plan, base, P.bm(), P.astar(), slot, sr, sc, (int)(goal & 255u), (int)((goal >> 8) & 255u), lane);
    while (len < 1) {
        count_res++;
        if (count_res > 5) { planb = true; break; }
        goal = P.sample_goal1(lane);
        len = astar_plan(w, plan, base, P.bm(), P.astar(), slot, sr, sc, (int)(goal & 255u), (int)((goal >> 8) & 255u), lane);
    }
    if (planb) {
        uint32_t acts[10];
        P.plan_b(acts, lane);
        if (lane == 0) nav_store_planb(plan, base, acts);
        len = 10;
    }
    __syncwarp();
    return len;
}
// the reference's semantics: the new plan REPLACES the old one (a_i = 0)
template <typename Policy>
__device__ void nav_plan_from(const World &w, int e, Policy &P, int slot, int sr, int sc, uint32_t goal, int lane) {
    const int len = nav_plan_into(w, w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES, -1, P, slot, sr, sc, goal, lane);
    if (lane == 0) {
        w.nav_meta[e] = (uint32_t)len; // a_i = 0
        w.nav_goal[e] = goal & 0xFFFFu;
    }
    __syncwarp();
}

// the cell an agent ends on after `len` planned actions from (r, c) under the env's move rule (track_1v1.py:271-285: a wall blocks)
__device__ uint32_t plan_end_cell(const uint32_t *bm, const uint8_t *plan, int base, int len, int r, int c) {
    for (int i = 0; i < len; i++) {
        const int a = plan_get(plan, base + i);
        const int nr = r + action_dr(a), nc = c + action_dc(a);
        if (!((bm[map_word_index(nr, nc)] >> ((nc + T2D_PAD) & 31)) & 1u)) { r = nr; c = nc; }
    }
    return (uint32_t)r | ((uint32_t)c << 8);
}

// ---- Philox reset ----------------------------------------------------------------------------------
// `nav` is non-NULL only for the Nav / RPF instantiation (it then also provides bm).
template <typename ObsT, int MAP>
__device__ __noinline__ void reset_env_philox(const World &w, int e, uint32_t *bm, uint32_t *walkw, PhiloxNavScratch *nav, int slot, ObsT *obs, int lane,
                                              bool init_only, bool defer_plan = false) {
    const uint32_t episode = w.episode[e];
    Philox rng; // stream 0: the same on every lane (scalar decisions need no shuffles)
    rng.init(w.seed, (uint32_t)e, episode, 0u);
    bm_init(bm, w.H, w.W, lane);

    if (MAP == T2D_MAP_MAZE) {
        double r = w.level > 0 ? w.level * 0.02 : .03 * rng.dbl();
        {   // every lane fills its share of the word stream (block b -> words 4b..4b+3), then lane 0 walks
            const int density = (int)(r * 1600.0), complexity = (int)(r * 810.0);
            const int need = density * (2 + complexity);
            Philox walk;
            walk.init(w.seed, (uint32_t)e, episode, 2u);
            for (int b = lane; 4 * b < need; b += 32) {
                walk.blk = (uint32_t)b;
                walk.block();
#pragma unroll
                for (int q = 0; q < 4; q++) walkw[4 * b + q] = walk.out[q];
            }
            __syncwarp();
            if (lane == 0) {
                WordStream ws{walkw, 0};
                maze_walk(bm, ws, r);
            }
        }
        __syncwarp();
    } else {
        double r = MAP == T2D_MAP_EMPTY ? 0.0 : (w.level > 0 ? w.level * 0.05 : 0.15 * rng.dbl());
        int k = (int)(r * 6400.0); // generators.py:166
        // Draw uniform interior cells until exactly k distinct ones are set (a uniform k-subset, the law of
        // np.random.choice(6400, k, replace=False)).
        Philox mine;
        mine.init(w.seed, (uint32_t)e, episode, 0x100u + (uint32_t)lane);
        int count = 0;
        // bulk rounds: 4 candidates per lane (128 draws) while even an all-fresh round cannot overshoot k; the
        // old value of the shared-memory atomicOr says whether the cell was new
        while (k - count >= 128) {
            mine.block();
            int added = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int cand = (int)(mine.out[q] & 8191u);
                if (cand < 6400) {
                    int rr = cand / 80 + 1, cq = cand - (rr - 1) * 80 + 1;
                    uint32_t bit = map_bit_of(cq);
                    uint32_t old = atomicOr(&bm[map_word_index(rr, cq)], bit);
                    added += (old & bit) ? 0 : 1;
                }
            }
            mine.have = 0;
            count += __reduce_add_sync(0xFFFFFFFFu, added);
        }
        __syncwarp();
        // tail: fewer than 128 cells missing.  Four sub-rounds per Philox block; in each, every lane offers one cell, the
        // old value of the atomicOr tells whether it was new, and if more were new than are still needed the surplus
        // (highest lanes) is taken back.  The procedure never looks at WHICH cell a lane holds, so all k-subsets stay
        // equally likely -- same law as the reference's sampling without replacement.
        while (count < k) {
            mine.block();
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (count < k) { // warp-uniform
                    const int cand = (int)(mine.out[q] & 8191u);
                    bool fresh = false;
                    uint32_t bit = 0;
                    uint32_t *word = bm;
                    if (cand < 6400) {
                        const int rr = cand / 80 + 1, cq = cand - (rr - 1) * 80 + 1;
                        bit = map_bit_of(cq);
                        word = &bm[map_word_index(rr, cq)];
                        fresh = !(atomicOr(word, bit) & bit);
                    }
                    const uint32_t fmask = __ballot_sync(0xFFFFFFFFu, fresh);
                    const int need = k - count;
                    int nf = __popc(fmask);
                    if (nf > need) {
                        if (fresh && __popc(fmask & ((1u << lane) - 1u)) >= need) atomicAnd(word, ~bit);
                        nf = need;
                    }
                    count += nf;
                }
            }
            mine.have = 0;
            __syncwarp();
        }
    }

    uint32_t *gm = w.maps + (size_t)e * T2D_MAP_WORDS;
    for (int i = lane; i < T2D_MAP_WORDS; i += 32) gm[i] = bm[i];
    if (w.target_mode == T2D_TARGET_RPF) {
        uint32_t wallbits = rpf_clear_corners(bm, w.H, w.W, lane);
        if (lane == 0) w.rpf[e] = wallbits << 8; // static_goals(): vector = 0
        __syncwarp();
    }

    // spawn (generators.py:53-77): tracker uniform over free cells, target uniform over the free cells of the 2x2 block
    int nfree = bm_count_free(bm, lane);
    uint32_t tr;
    uint32_t goals = 0;
    if (w.target_mode == T2D_TARGET_RPF) {
        uint32_t g = rpf_next_goal(w, e, lane); // sample_goal(2): one cursor step, both goals that corner
        goals = g | (g << 16);
        tr = rpf_corner(w.H, w.W, 0);
    } else {
        tr = bm_select_free(bm, (int)rng.below((uint32_t)nfree), lane);
    }
    int r0 = tr & 255, c0 = tr >> 8;
    int cr[4], cc[4];
    int m = around_free(bm, r0, c0, cr, cc);
    int pick = (int)rng.below((uint32_t)m);
    int r1 = cr[0], c1 = cc[0];
#pragma unroll
    for (int q = 1; q < 4; q++)
        if (pick == q) { r1 = cr[q]; c1 = cc[q]; }
    uint32_t p = (uint32_t)r0 | ((uint32_t)c0 << 8) | ((uint32_t)r1 << 16) | ((uint32_t)c1 << 24);

    if (w.target_mode == T2D_TARGET_NAV) { // goals matter only to the Navigator (track_1v1.py:139-141)
        // sample_goal(2): two distinct uniform free cells, redrawn while the tracker starts on one (track_1v1.py:237-240)
        uint32_t g0, g1;
        do {
            int j0 = (int)rng.below((uint32_t)nfree);
            int j1 = (int)rng.below((uint32_t)(nfree - 1));
            if (j1 >= j0) j1++;
            g0 = bm_select_free(bm, j0, lane);
            g1 = bm_select_free(bm, j1, lane);
        } while (g0 == tr || g1 == tr);
        goals = g0 | (g1 << 16);
    }

    if (lane == 0) {
        w.pos[e] = p;
        w.goals[e] = goals;
    }
    if (!init_only) {
        if (w.target_mode == T2D_TARGET_RAM) {
            uint32_t word = ram_new_plan(rng, false, 0); // RamAgent.reset (navigator.py:90-93)
            if (lane == 0) w.ram[e] = word;
        }
        if (nav && defer_plan) {
            // auto-reset inside step(): the first plan is left to the replan wave at the start of the NEXT step (same Philox stream, same
            // goal, same start cell -> the same plan), so a step waits for ONE wave of A* plans instead of two
            if (lane == 0) {
                w.nav_meta[e] = 0u;  // length 0, cursor 0: "exhausted"
                w.nav_goal[e] = (goals >> 16) & 0xFFFFu;
            }
        } else if (nav) {
            Philox nrng;
            nrng.init(w.seed, (uint32_t)e, episode, 3u);
            PhiloxNavPolicy P{w, e, *nav, nrng};
            nav_plan_from(w, e, P, slot, r1, c1, goals >> 16, lane); // Navigator.reset(init_states[1], goal_states[1])
            if (w.async_nav && lane == 0)
                w.nav_end[e] = plan_end_cell(bm, w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES, 0, (int)(w.nav_meta[e] & 0xFFFFu), r1, c1);
        }
        if (lane == 0) {
            w.ctr[e] = 0;
            w.episode[e] = episode + 1;
        }
        if (obs && w.obs_type == T2D_OBS_PARTIAL) {
            if (w.target_mode == T2D_TARGET_RPF) { // observations come from the ENV map (corners not cleared)
                __syncwarp();
                for (int i = lane; i < T2D_MAP_WORDS; i += 32) bm[i] = gm[i];
                __syncwarp();
            }
            write_partial_obs(bm, p, obs + (size_t)e * T2D_ENV_CELLS, lane);
        }
    }
}

template <typename ObsT, int MAP>
__global__ void __launch_bounds__(128) reset_philox_kernel(World w, const uint8_t *__restrict__ mask, int from_list, ObsT *obs, int init_only) {
    __shared__ uint32_t bms[4][T2D_MAP_WORDS];
    __shared__ uint32_t walkw[4][MAP == T2D_MAP_MAZE ? T2D_WALK_WORDS : 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * 4 + wib, nw = gridDim.x * 4;
    const int n = from_list ? (int)w.work_count[0] : w.E;
    for (int i = gw; i < n; i += nw) {
        int e = i;
        if (from_list) e = (int)w.work_list[i];
        else if (mask && !mask[e]) continue;
        reset_env_philox<ObsT, MAP>(w, e, bms[wib], walkw[wib], nullptr, 0, obs, lane, init_only != 0);
        __syncwarp();
    }
    if (from_list) finish_reset(w);
}

// Nav / RPF: one warp per CTA, A* scratch in shared memory
template <typename ObsT, int MAP>
__global__ void __launch_bounds__(32) reset_philox_nav_kernel(World w, const uint8_t *__restrict__ mask, int from_list, ObsT *obs, int init_only) {
    __shared__ PhiloxNavScratch s;
    __shared__ uint32_t walkw[MAP == T2D_MAP_MAZE ? T2D_WALK_WORDS : 1];
    const int lane = threadIdx.x;
    const int n = from_list ? (int)w.work_count[0] : w.E;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        int e = i;
        if (from_list) e = (int)w.work_list[i];
        else if (mask && !mask[e]) continue;
        reset_env_philox<ObsT, MAP>(w, e, s.bm, walkw, &s, blockIdx.x, obs, lane, init_only != 0, from_list && !w.async_nav);
        __syncwarp();
    }
    if (from_list) finish_reset(w);
}

// ---- numpy-compat reset ------------------------------------------------------------------------------
template <typename ObsT>
__device__ __noinline__ void reset_env_numpy(const World &w, int e, NumpyScratch &s, int slot, ObsT *obs, int lane, bool init_only) {
    uint32_t *gkey = w.mt_key + (size_t)e * T2D_MT_N;
    for (int i = lane; i < T2D_MT_N; i += 32) s.key[i] = gkey[i];
    MtRng rng;
    rng.key = s.key;
    rng.pos = w.mt_pos[e];
    __syncwarp();
    // Only lane 0 draws; scalars it decides are broadcast with shuffles.
    bm_init(s.bm, w.H, w.W, lane);

    if (w.map_type == T2D_MAP_MAZE) {
        if (lane == 0) {
            double r = w.level > 0 ? w.level * 0.02 : .03 * rng.dbl(); // track_1v1.py:220-223
            maze_walk(s.bm, rng, r);
        }
        __syncwarp();
    } else {
        int k = 0;
        if (lane == 0) {
            double r = w.map_type == T2D_MAP_EMPTY ? 0.0 : (w.level > 0 ? w.level * 0.05 : 0.15 * rng.dbl()); // :226-231
            k = (int)(r * 6400.0);
        }
        k = __shfl_sync(0xFFFFFFFFu, k, 0);
        np_permutation(s, rng, 6400, lane); // np.random.choice(6400, k, replace=False) == permutation(6400)[:k]
        for (int t = lane; t < k; t += 32) {
            int idx = s.perm[t];
            bm_set(s.bm, idx / 80 + 1, idx % 80 + 1);
        }
        __syncwarp();
    }

    uint32_t *gm = w.maps + (size_t)e * T2D_MAP_WORDS;
    for (int i = lane; i < T2D_MAP_WORDS; i += 32) gm[i] = s.bm[i];
    if (w.target_mode == T2D_TARGET_RPF) {
        uint32_t wallbits = rpf_clear_corners(s.bm, w.H, w.W, lane);
        if (lane == 0) w.rpf[e] = wallbits << 8; // static_goals(): vector = 0
        __syncwarp();
    }

    uint32_t g[2];
    np_sample_goal(w, s, rng, e, 2, g, lane);

    // sample_close_states(2, 1) (generators.py:53-77)
    int n = np_free_list(s, w.H, w.W, lane);
    np_permutation(s, rng, n, lane);
    uint32_t tr;
    if (w.target_mode != T2D_TARGET_RPF) {
        int cell = s.freel[s.perm[0]];
        tr = (uint32_t)(cell / w.W) | ((uint32_t)(cell % w.W) << 8);
    } else {
        tr = rpf_corner(w.H, w.W, 0);
    }
    int r0 = tr & 255, c0 = tr >> 8;
    int cr[4], cc[4];
    int m = around_free(s.bm, r0, c0, cr, cc);
    int pick = 0;
    if (lane == 0) { // permutation(m)[0] with m <= 4, in registers
        int pm[4] = {0, 1, 2, 3};
        for (int i = m - 1; i >= 1; i--) {
            int j = (int)rng.interval((uint32_t)i);
            int t = pm[i]; pm[i] = pm[j]; pm[j] = t;
        }
        pick = pm[0];
    }
    pick = __shfl_sync(0xFFFFFFFFu, pick, 0);
    int r1 = cr[0], c1 = cc[0];
#pragma unroll
    for (int q = 1; q < 4; q++)
        if (pick == q) { r1 = cr[q]; c1 = cc[q]; }
    __syncwarp();
    np_permutation(s, rng, n, lane); // sample_state(0): choice(n, size=0, replace=False) still shuffles (generators.py:28 via :73)

    while (g[0] == tr || g[1] == tr) np_sample_goal(w, s, rng, e, 2, g, lane); // track_1v1.py:239-240

    uint32_t p = (uint32_t)r0 | ((uint32_t)c0 << 8) | ((uint32_t)r1 << 16) | ((uint32_t)c1 << 24);
    if (lane == 0) {
        w.pos[e] = p;
        w.goals[e] = g[0] | (g[1] << 16);
    }
    if (!init_only) {
        if (w.target_mode == T2D_TARGET_RAM) {
            if (lane == 0) w.ram[e] = ram_new_plan(rng, false, 0);
        }
        if (w.target_mode == T2D_TARGET_NAV || w.target_mode == T2D_TARGET_RPF) {
            NumpyNavPolicy P{w, e, s, rng};
            nav_plan_from(w, e, P, slot, r1, c1, g[1], lane); // Navigator.reset(init_states[1], goal_states[1])
        }
        if (lane == 0) {
            w.ctr[e] = 0;
            w.episode[e] = w.episode[e] + 1;
        }
        __syncwarp();
        if (obs && w.obs_type == T2D_OBS_PARTIAL) {
            if (w.target_mode == T2D_TARGET_RPF) { // observations come from the ENV map (corners not cleared)
                for (int i = lane; i < T2D_MAP_WORDS; i += 32) s.bm[i] = gm[i];
                __syncwarp();
            }
            write_partial_obs(s.bm, p, obs + (size_t)e * T2D_ENV_CELLS, lane);
        }
    }
    __syncwarp();
    for (int i = lane; i < T2D_MT_N; i += 32) gkey[i] = s.key[i];
    if (lane == 0) w.mt_pos[e] = rng.pos;
}

template <typename ObsT>
__global__ void __launch_bounds__(32) reset_numpy_kernel(World w, const uint8_t *__restrict__ mask, int from_list, ObsT *obs, int init_only) {
    __shared__ NumpyScratch s;
    const int lane = threadIdx.x;
    const int n = from_list ? (int)w.work_count[0] : w.E;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        int e = i;
        if (from_list) e = (int)w.work_list[i];
        else if (mask && !mask[e]) continue;
        reset_env_numpy<ObsT>(w, e, s, blockIdx.x, obs, lane, init_only != 0);
        __syncwarp();
    }
    if (from_list) finish_reset(w);
}

// ---- Navigator.step replan (navigator.py:15-36), run BEFORE the step kernel for envs whose plan is used up ----
__global__ void nav_scan_kernel(World w) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w.E) return;
    uint32_t meta = w.nav_meta[e];
    if ((meta >> 16) >= (meta & 0xFFFFu)) { // a_i >= len(plan_actions)
        uint32_t slot = atomicAdd(&w.work_count[1], 1u);
        w.work_list[w.E + slot] = (uint32_t)e;
    }
}

// load the generator maze of env e (ENV map, RPF corners cleared) into shared memory
__device__ __forceinline__ void load_gen_maze(const World &w, int e, uint32_t *bm, int lane) {
    const uint32_t *gm = w.maps + (size_t)e * T2D_MAP_WORDS;
    for (int i = lane; i < T2D_MAP_WORDS; i += 32) bm[i] = gm[i];
    __syncwarp();
    if (w.target_mode == T2D_TARGET_RPF) rpf_clear_corners(bm, w.H, w.W, lane);
}

__global__ void __launch_bounds__(32) nav_replan_numpy_kernel(World w) {
    __shared__ NumpyScratch s;
    const int lane = threadIdx.x;
    const int n = (int)w.work_count[1];
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int e = (int)w.work_list[w.E + i];
        uint32_t *gkey = w.mt_key + (size_t)e * T2D_MT_N;
        for (int q = lane; q < T2D_MT_N; q += 32) s.key[q] = gkey[q];
        MtRng rng;
        rng.key = s.key;
        rng.pos = w.mt_pos[e];
        load_gen_maze(w, e, s.bm, lane);
        NumpyNavPolicy P{w, e, s, rng};
        uint32_t goal = P.sample_goal1(lane); // self.goal_states = maze_generator.sample_goal(1)[0]
        uint32_t p = w.pos[e];
        nav_plan_from(w, e, P, blockIdx.x, (int)(p >> 16) & 255, (int)(p >> 24), goal, lane); // from old_state[1]
        for (int q = lane; q < T2D_MT_N; q += 32) gkey[q] = s.key[q];
        if (lane == 0) w.mt_pos[e] = rng.pos;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) nav_replan_philox_kernel(World w) {
    __shared__ PhiloxNavScratch s;
    const int lane = threadIdx.x;
    const int n = (int)w.work_count[1];
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int e = (int)w.work_list[w.E + i];
        const uint32_t elapsed = w.ctr[e] >> 16, p = w.pos[e];
        load_gen_maze(w, e, s.bm, lane);
        Philox rng;
        if (elapsed == 0u && w.nav_meta[e] == 0u) {
            // the first plan of an episode the last auto-reset started: Navigator.reset(init_states[1], goal_states[1]) with the stream
            // and the goal the reset kernel would have used
            const uint32_t goal = w.nav_goal[e];
            rng.init(w.seed, (uint32_t)e, w.episode[e] - 1u, 3u);
            PhiloxNavPolicy P{w, e, s, rng};
            nav_plan_from(w, e, P, blockIdx.x, (int)(p >> 16) & 255, (int)(p >> 24), goal, lane);
        } else {
            rng.init(w.seed, (uint32_t)e, w.episode[e], 0x20000u + elapsed);
            PhiloxNavPolicy P{w, e, s, rng};
            const uint32_t goal = P.sample_goal1(lane);
            nav_plan_from(w, e, P, blockIdx.x, (int)(p >> 16) & 255, (int)(p >> 24), goal, lane);
        }
        __syncwarp();
    }
}

// AstarSolver.solve (Astar_solver.py:121-149) called directly: plan from (sr, sc) to (gr, gc) on the generator maze of env
// first + i, for i < count.  The plan lands in that env's Navigator plan buffer (a_i = 0); len_out[i] = its length or -1.
__global__ void __launch_bounds__(32) astar_direct_kernel(World w, int first, int count, const int32_t *__restrict__ sg, int32_t *__restrict__ len_out) {
    __shared__ PhiloxNavScratch s;
    const int lane = threadIdx.x;
    for (int i = blockIdx.x; i < count; i += gridDim.x) {
        const int e = first + i;
        load_gen_maze(w, e, s.bm, lane);
        const int len = astar_plan(w, w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES, -1, s.bm, s.astar, blockIdx.x, sg[4 * i], sg[4 * i + 1], sg[4 * i + 2], sg[4 * i + 3], lane);
        if (lane == 0) {
            len_out[i] = len;
            w.nav_meta[e] = (uint32_t)(len > 0 ? len : 0);
            w.nav_goal[e] = (uint32_t)sg[4 * i + 2] | ((uint32_t)sg[4 * i + 3] << 8);
        }
        __syncwarp();
    }
}

// =====================================================================================================================
// Plan-ahead (Philox + Nav + auto-reset).  A scripted Nav target never reacts to the tracker, so its whole trajectory is a function
// of (seed, env, episode): plan k of an episode starts where plan k-1 ends and draws its goal from Philox(seed, env, episode,
// 0x20000 + first step of the plan) -- exactly the key the synchronous replan uses when the plan runs out.  So plans can be made
// BEFORE they are needed: nav_plan becomes an append-only ring (length and cursor of nav_meta only grow), a planner on a side stream
// keeps ~NAV_LOW actions ahead of every target and hands segments over through ext_info / ext_plan, the main stream merges them
// between steps; next-episode worlds ("standby": map, spawn, goals, first plans) are prepared on the side stream too, so that an
// auto-reset is a copy.  The synchronous replan stays as the fallback for a starved env; results do not depend on the timing.
constexpr int NAV_LOW = 24;   // the side-stream planner extends a plan once fewer than this many actions lie ahead of the target
constexpr int NAV_MIN = 12;   // a freshly reset world starts with at least this many (the planner's hand-over lags a few steps)

// one more segment for env e of world w, planned from cell `start` at absolute step `base_abs`, written to out[out_base ...)
__device__ int plan_segment(const World &w, int e, uint32_t episode, int base_abs, uint32_t start, uint8_t *out, int out_base, PhiloxNavScratch &s, int slot,
                            int lane, uint32_t &goal, uint32_t &end) {
    Philox rng;
    rng.init(w.seed, (uint32_t)e, episode, 0x20000u + (uint32_t)base_abs);
    PhiloxNavPolicy P{w, e, s, rng};
    goal = P.sample_goal1(lane);  // self.goal_states = maze_generator.sample_goal(1)[0]
    const int len = nav_plan_into(w, out, out_base, P, slot, (int)(start & 255u), (int)((start >> 8) & 255u), goal, lane);
    uint32_t en = 0;
    if (lane == 0) en = plan_end_cell(s.bm, out, out_base, len, (int)(start & 255u), (int)((start >> 8) & 255u));
    end = __shfl_sync(0xFFFFFFFFu, en, 0);
    return len;
}

// direct mode: extend the plan of every listed (or masked) env until NAV_MIN actions lie ahead.  Used where nothing else touches the
// world: on the live world right after a full reset (main stream) and on the standby world (side stream).
__global__ void __launch_bounds__(32) nav_fill_kernel(World w, const uint8_t *__restrict__ mask, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count) {
    __shared__ PhiloxNavScratch s;
    const int lane = threadIdx.x;
    const int n = list ? (int)*count : w.E;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        int e = i;
        if (list) e = (int)list[i];
        else if (mask && !mask[e]) continue;
        load_gen_maze(w, e, s.bm, lane);
        uint32_t meta = w.nav_meta[e], end = w.nav_end[e];
        const uint32_t episode = w.episode[e];  // (the value the synchronous replan keys on: already advanced by the reset)
        uint8_t *plan = w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES;
        while ((int)((meta & 0xFFFFu) - (meta >> 16)) < NAV_MIN) {
            uint32_t goal;
            const int len = plan_segment(w, e, episode, (int)(meta & 0xFFFFu), end, plan, (int)(meta & 0xFFFFu), s, blockIdx.x, lane, goal, end);
            meta += (uint32_t)len;
            if (lane == 0) w.nav_goal[e] = goal & 0xFFFFu;
        }
        if (lane == 0) {
            w.nav_meta[e] = meta;
            w.nav_end[e] = end;
        }
        __syncwarp();
    }
}

// side stream: which envs need another segment?  (ext state 0 -> 2)
__global__ void nav_ahead_scan_kernel(World w, uint32_t *__restrict__ list, uint32_t *__restrict__ count) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w.E) return;
    const uint32_t meta = *reinterpret_cast<volatile uint32_t *>(w.nav_meta + e);
    if ((int)((meta & 0xFFFFu) - (meta >> 16)) < NAV_LOW && (*reinterpret_cast<volatile uint32_t *>(&w.ext_info[e].x) & 3u) == 0u) {
        w.ext_info[e].x = 2u;
        list[atomicAdd(count, 1u)] = (uint32_t)e;
    }
}

// side stream: plan one segment per listed env from a snapshot of its state into the hand-over slot (ext state 2 -> 1)
__global__ void __launch_bounds__(32) nav_ahead_plan_kernel(World w, const uint32_t *__restrict__ list, uint32_t *__restrict__ count) {
    __shared__ PhiloxNavScratch s;
    const int lane = threadIdx.x;
    const int n = (int)*count;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int e = (int)list[i];
        const uint32_t episode = *reinterpret_cast<volatile uint32_t *>(w.episode + e);
        const uint32_t base = *reinterpret_cast<volatile uint32_t *>(w.nav_meta + e) & 0xFFFFu;
        __threadfence();
        const uint32_t start = *reinterpret_cast<volatile uint32_t *>(w.nav_end + e);
        load_gen_maze(w, e, s.bm, lane);
        uint32_t goal, end;
        uint8_t *out = w.ext_plan + (size_t)e * T2D_NAV_PLAN_BYTES;
        const int len = plan_segment(w, e, episode, (int)base, start, out, 0, s, blockIdx.x, lane, goal, end);
        if (lane == 0) {
            w.ext_info[e].y = base | ((uint32_t)len << 16);
            w.ext_info[e].z = goal & 0xFFFFu;
            w.ext_info[e].w = end;
            __threadfence();
            w.ext_info[e].x = 1u | ((episode & 0xFFFFFFu) << 8);
        }
        __syncwarp();
    }
    // the last CTA clears the queue for the next scan
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(count + 1, 1u) == gridDim.x - 1) { count[0] = 0; count[1] = 0; }
    }
}

// main stream, between steps: append finished segments that still fit (same episode, same length as when they were planned)
__global__ void nav_merge_kernel(World w) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w.E) return;
    if (!(*reinterpret_cast<volatile uint32_t *>(&w.ext_info[e].x) & 1u)) return;
    __threadfence();
    uint4 info;
    info.x = *reinterpret_cast<volatile uint32_t *>(&w.ext_info[e].x);
    info.y = *reinterpret_cast<volatile uint32_t *>(&w.ext_info[e].y);
    info.z = *reinterpret_cast<volatile uint32_t *>(&w.ext_info[e].z);
    info.w = *reinterpret_cast<volatile uint32_t *>(&w.ext_info[e].w);
    const uint32_t meta = w.nav_meta[e], base = info.y & 0xFFFFu, len = info.y >> 16;
    if ((info.x >> 8) == (w.episode[e] & 0xFFFFFFu) && base == (meta & 0xFFFFu)) {
        uint8_t *plan = w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES;
        const uint8_t *src = w.ext_plan + (size_t)e * T2D_NAV_PLAN_BYTES;
        for (uint32_t i = 0; i < len; i++) plan_put(plan, (int)(base + i), plan_get(src, (int)i));
        w.nav_goal[e] = info.z;
        w.nav_end[e] = info.w;
        __threadfence();  // a planner that sees the new length also sees the new end cell
        atomicAdd(&w.nav_meta[e], len);
    }
    __threadfence();
    w.ext_info[e].x = 0u;
}

// main stream: an auto-reset as a copy of the standby world sw (prepared on the side stream) into the live world w.  One warp per
// finished env (queued by the step kernel); writes the reset observation; queues the env for a new standby world.
template <typename ObsT>
__global__ void __launch_bounds__(128) swap_standby_kernel(World w, World sw, ObsT *obs, uint32_t *__restrict__ regen_list, uint32_t *__restrict__ regen_count) {
    __shared__ uint32_t bms[4][T2D_MAP_WORDS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = (int)w.work_count[0];
    for (int i = blockIdx.x * 4 + wib; i < n; i += gridDim.x * 4) {
        const int e = (int)w.work_list[i];
        uint32_t *bm = bms[wib];
        const uint32_t *src = sw.maps + (size_t)e * T2D_MAP_WORDS;
        uint32_t *dst = w.maps + (size_t)e * T2D_MAP_WORDS;
        for (int k = lane; k < T2D_MAP_WORDS; k += 32) { const uint32_t v = src[k]; bm[k] = v; dst[k] = v; }
        const uint32_t p = sw.pos[e];
        if (w.nav_plan) {
            const uint32_t *ps = reinterpret_cast<const uint32_t *>(sw.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES);
            uint32_t *pd = reinterpret_cast<uint32_t *>(w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES);
            for (int k = lane; k < T2D_NAV_PLAN_BYTES / 4; k += 32) pd[k] = ps[k];
        }
        __syncwarp();
        if (lane == 0) {
            w.pos[e] = p;
            w.goals[e] = sw.goals[e];
            w.ctr[e] = 0;
            if (w.ram) w.ram[e] = sw.ram[e];
            if (w.nav_plan) {
                w.nav_goal[e] = sw.nav_goal[e];
                w.nav_end[e] = sw.nav_end[e];
                __threadfence();  // map, plan and end cell before the new length / episode (the side-stream planner reads them in the opposite order)
                w.nav_meta[e] = sw.nav_meta[e];
            }
            const uint32_t ep = w.episode[e] + 1u;
            w.episode[e] = ep;
            sw.episode[e] = ep;  // key of the next standby world
            regen_list[atomicAdd(regen_count, 1u)] = (uint32_t)e;
        }
        __syncwarp();
        if (obs) write_partial_obs(bm, p, obs + (size_t)e * T2D_ENV_CELLS, lane);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&w.work_count[2], 1u) == gridDim.x - 1) {
            w.stats[0] += w.work_count[0];
            regen_count[3] = w.work_count[0];  // (the standby reset kernel consumes and clears regen_count[0]; the fill kernel reads [3])
            w.work_count[0] = 0;
            w.work_count[2] = 0;
            __threadfence();
        }
    }
}

// the synchronous fallback in plan-ahead mode: a starved env (cursor == length) gets its next segment appended on the main stream
__global__ void nav_starved_scan_kernel(World w) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= w.E) return;
    const uint32_t meta = w.nav_meta[e];
    if (((meta >> 16) & 0xFFFFu) == (meta & 0xFFFFu)) w.work_list[w.E + atomicAdd(&w.work_count[1], 1u)] = (uint32_t)e;
}
__global__ void __launch_bounds__(32) nav_starved_plan_kernel(World w) {
    __shared__ PhiloxNavScratch s;
    const int lane = threadIdx.x;
    const int n = (int)w.work_count[1];
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int e = (int)w.work_list[w.E + i];
        load_gen_maze(w, e, s.bm, lane);
        const uint32_t meta = w.nav_meta[e], p = w.pos[e];
        uint32_t goal, end;
        const uint32_t start = ((p >> 16) & 255u) | ((p >> 24) << 8);
        const int len = plan_segment(w, e, w.episode[e], (int)(meta & 0xFFFFu), start, w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES, (int)(meta & 0xFFFFu), s,
                                     blockIdx.x, lane, goal, end);
        if (lane == 0) {
            w.nav_meta[e] = meta + (uint32_t)len;
            w.nav_goal[e] = goal & 0xFFFFu;
            w.nav_end[e] = end;
        }
        __syncwarp();
    }
}

__global__ void finish_replan_kernel(World w) {
    if (threadIdx.x == 0 && blockIdx.x == 0) w.work_count[1] = 0;
}

// np.random.seed(seed) for env `e`: mt19937_seed (init_genrand)
__global__ void seed_numpy_kernel(World w, int first, int count, unsigned long long base_seed, int add_index) {
    int e = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= first + count) return;
    uint32_t seed = (uint32_t)(base_seed + (add_index ? (unsigned long long)e : 0ull));
    uint32_t *key = w.mt_key + (size_t)e * T2D_MT_N;
    for (int pos = 0; pos < T2D_MT_N; pos++) {
        key[pos] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
    }
    w.mt_pos[e] = T2D_MT_N;
}

} // namespace

#define T2D_NAV_GRID (148 * 10) /* one-warp planners in flight: ~20 KB of shared memory each */

template <typename ObsT, int MAP>
static cudaError_t launch_reset_map(const World &w, const uint8_t *mask, int from_list, ObsT *obs, int init_only, cudaStream_t s) {
    const bool nav = w.target_mode == T2D_TARGET_NAV || w.target_mode == T2D_TARGET_RPF;
    if (nav) {
        int grid = from_list ? T2D_NAV_GRID : min(w.E, T2D_NAV_GRID);
        reset_philox_nav_kernel<ObsT, MAP><<<grid, 32, 0, s>>>(w, mask, from_list, obs, init_only);
    } else {
        int grid = from_list ? 148 * 4 : min((w.E + 3) / 4, 148 * 8);
        reset_philox_kernel<ObsT, MAP><<<grid, 128, 0, s>>>(w, mask, from_list, obs, init_only);
    }
    return cudaGetLastError();
}

template <typename ObsT>
static cudaError_t launch_reset_t(const World &w, const uint8_t *mask, int from_list, ObsT *obs, int init_only, cudaStream_t s) {
    if (w.rng_mode == T2D_RNG_NUMPY) {
        int grid = from_list ? T2D_NAV_GRID : min(w.E, T2D_NAV_GRID);
        reset_numpy_kernel<ObsT><<<grid, 32, 0, s>>>(w, mask, from_list, obs, init_only);
        return cudaGetLastError();
    }
    switch (w.map_type) {
        case T2D_MAP_MAZE: return launch_reset_map<ObsT, T2D_MAP_MAZE>(w, mask, from_list, obs, init_only, s);
        case T2D_MAP_EMPTY: return launch_reset_map<ObsT, T2D_MAP_EMPTY>(w, mask, from_list, obs, init_only, s);
        default: return launch_reset_map<ObsT, T2D_MAP_BLOCK>(w, mask, from_list, obs, init_only, s);
    }
}

cudaError_t t2d_launch_reset_f32(const World &w, const uint8_t *mask, int from_list, float *obs, int init_only, cudaStream_t s) {
    return launch_reset_t<float>(w, mask, from_list, obs, init_only, s);
}
cudaError_t t2d_launch_reset_u8(const World &w, const uint8_t *mask, int from_list, uint8_t *obs, int init_only, cudaStream_t s) {
    return launch_reset_t<uint8_t>(w, mask, from_list, obs, init_only, s);
}
cudaError_t t2d_launch_seed_numpy(const World &w, int first, int count, unsigned long long seed, int add_index, cudaStream_t s) {
    seed_numpy_kernel<<<(count + 127) / 128, 128, 0, s>>>(w, first, count, seed, add_index);
    return cudaGetLastError();
}
// Navigator.step's replan for every env whose plan is exhausted (no-op for other target modes)
cudaError_t t2d_launch_nav_replan(const World &w, cudaStream_t s) {
    if (w.target_mode != T2D_TARGET_NAV && w.target_mode != T2D_TARGET_RPF) return cudaSuccess;
    nav_scan_kernel<<<(w.E + 255) / 256, 256, 0, s>>>(w);
    if (w.rng_mode == T2D_RNG_NUMPY) nav_replan_numpy_kernel<<<T2D_NAV_GRID, 32, 0, s>>>(w);
    else nav_replan_philox_kernel<<<T2D_NAV_GRID, 32, 0, s>>>(w);
    finish_replan_kernel<<<1, 32, 0, s>>>(w);
    return cudaGetLastError();
}
cudaError_t t2d_launch_astar_direct(const World &w, int first, int count, const int32_t *sg_dev, int32_t *len_dev, cudaStream_t s) {
    astar_direct_kernel<<<min(count, T2D_NAV_GRID), 32, 0, s>>>(w, first, count, sg_dev, len_dev);
    return cudaGetLastError();
}
// ---- plan-ahead launchers (see the block comment above nav_fill_kernel) ------------------------------------------------------------
cudaError_t t2d_launch_nav_fill(const World &w, const uint8_t *mask, const uint32_t *list, const uint32_t *count, cudaStream_t s) {
    nav_fill_kernel<<<min(w.E, T2D_NAV_GRID), 32, 0, s>>>(w, mask, list, count);
    return cudaGetLastError();
}
cudaError_t t2d_launch_nav_ahead(const World &w, uint32_t *list, uint32_t *count, cudaStream_t s) {
    nav_ahead_scan_kernel<<<(w.E + 255) / 256, 256, 0, s>>>(w, list, count);
    nav_ahead_plan_kernel<<<T2D_NAV_GRID, 32, 0, s>>>(w, list, count);
    return cudaGetLastError();
}
cudaError_t t2d_launch_nav_merge(const World &w, cudaStream_t s) {
    nav_merge_kernel<<<(w.E + 255) / 256, 256, 0, s>>>(w);
    nav_starved_scan_kernel<<<(w.E + 255) / 256, 256, 0, s>>>(w);
    nav_starved_plan_kernel<<<T2D_NAV_GRID, 32, 0, s>>>(w);
    finish_replan_kernel<<<1, 32, 0, s>>>(w);
    return cudaGetLastError();
}
cudaError_t t2d_launch_swap_standby(const World &w, const World &sw, void *obs, int obs_u8, uint32_t *regen_list, uint32_t *regen_count, cudaStream_t s) {
    if (obs_u8) swap_standby_kernel<uint8_t><<<148 * 2, 128, 0, s>>>(w, sw, (uint8_t *)obs, regen_list, regen_count);
    else swap_standby_kernel<float><<<148 * 2, 128, 0, s>>>(w, sw, (float *)obs, regen_list, regen_count);
    return cudaGetLastError();
}
int t2d_nav_slots() { return T2D_NAV_GRID; }
