/*
 * track2d_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference's gym-track2d environment (zfw1226/active_tracking_rl),
 * one env per object, scalar, single-threaded per env.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (libtrack2d.so, CUDA)
 * never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file bit-for-bit against fixtures
 * recorded from the unmodified reference (oracle/refharness/make_golden.py -> tests/golden/ fixtures)
 * and, when /root/reference is present, against the live reference.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root;
 * "envs/" = envs/gym-track2d/gym_track2d/envs/).
 *
 * Third-party arithmetic restated here because it is not under the reference tree:
 *   - numpy legacy RandomState (reference pins numpy==1.14.0, requirements.txt:5): MT19937
 *     init_genrand seeding, genrand_int32, random_sample (53-bit double), rk_interval masked
 *     rejection (used by shuffle/permutation and by randint), choice(replace=False) == full
 *     permutation(n)[:k], choice(replace=True) == randint(0, n, size).
 *   - CPython heapq (_siftdown on push; leaf-then-up _siftup on pop) as used by envs/Astar_solver.py.
 *   - gym==0.12.5 TimeLimit (done once elapsed_steps >= max_episode_steps).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define O_MAXDIM 82
#define O_MAXCELLS (O_MAXDIM * O_MAXDIM)
#define O_POB 6
#define O_WIN 13
#define O_NAV_MAXPLAN 8192

enum { O_MAP_BLOCK = 0, O_MAP_MAZE = 1, O_MAP_EMPTY = 2 };
enum { O_OBS_PARTIAL = 0, O_OBS_FULL = 1 };
enum { O_T_ADV = 0, O_T_PZR = 1, O_T_FAR = 2, O_T_NAV = 3, O_T_RAM = 4, O_T_RPF = 5 };

/* ------------------------------------------------------------------------------------------- */
/* numpy legacy RandomState (MT19937)                                                           */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t key[624];
    int pos;
    uint64_t draws; /* number of 32-bit outputs consumed (diagnostics) */
} o_mt;

/* numpy/random/src/mt19937/mt19937.c: mt19937_seed (what np.random.seed(int) calls) */
static void mt_seed(o_mt *s, uint32_t seed) {
    for (int pos = 0; pos < 624; pos++) {
        s->key[pos] = seed;
        seed = (1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u);
    }
    s->pos = 624;
    s->draws = 0;
}

static void mt_gen(o_mt *s) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    int i;
    uint32_t y;
    for (i = 0; i < 624 - 397; i++) {
        y = (s->key[i] & UPPER) | (s->key[i + 1] & LOWER);
        s->key[i] = s->key[i + 397] ^ (y >> 1) ^ (-(y & 1) & MAT);
    }
    for (; i < 623; i++) {
        y = (s->key[i] & UPPER) | (s->key[i + 1] & LOWER);
        s->key[i] = s->key[i + (397 - 624)] ^ (y >> 1) ^ (-(y & 1) & MAT);
    }
    y = (s->key[623] & UPPER) | (s->key[0] & LOWER);
    s->key[623] = s->key[396] ^ (y >> 1) ^ (-(y & 1) & MAT);
    s->pos = 0;
}

static uint32_t mt_u32(o_mt *s) {
    if (s->pos == 624) mt_gen(s);
    uint32_t y = s->key[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    s->draws++;
    return y;
}

/* np.random.random() / random_sample(): legacy 53-bit double from two draws */
static double mt_double(o_mt *s) {
    int32_t a = (int32_t)(mt_u32(s) >> 5), b = (int32_t)(mt_u32(s) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* legacy rk_interval / random_interval and buffered_bounded_masked_uint32: uniform in [0, max]
 * by masked rejection; max == 0 consumes NO draw. */
static uint32_t mt_interval(o_mt *s, uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max, v;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    while ((v = (mt_u32(s) & mask)) > max) {
    }
    return v;
}

/* np.random.randint(low, high): low + interval(high - 1 - low) */
static int mt_randint(o_mt *s, int low, int high) { return low + (int)mt_interval(s, (uint32_t)(high - 1 - low)); }

/* np.random.permutation(n): arange + legacy shuffle (i = n-1 .. 1, j = interval(i), swap) */
static void mt_permutation(o_mt *s, int n, int *out) {
    for (int i = 0; i < n; i++) out[i] = i;
    for (int i = n - 1; i >= 1; i--) {
        int j = (int)mt_interval(s, (uint32_t)i);
        int t = out[i];
        out[i] = out[j];
        out[j] = t;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* env object                                                                                   */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    int map_type, obs_type, target_mode, level;
    int H, W;                       /* 82x82 (Block/Empty) or 81x81 (Maze) */
    uint8_t maze[O_MAXCELLS];       /* Track1v1Env.maze: a COPY taken in init_maze (track_1v1.py:234) */
    uint8_t gen_maze[O_MAXCELLS];   /* maze_generator.maze: differs from `maze` only for RPF (static_goals) */
    int state[2][2];                /* [agent][row, col] */
    int goals[2][2];
    int is_static, vector, cand[4][2];
    double w_p;
    int c_far, elapsed, max_episode_steps;
    int collided[2];
    double distance;
    /* RamAgent */
    int ram_plan[16], ram_len, ram_i;
    /* Navigator */
    int nav_plan[O_NAV_MAXPLAN], nav_len, nav_i, nav_goal[2];
    int nav_replans, nav_planb, astar_expansions, astar_replaces;
    o_mt rng;
    int perm[O_MAXCELLS];
    int freelist[O_MAXCELLS];
} OEnv;

#define CELL(e, r, c) ((r) * (e)->W + (c))

/* ------------------------------------------------------------------------------------------- */
/* generators.py                                                                                */
/* ------------------------------------------------------------------------------------------- */

/* generators.py:157-176 RandomBlockMazeGenerator._generate_maze: k = int(ratio * 80**2) distinct
 * interior cells via np.random.choice(6400, k, replace=False) (== permutation(6400)[:k]); flat index
 * i -> (i // 80, i % 80) (cartesian_product order); pad a 1-cell wall border -> 82x82. */
static void gen_block(OEnv *e, double ratio) {
    const int n = 80;
    e->H = e->W = n + 2;
    memset(e->gen_maze, 0, sizeof(e->gen_maze));
    for (int i = 0; i < e->H; i++) {
        e->gen_maze[CELL(e, i, 0)] = e->gen_maze[CELL(e, i, e->W - 1)] = 1;
        e->gen_maze[CELL(e, 0, i)] = e->gen_maze[CELL(e, e->H - 1, i)] = 1;
    }
    int k = (int)(ratio * (double)(n * n));
    mt_permutation(&e->rng, n * n, e->perm); /* the full shuffle happens even when k == 0 */
    for (int t = 0; t < k; t++) {
        int idx = e->perm[t];
        e->gen_maze[CELL(e, idx / n + 1, idx % n + 1)] = 1;
    }
}

/* generators.py:115-145 RandomMazeGenerator._generate_maze (width = height = 80 -> shape 81x81) */
static void gen_maze_walls(OEnv *e, double r) {
    const int sh = (80 / 2) * 2 + 1; /* 81 */
    e->H = e->W = sh;
    int complexity = (int)(r * (double)(5 * (sh + sh)));
    int density = (int)(r * (double)((sh / 2) * (sh / 2)));
    memset(e->gen_maze, 0, sizeof(e->gen_maze));
    for (int i = 0; i < sh; i++) {
        e->gen_maze[CELL(e, 0, i)] = e->gen_maze[CELL(e, sh - 1, i)] = 1;
        e->gen_maze[CELL(e, i, 0)] = e->gen_maze[CELL(e, i, sh - 1)] = 1;
    }
    for (int i = 0; i < density; i++) {
        int x = mt_randint(&e->rng, 0, sh / 2 + 1) * 2; /* x drawn before y (tuple evaluation order) */
        int y = mt_randint(&e->rng, 0, sh / 2 + 1) * 2;
        e->gen_maze[CELL(e, y, x)] = 1;
        for (int j = 0; j < complexity; j++) {
            int ny[4], nx[4], nn = 0;
            if (x > 1) { ny[nn] = y; nx[nn] = x - 2; nn++; }
            if (x < sh - 2) { ny[nn] = y; nx[nn] = x + 2; nn++; }
            if (y > 1) { ny[nn] = y - 2; nx[nn] = x; nn++; }
            if (y < sh - 2) { ny[nn] = y + 2; nx[nn] = x; nn++; }
            if (nn) {
                int pick = mt_randint(&e->rng, 0, nn);
                int y_ = ny[pick], x_ = nx[pick];
                if (e->gen_maze[CELL(e, y_, x_)] == 0) {
                    e->gen_maze[CELL(e, y_, x_)] = 1;
                    /* Python floor division: (y - y_) // 2 is +-1 or 0 */
                    e->gen_maze[CELL(e, y_ + (y - y_) / 2, x_ + (x - x_) / 2)] = 1;
                    x = x_;
                    y = y_;
                }
            }
        }
    }
}

/* np.where(maze == 0) in row-major order */
static int free_cells(const OEnv *e, int *out) {
    int n = 0;
    for (int i = 0; i < e->H * e->W; i++)
        if (e->gen_maze[i] == 0) out[n++] = i;
    return n;
}

/* generators.py:12-19 static_goals (RPF) */
static void static_goals(OEnv *e) {
    e->is_static = 1;
    int a0 = (int)(e->H / 6.0), a1 = (int)(e->H * 5 / 6.0), b0 = (int)(e->W / 6.0), b1 = (int)(e->W * 5 / 6.0);
    int c[4][2] = {{a0, b0}, {a1, b0}, {a1, b1}, {a0, b1}};
    memcpy(e->cand, c, sizeof(c));
    for (int g = 0; g < 4; g++) e->gen_maze[CELL(e, c[g][0], c[g][1])] = 0;
    e->vector = 0;
}

/* generators.py:38-51 sample_goal(num): num distinct free cells = free[permutation(n)[:num]];
 * static (RPF): advance the corner cursor once, every goal is that corner. */
static void sample_goal(OEnv *e, int num, int goals[][2]) {
    if (!e->is_static) {
        int n = free_cells(e, e->freelist);
        mt_permutation(&e->rng, n, e->perm);
        for (int i = 0; i < num; i++) {
            int cell = e->freelist[e->perm[i]];
            goals[i][0] = cell / e->W;
            goals[i][1] = cell % e->W;
        }
    } else {
        e->vector = (e->vector + 1) % 4;
        for (int i = 0; i < num; i++) {
            goals[i][0] = e->cand[e->vector][0];
            goals[i][1] = e->cand[e->vector][1];
        }
    }
}

/* generators.py:82-94 get_around(state, 1): window rows [max(0,r-1), min(H-1,r+1)) x cols likewise
 * (slice end EXCLUSIVE => the 2x2 block {r-1,r}x{c-1,c}); choice(len, 1, replace=False). */
static void get_around(OEnv *e, const int st[2], int md, int out[2]) {
    int x0 = st[0] - md > 0 ? st[0] - md : 0;
    int x1 = st[0] + md < e->H - 1 ? st[0] + md : e->H - 1;
    int y0 = st[1] - md > 0 ? st[1] - md : 0;
    int y1 = st[1] + md < e->W - 1 ? st[1] + md : e->W - 1;
    int pr[16], pc[16], n = 0;
    for (int r = x0; r < x1; r++)
        for (int c = y0; c < y1; c++)
            if (e->gen_maze[CELL(e, r, c)] == 0) { pr[n] = r; pc[n] = c; n++; }
    int perm[16];
    mt_permutation(&e->rng, n, perm);
    out[0] = pr[perm[0]];
    out[1] = pc[perm[0]];
}

/* generators.py:53-77 sample_close_states(2, 1): choice(n, 2, replace=False) (only idxs[0] is
 * used); tracker = that free cell (or corner 0 when static); target = get_around(tracker);
 * then sample_state(0) still runs a full permutation(n) (choice(n, size=0, replace=False)). */
static void sample_close_states(OEnv *e) {
    int n = free_cells(e, e->freelist);
    mt_permutation(&e->rng, n, e->perm);
    if (!e->is_static) {
        int cell = e->freelist[e->perm[0]];
        e->state[0][0] = cell / e->W;
        e->state[0][1] = cell % e->W;
    } else {
        e->state[0][0] = e->cand[0][0];
        e->state[0][1] = e->cand[0][1];
    }
    get_around(e, e->state[0], 1, e->state[1]);
    n = free_cells(e, e->freelist);         /* sample_state(num - 2 == 0), generators.py:21-36 */
    mt_permutation(&e->rng, n, e->perm);
}

/* track_1v1.py:218-240 init_maze */
void o_init_maze(OEnv *e) {
    e->is_static = 0;
    if (e->map_type == O_MAP_MAZE) {
        double r = e->level > 0 ? e->level * 0.02 : .03 * mt_double(&e->rng);
        gen_maze_walls(e, r);
    } else if (e->map_type == O_MAP_BLOCK) {
        double r = e->level > 0 ? e->level * 0.05 : 0.15 * mt_double(&e->rng);
        gen_block(e, r);
    } else {
        gen_block(e, 0.0);
    }
    memcpy(e->maze, e->gen_maze, sizeof(e->maze)); /* np.array(get_maze()) copy, BEFORE static_goals */
    if (e->target_mode == O_T_RPF) static_goals(e);
    sample_goal(e, 2, e->goals);
    sample_close_states(e);
    /* while goal_test(init_states[0]): resample goals (track_1v1.py:239-240, 264-269) */
    while ((e->state[0][0] == e->goals[0][0] && e->state[0][1] == e->goals[0][1]) ||
           (e->state[0][0] == e->goals[1][0] && e->state[0][1] == e->goals[1][1]))
        sample_goal(e, 2, e->goals);
}

/* ------------------------------------------------------------------------------------------- */
/* Astar_solver.py (+ CPython heapq)                                                            */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int cell, prev, action, g; } ANode;
typedef struct { double f; int node; } AEntry;
typedef struct {
    ANode *nodes; int nn, ncap;
    AEntry *heap; int hn, hcap;
    int *in_frontier; /* cell -> node id or -1  (Frontier.state_nodes) */
    uint8_t *explored;
} AStar;

static const int DR[4] = {-1, +1, 0, 0}, DC[4] = {0, 0, -1, +1}; /* track_1v1.py:276 / Astar_solver.py:162 */

/* list comparison [f, node] < [f2, node2]: floats first; on equal f falls through to
 * Node.__lt__ (path_cost <), Astar_solver.py:30-32,55 */
static int entry_lt(const AStar *a, AEntry x, AEntry y) {
    if (x.f != y.f) return x.f < y.f;
    if (x.node == y.node) return 0;
    return a->nodes[x.node].g < a->nodes[y.node].g;
}

static void hq_siftdown(AStar *a, int startpos, int pos) { /* heapq._siftdown */
    AEntry newitem = a->heap[pos];
    while (pos > startpos) {
        int parentpos = (pos - 1) >> 1;
        AEntry parent = a->heap[parentpos];
        if (entry_lt(a, newitem, parent)) {
            a->heap[pos] = parent;
            pos = parentpos;
            continue;
        }
        break;
    }
    a->heap[pos] = newitem;
}

static void hq_siftup(AStar *a, int pos) { /* heapq._siftup */
    int endpos = a->hn, startpos = pos;
    AEntry newitem = a->heap[pos];
    int childpos = 2 * pos + 1;
    while (childpos < endpos) {
        int rightpos = childpos + 1;
        if (rightpos < endpos && !entry_lt(a, a->heap[childpos], a->heap[rightpos])) childpos = rightpos;
        a->heap[pos] = a->heap[childpos];
        pos = childpos;
        childpos = 2 * pos + 1;
    }
    a->heap[pos] = newitem;
    hq_siftdown(a, startpos, pos);
}

static int new_node(AStar *a, int cell, int prev, int action) {
    if (a->nn == a->ncap) {
        a->ncap *= 2;
        a->nodes = (ANode *)realloc(a->nodes, sizeof(ANode) * a->ncap);
    }
    ANode *n = &a->nodes[a->nn];
    n->cell = cell; n->prev = prev; n->action = action;
    n->g = prev < 0 ? 0 : a->nodes[prev].g + 1; /* Node.__init__, uniform step cost (:157-159) */
    return a->nn++;
}

static double heuristic(const OEnv *e, int cell, const int goal[2]) { /* :151-153 np.linalg.norm */
    double dr = (double)(cell / e->W - goal[0]), dc = (double)(cell % e->W - goal[1]);
    return sqrt(dr * dr + dc * dc);
}

static void frontier_add(AStar *a, const OEnv *e, int node, const int goal[2]) { /* Frontier.add :53-56 */
    if (a->hn == a->hcap) {
        a->hcap *= 2;
        a->heap = (AEntry *)realloc(a->heap, sizeof(AEntry) * a->hcap);
    }
    AEntry en;
    en.f = (double)a->nodes[node].g + heuristic(e, a->nodes[node].cell, goal);
    en.node = node;
    a->heap[a->hn++] = en;
    hq_siftdown(a, 0, a->hn - 1);
    a->in_frontier[a->nodes[node].cell] = node;
}

/* Astar_solver.py:121-149 _astar_search on `maze`; returns plan length (>= 0) or -1 if unsolvable.
 * The plan (get_actions, :102-110) is written to `plan`. */
static int astar(OEnv *e, const uint8_t *maze, const int start[2], const int goal[2], int *plan, int maxplan) {
    AStar a;
    int ncell = e->H * e->W;
    a.ncap = 1024; a.nn = 0; a.nodes = (ANode *)malloc(sizeof(ANode) * a.ncap);
    a.hcap = 1024; a.hn = 0; a.heap = (AEntry *)malloc(sizeof(AEntry) * a.hcap);
    a.in_frontier = (int *)malloc(sizeof(int) * ncell);
    a.explored = (uint8_t *)calloc(ncell, 1);
    for (int i = 0; i < ncell; i++) a.in_frontier[i] = -1;
    int goal_cell = goal[0] * e->W + goal[1];
    int sol = -1;
    frontier_add(&a, e, new_node(&a, start[0] * e->W + start[1], -1, -1), goal);
    while (a.hn > 0) {
        /* Frontier.pop :58-63 == heapq.heappop */
        AEntry last = a.heap[--a.hn], ret = last;
        if (a.hn > 0) {
            ret = a.heap[0];
            a.heap[0] = last;
            hq_siftup(&a, 0);
        }
        int node = ret.node, cell = a.nodes[node].cell;
        a.in_frontier[cell] = -1;
        if (cell == goal_cell) { sol = node; break; }
        a.explored[cell] = 1;
        e->astar_expansions++;
        for (int act = 0; act < 4; act++) {
            int r = cell / e->W + DR[act], c = cell % e->W + DC[act];
            int child_cell = maze[r * e->W + c] == 1 ? cell : r * e->W + c; /* :156-172 wall => stay */
            int child_g = a.nodes[node].g + 1;
            if (!a.explored[child_cell] && a.in_frontier[child_cell] < 0) {
                frontier_add(&a, e, new_node(&a, child_cell, node, act), goal);
            } else if (a.in_frontier[child_cell] >= 0 && a.nodes[a.in_frontier[child_cell]].g < child_g) {
                /* Frontier.replace :65-73 (the comparison is inverted in the reference: it swaps in
                 * the WORSE node; restated as is).  Linear scan, overwrite, _siftdown(heap, 0, i). */
                e->astar_replaces++;
                int child = new_node(&a, child_cell, node, act);
                for (int i = 0; i < a.hn; i++) {
                    if (a.nodes[a.heap[i].node].cell == child_cell) {
                        a.heap[i].f = (double)child_g + heuristic(e, child_cell, goal);
                        a.heap[i].node = child;
                        hq_siftdown(&a, 0, i);
                        a.in_frontier[child_cell] = child;
                    }
                }
            }
        }
    }
    int len = -1;
    if (sol >= 0) {
        len = a.nodes[sol].g;
        if (len > maxplan) len = maxplan; /* never hit on 82x82 maps; guarded for safety */
        int n = sol;
        for (int i = a.nodes[sol].g - 1; i >= 0; i--) {
            if (i < maxplan) plan[i] = a.nodes[n].action;
            n = a.nodes[n].prev;
        }
    }
    free(a.nodes); free(a.heap); free(a.in_frontier); free(a.explored);
    return len;
}

/* ------------------------------------------------------------------------------------------- */
/* navigator.py                                                                                 */
/* ------------------------------------------------------------------------------------------- */

/* shared tail of Navigator.reset (:38-63) and Navigator.step's replan (:15-37): A* to the current
 * goal; while unsolvable or empty plan: up to 5 fresh goals via sample_goal(1); then plan B =
 * np.random.choice(all_actions, 10). */
static void nav_plan_from(OEnv *e, const int start[2]) {
    int count_res = 0, planb = 0;
    int len = astar(e, e->gen_maze, start, e->nav_goal, e->nav_plan, O_NAV_MAXPLAN);
    while (len < 1) {
        count_res++;
        if (count_res > 5) { planb = 1; break; }
        int g[1][2];
        sample_goal(e, 1, g);
        e->nav_goal[0] = g[0][0]; e->nav_goal[1] = g[0][1];
        len = astar(e, e->gen_maze, start, e->nav_goal, e->nav_plan, O_NAV_MAXPLAN);
    }
    if (planb) {
        e->nav_planb++;
        for (int i = 0; i < 10; i++) e->nav_plan[i] = mt_randint(&e->rng, 0, 4);
        len = 10;
    }
    e->nav_len = len;
    e->nav_i = 0;
}

/* navigator.py:11-36 Navigator.step.  _goal_test always returns None (goal_states is a flat
 * [r, c], :65-70) so a replan happens only when the plan is exhausted; ref_goal is None at the
 * only call site (track_1v1.py:84) so the new goal is always sample_goal(1)[0]. */
static int nav_step(OEnv *e, const int state[2]) {
    if (e->nav_i >= e->nav_len) {
        int g[1][2];
        sample_goal(e, 1, g);
        e->nav_goal[0] = g[0][0]; e->nav_goal[1] = g[0][1];
        e->nav_replans++;
        nav_plan_from(e, state);
    }
    return e->nav_plan[e->nav_i++];
}

/* navigator.py:77-93 RamAgent.  reset: plan = choice(4, randint(1, 10)) -- randint is evaluated
 * first (argument evaluation order). */
static void ram_reset(OEnv *e) {
    e->ram_len = mt_randint(&e->rng, 1, 10);
    for (int i = 0; i < e->ram_len; i++) e->ram_plan[i] = mt_randint(&e->rng, 0, 4);
    e->ram_i = 0;
}

static int ram_step(OEnv *e) {
    int action = e->ram_plan[e->ram_i];
    e->ram_i++;
    if (e->ram_i >= e->ram_len) {
        if (mt_randint(&e->rng, 0, 2) == 0) {       /* np.random.choice([0, 1], 1) == 0 */
            action = mt_randint(&e->rng, 0, 4);      /* the NEW action is what step() returns */
            e->ram_len = mt_randint(&e->rng, 1, 10); /* np.ones(randint(1, 10)) * action */
            for (int i = 0; i < e->ram_len; i++) e->ram_plan[i] = action;
        } else {
            e->ram_len = mt_randint(&e->rng, 1, 10);
            for (int i = 0; i < e->ram_len; i++) e->ram_plan[i] = mt_randint(&e->rng, 0, 4);
        }
        e->ram_i = 0;
    }
    return action;
}

/* ------------------------------------------------------------------------------------------- */
/* track_1v1.py: observations                                                                   */
/* ------------------------------------------------------------------------------------------- */

/* _get_full_obs :295-307: map copy, tracker cell = 2, then target cell = 4 */
static void full_obs(const OEnv *e, uint8_t *out) {
    memcpy(out, e->maze, (size_t)(e->H * e->W));
    out[CELL(e, e->state[0][0], e->state[0][1])] = 2;
    out[CELL(e, e->state[1][0], e->state[1][1])] = 4;
}

/* _get_partial_obs :309-326: 13x13 window round agent `id`; its own centre re-stamped with its own
 * colour; cells outside the map read 1 (np.pad constant 1). */
static void partial_obs(const OEnv *e, int id, uint8_t *out) {
    uint8_t local[O_MAXCELLS];
    full_obs(e, local);
    local[CELL(e, e->state[id][0], e->state[id][1])] = (uint8_t)(2 + 2 * id);
    for (int wr = 0; wr < O_WIN; wr++)
        for (int wc = 0; wc < O_WIN; wc++) {
            int r = e->state[id][0] - O_POB + wr, c = e->state[id][1] - O_POB + wc;
            out[wr * O_WIN + wc] = (r < 0 || c < 0 || r >= e->H || c >= e->W) ? 1 : local[CELL(e, r, c)];
        }
}

int o_obs_cells(const OEnv *e) { return e->obs_type == O_OBS_FULL ? e->H * e->W : O_WIN * O_WIN; }

/* _get_obs :287-293; out is [2][cells] uint8 (the reference's float64/int64 values are 0,1,2,4) */
void o_get_obs(const OEnv *e, uint8_t *out) {
    int n = o_obs_cells(e);
    for (int i = 0; i < 2; i++) {
        if (e->obs_type == O_OBS_FULL) full_obs(e, out + i * n);
        else partial_obs(e, i, out + i * n);
    }
}

/* ------------------------------------------------------------------------------------------- */
/* track_1v1.py: reset / step                                                                   */
/* ------------------------------------------------------------------------------------------- */

/* the tail of Track1v1Env.reset after init_maze (:137-168) */
static void reset_tail(OEnv *e) {
    if (e->target_mode == O_T_NAV || e->target_mode == O_T_RPF) {
        /* Navigator.reset(init_states[1], goal_states[1], gen) :38-63 */
        e->nav_goal[0] = e->goals[1][0];
        e->nav_goal[1] = e->goals[1][1];
        nav_plan_from(e, e->state[1]);
    }
    if (e->target_mode == O_T_RAM) ram_reset(e);
    e->w_p = e->target_mode == O_T_PZR ? 1.0 : (e->target_mode == O_T_FAR ? -0.5 : 0.0);
    e->c_far = 0;
    e->elapsed = 0; /* TimeLimit.reset */
    e->collided[0] = e->collided[1] = 0;
}

void o_reset(OEnv *e, uint8_t *obs) {
    o_init_maze(e);
    reset_tail(e);
    if (obs) o_get_obs(e, obs);
}

/* Track1v1Env.step :71-127 followed by gym 0.12.5 TimeLimit.step.  actions[1] is overridden for
 * Ram / Nav / RPF.  Returns done (after TimeLimit); *done_env gets the env's own far-counter done;
 * *target_action the action the target actually executed. */
int o_step(OEnv *e, const int *actions, uint8_t *obs, double *rewards, int *done_env, int *target_action) {
    int act[2] = {actions[0], actions[1]};
    int old1[2] = {e->state[1][0], e->state[1][1]};
    if (e->target_mode == O_T_RAM) act[1] = ram_step(e);
    if (e->target_mode == O_T_NAV || e->target_mode == O_T_RPF) act[1] = nav_step(e, old1);
    for (int i = 0; i < 2; i++) { /* _next_state :271-285 */
        int r = e->state[i][0] + DR[act[i]], c = e->state[i][1] + DC[act[i]];
        if (e->maze[CELL(e, r, c)] == 1) {
            e->collided[i] = 1;
        } else {
            e->state[i][0] = r; e->state[i][1] = c; e->collided[i] = 0;
        }
    }
    double dr = (double)(e->state[1][0] - e->state[0][0]), dc = (double)(e->state[1][1] - e->state[0][1]);
    double distance = sqrt(dr * dr + dc * dc);      /* np.linalg.norm :94 */
    double max_distance = (double)O_POB;
    double r_track = 1 - 2 * distance / max_distance; /* :99-104 */
    r_track = r_track > -1 ? r_track : -1;
    double over = distance - max_distance > 0 ? distance - max_distance : 0;
    double r_target = -r_track - e->w_p * over / max_distance;
    r_target = r_target > -1 ? r_target : -1;
    rewards[0] = r_track;
    rewards[1] = r_target;
    if (distance <= max_distance) e->c_far = 0; else e->c_far += 1; /* :106-111 */
    int done = e->c_far > 10;
    e->distance = distance;
    if (obs) o_get_obs(e, obs);
    if (done_env) *done_env = done;
    if (target_action) *target_action = act[1];
    e->elapsed += 1; /* TimeLimit */
    if (e->max_episode_steps > 0 && e->elapsed >= e->max_episode_steps) done = 1;
    return done;
}

/* ------------------------------------------------------------------------------------------- */
/* object plumbing + injection (tests)                                                          */
/* ------------------------------------------------------------------------------------------- */
OEnv *o_create(int map_type, int obs_type, int target_mode, int level) {
    OEnv *e = (OEnv *)calloc(1, sizeof(OEnv));
    e->map_type = map_type; e->obs_type = obs_type; e->target_mode = target_mode; e->level = level;
    e->max_episode_steps = 500; /* gym_track2d/__init__.py:17 */
    e->H = e->W = map_type == O_MAP_MAZE ? 81 : 82;
    mt_seed(&e->rng, 0);
    return e;
}
void o_destroy(OEnv *e) { free(e); }
void o_seed(OEnv *e, uint32_t seed) { mt_seed(&e->rng, seed); }
uint64_t o_rng_draws(const OEnv *e) { return e->rng.draws; }
uint32_t o_rng_u32(OEnv *e) { return mt_u32(&e->rng); }
double o_rng_double(OEnv *e) { return mt_double(&e->rng); }
int o_rng_randint(OEnv *e, int lo, int hi) { return mt_randint(&e->rng, lo, hi); }
void o_rng_permutation(OEnv *e, int n, int *out) { mt_permutation(&e->rng, n, out); }
void o_rng_get_state(const OEnv *e, uint32_t *key, int *pos) { memcpy(key, e->rng.key, sizeof(e->rng.key)); *pos = e->rng.pos; }
void o_rng_set_state(OEnv *e, const uint32_t *key, int pos) { memcpy(e->rng.key, key, sizeof(e->rng.key)); e->rng.pos = pos; }

int o_height(const OEnv *e) { return e->H; }
int o_width(const OEnv *e) { return e->W; }
void o_get_maze(const OEnv *e, uint8_t *out) { memcpy(out, e->maze, (size_t)(e->H * e->W)); }
void o_get_gen_maze(const OEnv *e, uint8_t *out) { memcpy(out, e->gen_maze, (size_t)(e->H * e->W)); }
void o_get_state(const OEnv *e, int *st /*4*/, int *goals /*4*/, int *counters /*2: c_far, elapsed*/) {
    memcpy(st, e->state, sizeof(e->state));
    memcpy(goals, e->goals, sizeof(e->goals));
    counters[0] = e->c_far; counters[1] = e->elapsed;
}
/* inject a full deterministic state: map (H*W uint8), positions, counters; target policy state is
 * left alone (use o_set_ram / o_set_nav). */
void o_set_state(OEnv *e, int H, int W, const uint8_t *maze, const int *st, int c_far, int elapsed) {
    e->H = H; e->W = W;
    memcpy(e->maze, maze, (size_t)(H * W));
    memcpy(e->gen_maze, maze, (size_t)(H * W));
    memcpy(e->state, st, sizeof(e->state));
    e->c_far = c_far; e->elapsed = elapsed;
    e->w_p = e->target_mode == O_T_PZR ? 1.0 : (e->target_mode == O_T_FAR ? -0.5 : 0.0);
}
void o_get_ram(const OEnv *e, int *plan /*16*/, int *len, int *idx) { memcpy(plan, e->ram_plan, sizeof(e->ram_plan)); *len = e->ram_len; *idx = e->ram_i; }
void o_set_ram(OEnv *e, const int *plan, int len, int idx) { memcpy(e->ram_plan, plan, sizeof(int) * (size_t)len); e->ram_len = len; e->ram_i = idx; }
int o_get_nav(const OEnv *e, int *plan, int maxlen, int *idx, int *goal) {
    int n = e->nav_len < maxlen ? e->nav_len : maxlen;
    memcpy(plan, e->nav_plan, sizeof(int) * (size_t)n);
    *idx = e->nav_i; goal[0] = e->nav_goal[0]; goal[1] = e->nav_goal[1];
    return e->nav_len;
}
void o_set_nav(OEnv *e, const int *plan, int len, int idx, const int *goal) {
    memcpy(e->nav_plan, plan, sizeof(int) * (size_t)len); e->nav_len = len; e->nav_i = idx;
    e->nav_goal[0] = goal[0]; e->nav_goal[1] = goal[1];
}
void o_get_nav_stats(const OEnv *e, int *out /*4*/) { out[0] = e->nav_replans; out[1] = e->nav_planb; out[2] = e->astar_expansions; out[3] = e->astar_replaces; }
/* standalone A* on the env's generator maze (tests) */
int o_astar(OEnv *e, const int *start, const int *goal, int *plan, int maxplan) { return astar(e, e->gen_maze, start, goal, plan, maxplan); }

/* ------------------------------------------------------------------------------------------- */
/* CPU baseline driver: `n_envs` independent envs stepped round-robin with LCG-random tracker /
 * target actions and reset-on-done, like random_agent_multi.py:34-51.  Returns env-steps done.  */
/* ------------------------------------------------------------------------------------------- */
long o_run_random(int map_type, int obs_type, int target_mode, int level, uint32_t seed, long n_steps, double *reward_sum, long *n_episodes) {
    OEnv *e = o_create(map_type, obs_type, target_mode, level);
    o_seed(e, seed);
    uint8_t *obs = (uint8_t *)malloc(2 * O_MAXCELLS);
    uint32_t lcg = seed * 2654435761u + 12345u;
    long steps = 0, eps = 0;
    double rs[2] = {0, 0}, rew[2];
    o_reset(e, obs);
    while (steps < n_steps) {
        int a[2];
        lcg = lcg * 1664525u + 1013904223u; a[0] = (int)((lcg >> 24) & 3);
        lcg = lcg * 1664525u + 1013904223u; a[1] = (int)((lcg >> 24) & 3);
        int done = o_step(e, a, obs, rew, 0, 0);
        rs[0] += rew[0]; rs[1] += rew[1];
        steps++;
        if (done) { eps++; o_reset(e, obs); }
    }
    if (reward_sum) { reward_sum[0] = rs[0]; reward_sum[1] = rs[1]; }
    if (n_episodes) *n_episodes = eps;
    free(obs);
    o_destroy(e);
    return steps;
}

/* ------------------------------------------------------------------------------------------- */
/* Batched replay for the full-size parity tests (pthreads: envs are independent).              */
/*                                                                                              */
/* o_batch_pipeline: env e of an E-wide batch behaves like the reference after                  */
/* np.random.seed(seed0 + e): reset, then T steps with reset-on-done (the obs of a finished env */
/* is replaced by its reset obs; rewards / done are the finishing step's -- what libtrack2d's   */
/* T2D_FLAG_AUTO_RESET does).  One call handles envs [first, first + n); arrays are indexed     */
/* [t][e] over the whole batch width E.  Any output pointer may be NULL.                        */
/* o_batch_inject_steps: env e starts from an INJECTED (maze[e], pos0[e], ctr0[e]) and takes T  */
/* steps without reset (no RNG involved for Adv / PZR / Far targets).                           */
/* ------------------------------------------------------------------------------------------- */
#include <pthread.h>

typedef struct {
    int map_type, obs_type, target_mode, level, T, H, W, inject;
    long E, first, n;
    uint32_t seed0;
    const int *actions;
    const uint8_t *maze0; const int *pos0, *ctr0;
    uint8_t *reset_obs, *reset_maze; int *reset_pos;
    int *pos, *ctr; double *rew; uint8_t *done; int *tgt_act; uint8_t *obs; uint32_t *rng_ckpt; int *plan_meta;
    int tid, nthreads;
    long n_done;
} BatchJob;

static void *batch_worker(void *arg) {
    BatchJob *j = (BatchJob *)arg;
    OEnv *e = o_create(j->map_type, j->obs_type, j->target_mode, j->level);
    const size_t ob = (size_t)2 * (size_t)o_obs_cells(e);
    uint8_t *scratch = (uint8_t *)malloc(2 * O_MAXCELLS);
    for (long i = j->first + j->tid; i < j->first + j->n; i += j->nthreads) {
        if (j->inject) {
            o_set_state(e, j->H, j->W, j->maze0 + (size_t)i * (size_t)(j->H * j->W), j->pos0 + i * 4, j->ctr0[i * 2], j->ctr0[i * 2 + 1]);
        } else {
            o_seed(e, j->seed0 + (uint32_t)i);
            o_reset(e, scratch);
            if (j->reset_obs) memcpy(j->reset_obs + (size_t)i * ob, scratch, ob);
            if (j->reset_maze) memcpy(j->reset_maze + (size_t)i * (size_t)(e->H * e->W), e->maze, (size_t)(e->H * e->W));
            if (j->reset_pos) memcpy(j->reset_pos + i * 4, e->state, sizeof(e->state));
        }
        for (int t = 0; t < j->T; t++) {
            const size_t k = (size_t)t * (size_t)j->E + (size_t)i;
            double r[2];
            int ta = 0;
            int d = o_step(e, j->actions + k * 2, scratch, r, 0, &ta);
            if (d && !j->inject) { o_reset(e, scratch); j->n_done++; }
            if (j->pos) memcpy(j->pos + k * 4, e->state, sizeof(e->state));
            if (j->ctr) { j->ctr[k * 2] = e->c_far; j->ctr[k * 2 + 1] = e->elapsed; }
            if (j->rew) { j->rew[k * 2] = r[0]; j->rew[k * 2 + 1] = r[1]; }
            if (j->done) j->done[k] = (uint8_t)d;
            if (j->tgt_act) j->tgt_act[k] = ta;
            if (j->obs) memcpy(j->obs + k * ob, scratch, ob);
            if (j->plan_meta) {
                int is_ram = e->target_mode == O_T_RAM;
                j->plan_meta[k * 2] = is_ram ? e->ram_i : e->nav_i;
                j->plan_meta[k * 2 + 1] = is_ram ? e->ram_len : e->nav_len;
            }
        }
        if (j->rng_ckpt) {
            memcpy(j->rng_ckpt + i * 9, e->rng.key, 8 * sizeof(uint32_t));
            j->rng_ckpt[i * 9 + 8] = (uint32_t)e->rng.pos;
        }
    }
    free(scratch);
    o_destroy(e);
    return 0;
}

static long batch_run(BatchJob *proto, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    BatchJob jobs[64];
    pthread_t th[64];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = *proto; jobs[t].tid = t; jobs[t].nthreads = nthreads; jobs[t].n_done = 0;
        pthread_create(&th[t], 0, batch_worker, &jobs[t]);
    }
    long nd = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], 0); nd += jobs[t].n_done; }
    return nd;
}

long o_batch_pipeline(int map_type, int obs_type, int target_mode, int level, long E, long first, long n, uint32_t seed0, int T,
                      const int *actions /*[T][E][2]*/, uint8_t *reset_obs /*[E][2][cells]*/, uint8_t *reset_maze /*[E][H][W]*/,
                      int *reset_pos /*[E][4]*/, int *pos /*[T][E][4]*/, int *ctr /*[T][E][2]*/, double *rew /*[T][E][2]*/,
                      uint8_t *done /*[T][E]*/, int *tgt_act /*[T][E]*/, uint8_t *obs /*[T][E][2][cells]*/, uint32_t *rng_ckpt /*[E][9]*/,
                      int *plan_meta /*[T][E][2] = (idx, len) of the Navigator / RamAgent plan after the step*/, int nthreads) {
    BatchJob j;
    memset(&j, 0, sizeof(j));
    j.map_type = map_type; j.obs_type = obs_type; j.target_mode = target_mode; j.level = level; j.T = T;
    j.E = E; j.first = first; j.n = n; j.seed0 = seed0; j.actions = actions;
    j.reset_obs = reset_obs; j.reset_maze = reset_maze; j.reset_pos = reset_pos;
    j.pos = pos; j.ctr = ctr; j.rew = rew; j.done = done; j.tgt_act = tgt_act; j.obs = obs; j.rng_ckpt = rng_ckpt; j.plan_meta = plan_meta;
    return batch_run(&j, nthreads);
}

void o_batch_inject_steps(int map_type, int obs_type, int target_mode, int level, long E, long first, long n, int H, int W, int T,
                          const uint8_t *maze /*[E][H][W]*/, const int *pos0 /*[E][4]*/, const int *ctr0 /*[E][2]*/,
                          const int *actions /*[T][E][2]*/, int *pos, int *ctr, double *rew, uint8_t *done, uint8_t *obs, int nthreads) {
    BatchJob j;
    memset(&j, 0, sizeof(j));
    j.map_type = map_type; j.obs_type = obs_type; j.target_mode = target_mode; j.level = level; j.T = T; j.H = H; j.W = W; j.inject = 1;
    j.E = E; j.first = first; j.n = n; j.actions = actions; j.maze0 = maze; j.pos0 = pos0; j.ctr0 = ctr0;
    j.pos = pos; j.ctr = ctr; j.rew = rew; j.done = done; j.obs = obs;
    batch_run(&j, nthreads);
}
