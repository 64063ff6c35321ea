// track2d_nav.cuh -- Navigator (envs/navigator.py:5-70) and its A* planner (envs/Astar_solver.py) on device.
//
// The action stream of a Nav/RPF target depends on WHICH shortest path A* returns, and that depends on
// CPython heapq's tie placement (push = _siftdown, pop = move-smaller-child-up-to-a-leaf then
// _siftdown) over keys [f, node] with f = g + float64 euclidean distance and Node.__lt__ = path_cost <
// (Astar_solver.py:30-32,53-63,127).  The planner below replays exactly those array operations, one
// lane per plan (the search is inherently sequential), with g / parent-action maps and the first
// T2D_HEAP_SMEM heap entries in shared memory and the rest of the heap spilled to an HBM workspace.  Keys are compared
// exactly in integer arithmetic (f_cmp), so the search needs no float64 at all.
//
// Frontier.replace (Astar_solver.py:65-73,146-147) can never fire on a unit-cost 4-connected grid with
// the euclidean heuristic (adjacent cells have path costs of opposite parity, and a node that could
// trigger it would have to be popped after the node it wants to replace); the condition is still
// tested and reported through T2D_STATUS_ASTAR_REPLACE rather than silently ignored.
#pragma once
#include "track2d_common.cuh"

#define T2D_HEAP_SMEM 1024
#define T2D_MAX_CELLS (82 * 82)

struct AStarScratch {
    uint16_t gcost[T2D_MAX_CELLS]; // 0xFFFF = never seen; bit 15 = explored; low 15 bits = path cost
    uint8_t pact[T2D_MAX_CELLS];   // action that led into the cell
    unsigned long long hk[T2D_HEAP_SMEM]; // heap entries: cell | g << 16 | d2 << 32  (f = g + sqrt(d2))
};

struct HeapView {
    unsigned long long *sk; // shared part
    unsigned long long *gk; // HBM spill (indices >= T2D_HEAP_SMEM)
    __device__ __forceinline__ unsigned long long get(int i) const { return i < T2D_HEAP_SMEM ? sk[i] : gk[i - T2D_HEAP_SMEM]; }
    __device__ __forceinline__ void set(int i, unsigned long long v) {
        if (i < T2D_HEAP_SMEM) sk[i] = v; else gk[i - T2D_HEAP_SMEM] = v;
    }
};

// Exact comparison of f = g + sqrt(d2) without floating point.  The reference compares float64 values
// g + np.linalg.norm(...) (Astar_solver.py:127,151-153).  For integers g <= 2^15, d2 <= 2 * 81^2 two such values are
// either exactly equal (same g and d2, or perfect squares with equal sums -- float64 represents those exactly) or at
// least ~4.5e-8 apart (|N - 2k sqrt(b)| >= 1 / (N + 2k sqrt(b)) for a non-square b), nine orders of magnitude more than
// float64 rounding can move them: the integer test below therefore orders and ties exactly like the reference's
// floats, and the heap needs neither a square root nor 8-byte float keys.
// Returns -1 / 0 / +1 for  ga + sqrt(da)  <, ==, >  gb + sqrt(db).
__device__ __forceinline__ int f_cmp(int ga, int da, int gb, int db) {
    int k = gb - ga; // sign(sqrt(da) - sqrt(db) - k)
    if (k == 0) return da < db ? -1 : (da > db ? 1 : 0);
    bool flip = k < 0;
    if (flip) { k = -k; int t = da; da = db; db = t; } // now: sign(sqrt(da) - sqrt(db) - k), k > 0, result negated if flipped
    int r;
    if (da <= db) {
        r = -1;
    } else {
        long long L = (long long)da - db - (long long)k * k; // sqrt(da) - sqrt(db) < k  <=>  L < 2 k sqrt(db)
        if (L < 0) {
            r = -1;
        } else {
            long long lhs = L * L, rhs = 4ll * k * k * db;
            r = lhs < rhs ? -1 : (lhs > rhs ? 1 : 0);
        }
    }
    return flip ? -r : r;
}

// [f, node] < [f2, node2]: floats first; on equal f Node.__lt__ (path_cost <), Astar_solver.py:30-32,55
__device__ __forceinline__ bool heap_lt(unsigned long long a, unsigned long long b) {
    int ga = (int)((a >> 16) & 0xFFFFu), gb = (int)((b >> 16) & 0xFFFFu);
    int c = f_cmp(ga, (int)(a >> 32), gb, (int)(b >> 32));
    return c != 0 ? c < 0 : ga < gb;
}

__device__ __forceinline__ void hq_siftdown(HeapView &h, int startpos, int pos) { // heapq._siftdown
    const unsigned long long nv = h.get(pos);
    while (pos > startpos) {
        int parent = (pos - 1) >> 1;
        unsigned long long pv = h.get(parent);
        if (heap_lt(nv, pv)) {
            h.set(pos, pv);
            pos = parent;
            continue;
        }
        break;
    }
    h.set(pos, nv);
}

__device__ __forceinline__ void hq_siftup(HeapView &h, int pos, int endpos) { // heapq._siftup
    int startpos = pos;
    const unsigned long long nv = h.get(pos);
    int child = 2 * pos + 1;
    while (child < endpos) {
        int right = child + 1;
        unsigned long long cv = h.get(child);
        if (right < endpos) {
            unsigned long long rv = h.get(right);
            if (!heap_lt(cv, rv)) { child = right; cv = rv; }
        }
        h.set(pos, cv);
        pos = child;
        child = 2 * pos + 1;
    }
    h.set(pos, nv);
    hq_siftdown(h, startpos, pos);
}

__device__ __forceinline__ uint32_t rpf_corner(int H, int W, int q) { // generators.py:12-19 candidate_goals
    int a0 = (int)(H / 6.0), a1 = (int)(H * 5 / 6.0), b0 = (int)(W / 6.0), b1 = (int)(W * 5 / 6.0);
    int r = (q == 0 || q == 3) ? a0 : a1;
    int c = (q == 0 || q == 1) ? b0 : b1;
    return (uint32_t)r | ((uint32_t)c << 8);
}

// Astar_solver.py:121-149 on the bit grid `bm` (generator maze).  Whole warp enters; lane 0 searches.
// Writes the action list (get_actions, :102-110) 2 bits per action into w.nav_plan[e] and returns its
// length, or -1 when the goal is unreachable.  start/goal are (row, col) map coordinates.
__device__ int astar_plan(const World &w, int e, const uint32_t *bm, AStarScratch &a, int slot, int sr, int sc, int gr, int gc, int lane) {
    const int W = w.W, cells = w.H * w.W;
    for (int i = lane; i < cells; i += 32) a.gcost[i] = 0xFFFFu;
    uint8_t *plan = w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES;
    for (int i = lane; i < T2D_NAV_PLAN_BYTES; i += 32) plan[i] = 0;
    __syncwarp();
    int len = -1;
    if (lane == 0) {
        HeapView h;
        h.sk = a.hk;
        h.gk = reinterpret_cast<unsigned long long *>(w.astar_ws + (size_t)slot * T2D_MAX_CELLS * 12);
        const int goal = gr * W + gc;
        int hn = 0;
        {
            int start = sr * W + sc;
            int dr = sr - gr, dc = sc - gc;
            a.gcost[start] = 0;
            h.set(0, (unsigned long long)start | ((unsigned long long)(dr * dr + dc * dc) << 32));
            hn = 1;
        }
        int sol = -1;
        while (hn > 0) {
            // Frontier.pop == heapq.heappop
            hn--;
            const unsigned long long last = h.get(hn);
            unsigned long long top = last;
            if (hn > 0) {
                top = h.get(0);
                h.set(0, last);
                hq_siftup(h, 0, hn);
            }
            int cell = (int)(top & 0xFFFFu), g = (int)((top >> 16) & 0xFFFFu);
            if (cell == goal) { sol = cell; break; }
            a.gcost[cell] |= 0x8000u; // explored.add
            int r = cell / W, c = cell - r * W;
#pragma unroll
            for (int act = 0; act < 4; act++) {
                int nr = r + action_dr(act), nc = c + action_dc(act);
                if ((bm[map_word_index(nr, nc)] >> ((nc + T2D_PAD) & 31)) & 1u) continue; // wall: child == parent, explored
                int child = nr * W + nc;
                uint32_t gv = a.gcost[child];
                if (gv == 0xFFFFu) {
                    a.gcost[child] = (uint16_t)(g + 1);
                    a.pact[child] = (uint8_t)act;
                    int dr = nr - gr, dc = nc - gc;
                    h.set(hn, (unsigned long long)child | ((unsigned long long)(g + 1) << 16) | ((unsigned long long)(dr * dr + dc * dc) << 32));
                    hn++; // Frontier.add == heappush
                    hq_siftdown(h, 0, hn - 1);
                } else if (!(gv & 0x8000u) && (int)(gv & 0x7FFFu) < g + 1) {
                    atomicOr(w.status, (uint32_t)T2D_STATUS_ASTAR_REPLACE);
                }
            }
        }
        if (sol >= 0) {
            len = a.gcost[sol] & 0x7FFF;
            if (len > TRACK2D_NAV_MAXPLAN) atomicOr(w.status, (uint32_t)T2D_STATUS_PLAN_OVERFLOW);
            int cell = sol;
            for (int i = len - 1; i >= 0; i--) {
                int act = a.pact[cell];
                if (i < TRACK2D_NAV_MAXPLAN) plan[i >> 2] |= (uint8_t)(act << (2 * (i & 3)));
                cell -= action_dr(act) * W + action_dc(act);
            }
            if (len > TRACK2D_NAV_MAXPLAN) len = TRACK2D_NAV_MAXPLAN;
        }
    }
    len = __shfl_sync(0xFFFFFFFFu, len, 0);
    __syncwarp();
    return len;
}

__device__ __forceinline__ void nav_store_planb(const World &w, int e, const uint32_t acts[10]) {
    uint8_t *plan = w.nav_plan + (size_t)e * T2D_NAV_PLAN_BYTES;
    plan[0] = (uint8_t)(acts[0] | (acts[1] << 2) | (acts[2] << 4) | (acts[3] << 6));
    plan[1] = (uint8_t)(acts[4] | (acts[5] << 2) | (acts[6] << 4) | (acts[7] << 6));
    plan[2] = (uint8_t)(acts[8] | (acts[9] << 2));
}
