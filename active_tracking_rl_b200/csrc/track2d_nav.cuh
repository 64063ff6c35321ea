// track2d_nav.cuh -- Navigator (envs/navigator.py:5-70) and its A* planner (envs/Astar_solver.py) on device.
//
// The action stream of a Nav/RPF target depends on WHICH shortest path A* returns, and that depends on
// CPython heapq's tie placement (push = _siftdown, pop = move-smaller-child-up-to-a-leaf then
// _siftdown) over keys [f, node] with f = g + float64 euclidean distance and Node.__lt__ = path_cost <
// (Astar_solver.py:30-32,53-63,127).  The planner below replays exactly those array operations, one
// warp per plan (the heap replay is inherently sequential; the neighbour expansion is not), with g / parent-action maps and
// the first T2D_HEAP_SMEM heap entries in shared memory and the rest of the heap spilled to an HBM workspace.
//
// Frontier.replace (Astar_solver.py:65-73,146-147) can never fire on a unit-cost 4-connected grid with
// the euclidean heuristic (adjacent cells have path costs of opposite parity, and a node that could
// trigger it would have to be popped after the node it wants to replace); the condition is still
// tested and reported through T2D_STATUS_ASTAR_REPLACE rather than silently ignored.
#pragma once
#include "track2d_common.cuh"

#define T2D_HEAP_SMEM 320
#define T2D_MAX_CELLS (82 * 82)

struct __align__(16) HeapEntry {  // one 16-byte shared-memory access per heap slot
    double f;
    uint32_t c;   // cell | g << 16
    uint32_t rc;  // row << 8 | col of the cell (saves the division by the map width on every pop)
};

struct AStarScratch {
    uint16_t gcost[T2D_MAX_CELLS]; // 0xFFFF = never seen; bit 15 = explored; bits 13-14 = action that led into the cell; low 13 bits = path cost
    HeapEntry heap[T2D_HEAP_SMEM]; // the first T2D_HEAP_SMEM slots of the frontier (frontiers of real maps stay below ~250)
    // 19.7 KB per plan: eleven one-warp planners per SM (the planner is latency-bound on one lane: throughput = plans in flight)
};

// FAST: every index touched is < T2D_HEAP_SMEM (checked once per heap operation by the caller), so the accesses are plain
// LDS.128 / STS.128 with no per-access branch; otherwise slots >= T2D_HEAP_SMEM live in the HBM spill area
template <bool FAST>
struct HeapView {
    HeapEntry *s, *g;
    __device__ __forceinline__ HeapEntry get(int i) const {
        if (FAST) return s[i];
        return i < T2D_HEAP_SMEM ? s[i] : g[i - T2D_HEAP_SMEM];
    }
    __device__ __forceinline__ void set(int i, const HeapEntry &v) const {
        if (FAST) s[i] = v;
        else if (i < T2D_HEAP_SMEM) s[i] = v;
        else g[i - T2D_HEAP_SMEM] = v;
    }
};

// [f, node] < [f2, node2].  f = g + distance is a non-negative finite double, so its bit pattern orders like the value: integer
// compares instead of two FP64 predicate instructions on the critical path of every sift level.
__device__ __forceinline__ bool heap_lt(const HeapEntry &a, const HeapEntry &b) {
    const long long fa = __double_as_longlong(a.f), fb = __double_as_longlong(b.f);
    return fa != fb ? fa < fb : (a.c >> 16) < (b.c >> 16);
}

template <bool FAST>
__device__ __forceinline__ void hq_siftdown(const HeapView<FAST> &h, int startpos, int pos, const HeapEntry &newitem) { // heapq._siftdown
    while (pos > startpos) {
        const int parent = (pos - 1) >> 1;
        const HeapEntry pe = h.get(parent);
        if (heap_lt(newitem, pe)) {
            h.set(pos, pe);
            pos = parent;
            continue;
        }
        break;
    }
    h.set(pos, newitem);
}

template <bool FAST>
__device__ __forceinline__ void hq_siftup(const HeapView<FAST> &h, int pos, int endpos, const HeapEntry &newitem) { // heapq._siftup
    const int startpos = pos;
    int child = 2 * pos + 1;
    while (child < endpos) {
        const int right = child + 1;
        HeapEntry ce = h.get(child);
        if (right < endpos) {
            const HeapEntry re = h.get(right);  // adjacent slot: the two loads are in flight together
            if (!heap_lt(ce, re)) { child = right; ce = re; }
        }
        h.set(pos, ce);
        pos = child;
        child = 2 * pos + 1;
    }
    hq_siftdown(h, startpos, pos, newitem);
}

__device__ __forceinline__ uint32_t rpf_corner(int H, int W, int q) { // generators.py:12-19 candidate_goals
    int a0 = (int)(H / 6.0), a1 = (int)(H * 5 / 6.0), b0 = (int)(W / 6.0), b1 = (int)(W * 5 / 6.0);
    int r = (q == 0 || q == 3) ? a0 : a1;
    int c = (q == 0 || q == 1) ? b0 : b1;
    return (uint32_t)r | ((uint32_t)c << 8);
}

// Astar_solver.py:121-149 on the bit grid `bm` (generator maze), one warp per plan.  The frontier is CPython's heapq replayed
// array operation by array operation on lane 0 (that fixes WHICH shortest path comes out); lanes 0..3 expand the four
// neighbours of the popped node in parallel (wall bit, seen / explored lookup, f = g + float64 euclidean distance), and lane 0
// pushes the new ones in action order.  Writes the action list (get_actions, :102-110) 2 bits per action into `plan` (a buffer of
// TRACK2D_NAV_MAXPLAN 2-bit slots, used as a ring) starting at slot `base`, and returns its length, or -1 when the goal is
// unreachable.  base < 0: the buffer is cleared first and the plan starts at slot 0 (the reference's "new plan replaces the old one").
// start/goal are (row, col) map coordinates.
__device__ __forceinline__ void plan_put(uint8_t *plan, int slot, int act) {
    slot &= TRACK2D_NAV_MAXPLAN - 1;
    const int sh = 2 * (slot & 3);
    plan[slot >> 2] = (uint8_t)((plan[slot >> 2] & ~(3 << sh)) | (act << sh));
}
__device__ __forceinline__ int plan_get(const uint8_t *plan, int slot) {
    slot &= TRACK2D_NAV_MAXPLAN - 1;
    return (plan[slot >> 2] >> (2 * (slot & 3))) & 3;
}
__device__ int astar_plan(const World &w, uint8_t *plan, int base, const uint32_t *bm, AStarScratch &a, int slot, int sr, int sc, int gr, int gc, int lane) {
    const int W = w.W, cells = w.H * w.W;
    for (int i = lane; i < cells; i += 32) a.gcost[i] = 0xFFFFu;
    if (base < 0) {
        for (int i = lane; i < T2D_NAV_PLAN_BYTES; i += 32) plan[i] = 0;
        base = 0;
    }
    __syncwarp();
    HeapEntry *spill = reinterpret_cast<HeapEntry *>(w.astar_ws + (size_t)slot * T2D_MAX_CELLS * sizeof(HeapEntry));
    const HeapView<true> hfast{a.heap, spill};
    const HeapView<false> hslow{a.heap, spill};
    const int goal = gr * W + gc;
    int hn = 0;  // heap size (lane 0)
    if (lane == 0) {
        const int start = sr * W + sc;
        const double dr = (double)(sr - gr), dc = (double)(sc - gc);
        a.gcost[start] = 0;
        HeapEntry en;
        en.f = 0.0 + __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc)));
        en.c = (uint32_t)start;
        en.rc = ((uint32_t)sr << 8) | (uint32_t)sc;
        a.heap[0] = en;
        hn = 1;
    }
    int sol = -1;
    for (;;) {
        uint32_t top = 0xFFFFFFFFu, top_rc = 0;  // heap empty
        if (lane == 0 && hn > 0) {
            // Frontier.pop == heapq.heappop
            hn--;
            const bool fast = hn < T2D_HEAP_SMEM;
            const HeapEntry last = fast ? hfast.get(hn) : hslow.get(hn);
            top = last.c;
            top_rc = last.rc;
            if (hn > 0) {
                const HeapEntry first = a.heap[0];
                top = first.c;
                top_rc = first.rc;
                if (fast) hq_siftup(hfast, 0, hn, last);
                else hq_siftup(hslow, 0, hn, last);
            }
        }
        top = __shfl_sync(0xFFFFFFFFu, top, 0);
        top_rc = __shfl_sync(0xFFFFFFFFu, top_rc, 0);
        if (top == 0xFFFFFFFFu) break;
        const int cell = (int)(top & 0xFFFFu), g = (int)(top >> 16);
        const int r = (int)(top_rc >> 8), c = (int)(top_rc & 255u);
        if (cell == goal) { sol = cell; break; }
        // neighbour `lane` (actions 0..3 in the reference's order)
        bool add = false;
        double f = 0.0;
        if (lane < 4) {
            if (lane == 0) a.gcost[cell] |= 0x8000u;  // explored.add (never one of its own neighbours)
            const int nr = r + action_dr(lane), nc = c + action_dc(lane);
            if (!((bm[map_word_index(nr, nc)] >> ((nc + T2D_PAD) & 31)) & 1u)) {  // wall: child == parent, already explored
                const int child = nr * W + nc;
                const uint32_t gv = a.gcost[child];
                if (gv == 0xFFFFu) {
                    a.gcost[child] = (uint16_t)((g + 1) | (lane << 13));
                    const double dr = (double)(nr - gr), dc = (double)(nc - gc);
                    f = __dadd_rn((double)(g + 1), __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc))));
                    add = true;
                } else if (!(gv & 0x8000u) && (int)(gv & 0x1FFFu) < g + 1) {
                    atomicOr(w.status, (uint32_t)T2D_STATUS_ASTAR_REPLACE);
                }
            }
        }
        const uint32_t addmask = __ballot_sync(0xFFFFFFFFu, add) & 0xFu;
        double fa[4];
#pragma unroll
        for (int act = 0; act < 4; act++) fa[act] = __shfl_sync(0xFFFFFFFFu, f, act);
        if (lane == 0 && addmask) {
#pragma unroll
            for (int act = 0; act < 4; act++) {
                if (!((addmask >> act) & 1u)) continue;
                HeapEntry en;  // Frontier.add == heappush
                en.f = fa[act];
                en.c = (uint32_t)((r + action_dr(act)) * W + c + action_dc(act)) | ((uint32_t)(g + 1) << 16);
                en.rc = ((uint32_t)(r + action_dr(act)) << 8) | (uint32_t)(c + action_dc(act));
                hn++;
                if (hn <= T2D_HEAP_SMEM) hq_siftdown(hfast, 0, hn - 1, en);
                else hq_siftdown(hslow, 0, hn - 1, en);
            }
        }
        __syncwarp();  // lane 0's heap / the neighbours' gcost stores are ordered before the next round
    }
    int len = -1;
    if (lane == 0 && sol >= 0) {
        len = a.gcost[sol] & 0x1FFF;
        if (len > TRACK2D_NAV_MAXPLAN) atomicOr(w.status, (uint32_t)T2D_STATUS_PLAN_OVERFLOW);
        int cell = sol;
        for (int i = len - 1; i >= 0; i--) {
            int act = (a.gcost[cell] >> 13) & 3;
            if (i < TRACK2D_NAV_MAXPLAN) plan_put(plan, base + i, act);
            cell -= action_dr(act) * W + action_dc(act);
        }
        if (len > TRACK2D_NAV_MAXPLAN) len = TRACK2D_NAV_MAXPLAN;
    }
    len = __shfl_sync(0xFFFFFFFFu, len, 0);
    __syncwarp();
    return len;
}

__device__ __forceinline__ void nav_store_planb(uint8_t *plan, int base, const uint32_t acts[10]) {
    if (base < 0) base = 0;
    for (int i = 0; i < 10; i++) plan_put(plan, base + i, (int)acts[i]);
}
