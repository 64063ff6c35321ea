// track2d_nav.cuh -- Navigator (envs/navigator.py:5-70) and its A* planner (envs/Astar_solver.py) on device.
//
// The action stream of a Nav/RPF target depends on WHICH shortest path A* returns, and that depends on
// CPython heapq's tie placement (push = _siftdown, pop = move-smaller-child-up-to-a-leaf then
// _siftdown) over keys [f, node] with f = g + float64 euclidean distance and Node.__lt__ = path_cost <
// (Astar_solver.py:30-32,53-63,127).  The planner below replays exactly those array operations, one
// lane per plan (the search is inherently sequential), with g / parent-action maps and the first
// T2D_HEAP_SMEM heap entries in shared memory and the rest of the heap spilled to an HBM workspace.
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
    double hf[T2D_HEAP_SMEM];
    uint32_t hc[T2D_HEAP_SMEM];    // cell | g << 16
};

struct HeapView {
    double *sf; uint32_t *sc; // shared part
    double *gf; uint32_t *gc; // HBM spill (indices >= T2D_HEAP_SMEM)
    __device__ __forceinline__ double f(int i) const { return i < T2D_HEAP_SMEM ? sf[i] : gf[i - T2D_HEAP_SMEM]; }
    __device__ __forceinline__ uint32_t c(int i) const { return i < T2D_HEAP_SMEM ? sc[i] : gc[i - T2D_HEAP_SMEM]; }
    __device__ __forceinline__ void set(int i, double fv, uint32_t cv) {
        if (i < T2D_HEAP_SMEM) { sf[i] = fv; sc[i] = cv; } else { gf[i - T2D_HEAP_SMEM] = fv; gc[i - T2D_HEAP_SMEM] = cv; }
    }
};

// [f, node] < [f2, node2]
__device__ __forceinline__ bool heap_lt(double fa, uint32_t ca, double fb, uint32_t cb) {
    return fa != fb ? fa < fb : (ca >> 16) < (cb >> 16);
}

__device__ __forceinline__ void hq_siftdown(HeapView &h, int startpos, int pos) { // heapq._siftdown
    double nf = h.f(pos);
    uint32_t nc = h.c(pos);
    while (pos > startpos) {
        int parent = (pos - 1) >> 1;
        double pf = h.f(parent);
        uint32_t pc = h.c(parent);
        if (heap_lt(nf, nc, pf, pc)) {
            h.set(pos, pf, pc);
            pos = parent;
            continue;
        }
        break;
    }
    h.set(pos, nf, nc);
}

__device__ __forceinline__ void hq_siftup(HeapView &h, int pos, int endpos) { // heapq._siftup
    int startpos = pos;
    double nf = h.f(pos);
    uint32_t nc = h.c(pos);
    int child = 2 * pos + 1;
    while (child < endpos) {
        int right = child + 1;
        double cf = h.f(child);
        uint32_t cc = h.c(child);
        if (right < endpos) {
            double rf = h.f(right);
            uint32_t rc = h.c(right);
            if (!heap_lt(cf, cc, rf, rc)) { child = right; cf = rf; cc = rc; }
        }
        h.set(pos, cf, cc);
        pos = child;
        child = 2 * pos + 1;
    }
    h.set(pos, nf, nc);
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
        h.sf = a.hf; h.sc = a.hc;
        h.gf = reinterpret_cast<double *>(w.astar_ws + (size_t)slot * T2D_MAX_CELLS * 12);
        h.gc = reinterpret_cast<uint32_t *>(w.astar_ws + (size_t)slot * T2D_MAX_CELLS * 12 + (size_t)T2D_MAX_CELLS * 8);
        const int goal = gr * W + gc;
        int hn = 0;
        {
            int start = sr * W + sc;
            double dr = (double)(sr - gr), dc = (double)(sc - gc);
            a.gcost[start] = 0;
            h.set(0, 0.0 + __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc))), (uint32_t)start);
            hn = 1;
        }
        int sol = -1;
        while (hn > 0) {
            // Frontier.pop == heapq.heappop
            hn--;
            double lf = h.f(hn);
            uint32_t lc = h.c(hn);
            uint32_t top = lc;
            if (hn > 0) {
                top = h.c(0);
                h.set(0, lf, lc);
                hq_siftup(h, 0, hn);
            }
            int cell = top & 0xFFFFu, g = top >> 16;
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
                    double dr = (double)(nr - gr), dc = (double)(nc - gc);
                    double f = __dadd_rn((double)(g + 1), __dsqrt_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc))));
                    h.set(hn, f, (uint32_t)child | ((uint32_t)(g + 1) << 16)); // Frontier.add == heappush
                    hn++;
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
