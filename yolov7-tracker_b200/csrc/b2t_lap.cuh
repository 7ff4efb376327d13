// b2t_lap.cuh -- exact thresholded linear assignment, one CTA per problem.
//
// Replaces ``lap.lapjv(cost, extend_cost=True, cost_limit=t)`` as called from
// tracker/matching.py:30-41.  lap solves the (N+M)^2 cost-limit-extended square problem
// (oracle/lapjv.py), whose optimum equals
//        minimise  sum_matched (c_ij - t)   over partial matchings,
// so only entries c_ij < t can ever be matched.  With IoU costs that graph is ~99 % empty and
// falls apart into many small connected components, hence:
//   1. the caller hands over the sub-threshold entries as CSR rows (col, cost);
//   2. connected components by min-label propagation over the CSR (shared-memory atomics);
//   3. a stable counting sort groups rows by component;
//   4. warps pull components off a queue and run Jonker-Volgenant-style shortest augmenting
//      paths with dual potentials (u, v) -- one warp per component, lanes over the row's
//      edges for the relaxation and over the frontier for the arg-min.
// Rectangular formulation: each row owns a private "stay unmatched" column of cost t/2 (its
// dual never moves, so it is never stored) and real edges are shifted by -t/2 -- every row is
// assigned exactly once (to a real column or to its own dummy), N augmentations in total.
// The result is the unique optimum whenever that is unique (ties are solver-dependent in lap
// as well; the generators are tie-free).
#pragma once
#include "b2t_prims.cuh"

namespace b2t {

template <class T> struct LapCsr {
    const int* row_start;   // nullptr -> row i starts at i * row_stride
    int row_stride;
    const int* row_cnt;
    const int* e_col;
    const T* e_cost;
    B2T_DEV int start(int i) const { return row_start ? row_start[i] : i * row_stride; }
};

template <class T> struct LapWork {
    T *u, *v, *dist;
    int *x, *y, *pred, *lab_r, *lab_c, *tl, *rows_sorted, *comp_root, *row_off, *tl_off, *cur;
    unsigned char* sc;
    int* scratch;   // 64 ints
    template <class A> B2T_DEV void carve(A& a, int nmax, int mmax) {
        u = a.template take<T>(nmax); v = a.template take<T>(mmax); dist = a.template take<T>(mmax);
        x = a.template take<int>(nmax); y = a.template take<int>(mmax); pred = a.template take<int>(mmax);
        lab_r = a.template take<int>(nmax); lab_c = a.template take<int>(mmax); tl = a.template take<int>(mmax);
        rows_sorted = a.template take<int>(nmax); comp_root = a.template take<int>(nmax);
        row_off = a.template take<int>(nmax + 1); tl_off = a.template take<int>(nmax + 1); cur = a.template take<int>(nmax);
        sc = a.template take<unsigned char>(mmax); scratch = a.template take<int>(64);
    }
    static void size(ArenaSize& a, int nmax, int mmax) {
        a.take<T>(nmax); a.take<T>(mmax); a.take<T>(mmax);
        a.take<int>(nmax); a.take<int>(mmax); a.take<int>(mmax);
        a.take<int>(nmax); a.take<int>(mmax); a.take<int>(mmax);
        a.take<int>(nmax); a.take<int>(nmax);
        a.take<int>(nmax + 1); a.take<int>(nmax + 1); a.take<int>(nmax);
        a.take<unsigned char>(mmax); a.take<int>(64);
    }
};

#define B2T_LAP_BIG ((T)1e30)
#define B2T_LAP_NOLAB 0x3fffffff

// One warp, one connected component: rows[0..nrows) ascending, tl = private frontier storage.
template <class T>
B2T_DEV void lap_component(const LapCsr<T>& g, const T half_t, LapWork<T>& w, const int* rows, int nrows, int* tl) {
    const int lane = lane_id();
    const unsigned lt = lanemask_lt();
    for (int ri = 0; ri < nrows; ++ri) {
        const int r0 = rows[ri];
        int nt = 0;
        T minval = (T)0;
        int i = r0;
        T best_dummy = half_t - w.u[r0];
        int dummy_row = r0;
        int sink = -1;   // -1: dummy of dummy_row, >= 0: free real column
        for (;;) {
            const T ui = w.u[i];
            const int es = g.start(i), ec = g.row_cnt[i];
            for (int e0 = 0; e0 < ec; e0 += 32) {
                const int e = e0 + lane;
                bool act = e < ec, fresh = false;
                int j = -1;
                if (act) { j = g.e_col[es + e]; if (w.sc[j]) act = false; }
                if (act) {
                    const T red = minval + (((g.e_cost[es + e] - half_t) - ui) - w.v[j]);
                    const T dj = w.dist[j];
                    fresh = dj >= B2T_LAP_BIG;
                    if (red < dj) { w.dist[j] = red; w.pred[j] = i; }
                }
                const unsigned fm = __ballot_sync(B2T_FULL, fresh);
                if (fresh) tl[nt + __popc(fm & lt)] = j;
                nt += __popc(fm);
            }
            __syncwarp();
            T bv = B2T_LAP_BIG;
            int bj = -1;
            for (int k = lane; k < nt; k += 32) {
                const int j = tl[k];
                if (!w.sc[j]) { const T d = w.dist[j]; if (d < bv) { bv = d; bj = j; } }
            }
            for (int o = 16; o; o >>= 1) {
                const T ov = shfl_xor(bv, o);
                const int oj = shfl_xor(bj, o);
                if (oj >= 0 && (bj < 0 || ov < bv || (ov == bv && oj < bj))) { bv = ov; bj = oj; }
            }
            if (bj < 0 || best_dummy <= bv) { sink = -1; minval = best_dummy; break; }
            minval = bv;
            if (lane == 0) w.sc[bj] = 1;
            __syncwarp();
            const int yi = w.y[bj];
            if (yi < 0) { sink = bj; break; }
            i = yi;
            const T dd = minval + (half_t - w.u[i]);
            if (dd < best_dummy) { best_dummy = dd; dummy_row = i; }
        }
        // dual update (pre-augmentation y)
        for (int k = lane; k < nt; k += 32) {
            const int j = tl[k];
            if (w.sc[j]) {
                const T d = minval - w.dist[j];
                w.v[j] = w.v[j] - d;
                const int yi = w.y[j];
                if (yi >= 0) w.u[yi] = w.u[yi] + d;
            }
        }
        if (lane == 0) w.u[r0] = w.u[r0] + minval;
        __syncwarp();
        if (lane == 0) {
            if (sink < 0) {
                int ii = dummy_row;
                int jprev = w.x[ii];
                w.x[ii] = -1;
                while (ii != r0) {
                    const int j = jprev;
                    ii = w.pred[j];
                    w.y[j] = ii;
                    jprev = w.x[ii];
                    w.x[ii] = j;
                }
            } else {
                int j = sink;
                for (;;) {
                    const int ii = w.pred[j];
                    w.y[j] = ii;
                    const int jn = w.x[ii];
                    w.x[ii] = j;
                    j = jn;
                    if (ii == r0) break;
                }
            }
        }
        __syncwarp();
        for (int k = lane; k < nt; k += 32) { const int j = tl[k]; w.dist[j] = B2T_LAP_BIG; w.sc[j] = 0; }
        __syncwarp();
    }
}

// All threads of the CTA.  On return w.x[0..n) / w.y[0..m) hold the assignment (-1 = unmatched).
template <class T>
B2T_DEV void lap_solve_cta(int n, int m, const LapCsr<T>& g, T thresh, LapWork<T>& w) {
    const int tid = (int)threadIdx.x, nthr = (int)blockDim.x;
    const int lane = lane_id(), wid = warp_id(), nw = num_warps();
    const T half_t = thresh / (T)2;
    for (int i = tid; i < n; i += nthr) {
        w.u[i] = (T)0; w.x[i] = -1; w.lab_r[i] = g.row_cnt[i] > 0 ? i : -1;
        w.row_off[i] = 0; w.tl_off[i] = 0;
    }
    for (int j = tid; j < m; j += nthr) { w.v[j] = (T)0; w.y[j] = -1; w.dist[j] = B2T_LAP_BIG; w.sc[j] = 0; w.lab_c[j] = B2T_LAP_NOLAB; }
    if (tid == 0) { w.row_off[n] = 0; w.tl_off[n] = 0; }
    __syncthreads();
    if (n == 0 || m == 0) return;
    // ---- connected components: min-label propagation
    for (;;) {
        if (tid == 0) w.scratch[40] = 0;
        __syncthreads();
        for (int i = wid; i < n; i += nw) {
            const int ec = g.row_cnt[i];
            if (ec == 0) continue;
            const int es = g.start(i);
            int l = w.lab_r[i];
            for (int e = lane; e < ec; e += 32) { const int lc = w.lab_c[g.e_col[es + e]]; if (lc < l) l = lc; }
            for (int o = 16; o; o >>= 1) { const int ol = shfl_xor(l, o); if (ol < l) l = ol; }
            if (lane == 0 && l < w.lab_r[i]) { w.lab_r[i] = l; w.scratch[40] = 1; }
            for (int e = lane; e < ec; e += 32) {
                const int old = atomicMin(&w.lab_c[g.e_col[es + e]], l);
                if (old > l) w.scratch[40] = 1;
            }
        }
        __syncthreads();
        const int changed = w.scratch[40];
        __syncthreads();
        if (!changed) break;
    }
    // ---- component sizes (rows, columns) -> offsets
    for (int i = tid; i < n; i += nthr) if (w.lab_r[i] >= 0) atomicAdd(&w.row_off[w.lab_r[i]], 1);
    for (int j = tid; j < m; j += nthr) if (w.lab_c[j] != B2T_LAP_NOLAB) atomicAdd(&w.tl_off[w.lab_c[j]], 1);
    __syncthreads();
    block_exscan(w.row_off, n + 1, w.scratch);
    block_exscan(w.tl_off, n + 1, w.scratch);
    const int ncomp = block_compact(n, [&](int i) { return w.lab_r[i] == i; }, w.comp_root, w.scratch);
    for (int i = tid; i < n; i += nthr) w.cur[i] = w.row_off[i];
    if (tid == 0) w.scratch[41] = 0;
    __syncthreads();
    // ---- stable counting sort of the rows by component label (one warp, n/32 steps)
    if (wid == 0) {
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int i = c0 + lane;
            const int lab = i < n ? w.lab_r[i] : -1;
            const unsigned mm = __match_any_sync(B2T_FULL, lab);
            if (lab >= 0) w.rows_sorted[w.cur[lab] + __popc(mm & lanemask_lt())] = i;
            __syncwarp();
            if (lab >= 0 && (mm & lanemask_lt()) == 0) w.cur[lab] += __popc(mm);
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- solve: warps pull components off the queue
    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&w.scratch[41], 1);
        c = shfl(c, 0);
        if (c >= ncomp) break;
        const int root = w.comp_root[c];
        const int r_begin = w.row_off[root], r_end = w.row_off[root + 1];
        lap_component<T>(g, half_t, w, w.rows_sorted + r_begin, r_end - r_begin, w.tl + w.tl_off[root]);
    }
    __syncthreads();
}

}  // namespace b2t
