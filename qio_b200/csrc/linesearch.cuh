// CTA-wide angle line search for ONE pair (reference qio/grad/jacobi.py:117-158), shared by the
// batched all-pairs kernel (pairs.cu) and the sequential Jacobi sweep (sweep.cu).
//
// One thread per candidate angle.  The reference's sequential scans are reproduced exactly:
//   coarse: running minimum with strict '>'  == first index of the global minimum, if it is
//           strictly below cost(0); otherwise t_opt falls back to 0.01 == coarse point 0;
//   fine:   accept k iff cost > cost_k + 1e-8, then cost = cost_k  (inherently sequential; one
//           thread walks the <= 201 values that the CTA evaluated in parallel).
#pragma once
#include "pair_math.cuh"

namespace qio {

constexpr int LS_THREADS = 320;           // >= 314 coarse points, 10 warps
constexpr int LS_MAX_FINE = 256;          // capacity of the fine-value buffer

struct ThetaTablesDev {                   // device copy of qio_b200_theta_tables (pointers are device pointers)
    int n_coarse, fine_stride;
    const double *coarse_theta, *coarse_cs;
    const int32_t *fine_len;
    const double *fine_theta, *fine_cs;
};

static inline ThetaTablesDev to_dev(const qio_b200_theta_tables &t) {
    ThetaTablesDev d;
    d.n_coarse = t.n_coarse;
    d.fine_stride = t.fine_stride;
    d.coarse_theta = t.coarse_theta;
    d.coarse_cs = t.coarse_cs;
    d.fine_len = t.fine_len;
    d.fine_theta = t.fine_theta;
    d.fine_cs = t.fine_cs;
    return d;
}

struct LineSearchSmem {
    double fine_val[LS_MAX_FINE];
    MinIdx warp_best[LS_THREADS / 32];
    MinIdx best;
};

// All LS_THREADS threads of the CTA must call this with the same arguments.  The result is
// returned in every thread (res is filled identically) -- thread 0's scan is broadcast via smem.
__device__ __forceinline__ void cta_linesearch(const PairBlock &b, bool i_in, bool j_in, const ThetaTablesDev &tab,
                                               LineSearchSmem &sm, qio_b200_ls_result &res) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int flags = 0;

    // cost at theta = 0 with cos = 1, sin = 0 exactly as np.cos(0), np.sin(0)
    const double cost0 = pair_cost(b, 1.0, 0.0, i_in, j_in, flags);

    // ---- coarse grid ---------------------------------------------------------------------
    MinIdx mine;
    mine.v = INFINITY;
    mine.idx = 0x7fffffff;
    mine.v2 = INFINITY;
    for (int t = tid; t < tab.n_coarse; t += LS_THREADS) {
        const double2 cs = reinterpret_cast<const double2 *>(tab.coarse_cs)[t];
        double v = pair_cost(b, cs.x, cs.y, i_in, j_in, flags);
        if (!(v == v)) v = INFINITY;            // a NaN never wins a '>' comparison
        MinIdx cand;
        cand.v = v;
        cand.idx = t;
        cand.v2 = INFINITY;
        mine = merge(mine, cand);
    }
    mine = warp_min(mine);
    if (lane == 0) sm.warp_best[warp] = mine;
    __syncthreads();
    if (warp == 0) {
        MinIdx x = (lane < LS_THREADS / 32) ? sm.warp_best[lane] : MinIdx{INFINITY, 0x7fffffff, INFINITY};
        x = warp_min(x);
        if (lane == 0) sm.best = x;
    }
    __syncthreads();
    const MinIdx best = sm.best;
    const bool improved = cost0 > best.v;
    const int row = improved ? best.idx : 0;
    double cur = improved ? best.v : cost0;
    const double coarse_margin = fmin(fabs(cost0 - best.v), best.v2 - best.v);

    // ---- fine grid of that coarse point ---------------------------------------------------
    const int L = min(tab.fine_len[row], LS_MAX_FINE);
    const double2 *fcs = reinterpret_cast<const double2 *>(tab.fine_cs) + (size_t)row * tab.fine_stride;
    for (int t = tid; t < L; t += LS_THREADS) {
        const double2 cs = fcs[t];
        sm.fine_val[t] = pair_cost(b, cs.x, cs.y, i_in, j_in, flags);
    }
    const int any_warn = __syncthreads_or(flags & QIO_B200_FLAG_WARN_NEGATIVE);
    const int any_raise = __syncthreads_or(flags & QIO_B200_FLAG_RAISE_GT1);

    if (tid == 0) {
        int fidx = -1;
        double margin = INFINITY;
#pragma unroll 8
        for (int k = 0; k < L; ++k) {
            const double v = sm.fine_val[k];
            const double thr = __dadd_rn(v, 1e-8);
            margin = fmin(margin, fabs(cur - thr));
            if (cur > thr) {
                cur = v;
                fidx = k;
            }
        }
        // stash the scan result in the (now free) reduction slot
        sm.best.v = cur;
        sm.best.idx = fidx;
        sm.best.v2 = margin;
    }
    __syncthreads();
    const int fidx = sm.best.idx;
    res.coarse_index = improved ? best.idx : -1;
    res.fine_index = fidx;
    res.flags = (any_warn ? QIO_B200_FLAG_WARN_NEGATIVE : 0) | (any_raise ? QIO_B200_FLAG_RAISE_GT1 : 0);
    res.accepted = fidx >= 0 ? 1 : 0;
    res.cost = sm.best.v;
    res.cost0 = cost0;
    res.coarse_margin = coarse_margin;
    res.fine_margin = sm.best.v2;
    if (fidx >= 0) {
        const size_t e = (size_t)row * tab.fine_stride + fidx;
        res.theta = tab.fine_theta[e];
        res.cos_theta = tab.fine_cs[2 * e];
        res.sin_theta = tab.fine_cs[2 * e + 1];
    } else {
        res.theta = __longlong_as_double(0x7ff8000000000000LL);
        res.cos_theta = 1.0;
        res.sin_theta = 0.0;
    }
    res.reserved = 0.0;
    __syncthreads();   // smem may be reused by the caller
}

}  // namespace qio
