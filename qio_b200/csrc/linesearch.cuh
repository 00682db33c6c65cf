// CTA-wide angle line search for ONE pair (reference qio/grad/jacobi.py:117-158), shared by the
// batched all-pairs kernel (pairs.cu) and the Jacobi sweeps (sweep.cu, sweep2.cu).
//
// One thread per candidate angle.  The reference's sequential scans are reproduced exactly:
//   coarse: running minimum with strict '>'  == first index of the global minimum, if it is
//           strictly below cost(0); otherwise t_opt falls back to 0.01 == coarse point 0;
//   fine:   accept k iff cost > cost_k + 1e-8, then cost = cost_k.  The chain is sequential in
//           principle, but it decomposes into "find the first k beyond the last accepted one with
//           cost_k + 1e-8 < cost" (ballots of one warp that holds all entries) followed by a run of consecutive accepts
//           (a_k := cost_{k-1} > cost_k + 1e-8, precomputed as a bit mask).  A smooth cost curve needs
//           two rounds; any other sequence just takes more rounds with the same result.
// The NT participating threads (the first NT of the CTA) synchronise on a NAMED barrier, so a CTA may
// have more threads that do something else (or wait elsewhere) meanwhile.
#pragma once
#include <type_traits>

#include "pair_math.cuh"

namespace qio {

constexpr int LS_THREADS = 320;           // >= 314 coarse points + 1, 10 warps
constexpr int LS_MAX_FINE = 256;          // capacity of the fine-value buffers (8 warps x 32)
constexpr int LS_BARRIER = 1;             // named barrier id used by the line search

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

template <int NT>
struct LineSearchSmemT {
    double fine_val[LS_MAX_FINE];     // cost at each fine angle
    double fine_thr[LS_MAX_FINE];     // cost + 1e-8 (rounded add, as the reference's `new_cost + 1e-8`)
    double2 fine_cs[LS_MAX_FINE];     // cos, sin of each fine angle
    int scan_last;                    // results of warp 0's accept scan
    double scan_cur;
    double scan_margin;               // smallest |cost - (new_cost + 1e-8)| over every comparison of the scan
    MinIdx warp_best[NT / 32];
    MinIdx best;
    double cost0;
    int flags;
    long long prof[8];                // cycles thread 0 spent per stage (coarse, reduce, fine, scan, margin, tail)
    long long prof_last;
};
using LineSearchSmem = LineSearchSmemT<LS_THREADS>;

// stage timers of thread 0 (cycles): only in translation units that define QIO_LS_TIMED (the timed sweep kernels)
#ifdef QIO_LS_TIMED
#define LS_PROF(k)                                  \
    if (tid == 0) {                                 \
        const long long now_ = clock64();           \
        sm.prof[k] += now_ - sm.prof_last;          \
        sm.prof_last = now_;                        \
    }
#else
#define LS_PROF(k)
#endif

template <int NT>
__device__ __forceinline__ void ls_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(LS_BARRIER), "n"(NT) : "memory");
}

__device__ __forceinline__ double warp_min_double(double v) { return dunkey(warp_min_u64(dkey(v))); }

// Threads 0..NT-1 of the CTA (whole warps, NT >= 256) must call this with the same arguments; the others must not.
// The result is returned in every participating thread.  `b` is either a PairBlock / PairBlockRef (reference operation
// order) or a PairCoef (fast polynomial form): pair_cost() is overloaded.
struct NoDecisionHook {
    __device__ __forceinline__ void operator()(bool, double, double) const {}
};

// `decided(accepted, cos, sin)` is called by thread 0 as soon as the angle is known (before the margins are taken): the
// pipelined sweep publishes the angle there.
// FINAL_SYNC = false leaves out the closing barrier: the caller then must not touch `sm` before its own next barrier of
// the same NT threads (the pipelined sweep's block contraction provides two).
// MARGINS = false skips the two diagnostic margins (reported as +inf): the pipelined sweep keeps them off its per-pair
// critical path -- the decision audit re-derives every decision, margins included, from the pair blocks in one batched
// launch afterwards.
template <int NT, typename Block, typename Hook = NoDecisionHook, bool FINAL_SYNC = true, bool MARGINS = true>
__device__ __forceinline__ void cta_linesearch_t(const Block &b, bool i_in, bool j_in, const ThetaTablesDev &tab,
                                                 LineSearchSmemT<NT> &sm, qio_b200_ls_result &res, Hook decided = Hook()) {
    static_assert(NT >= LS_MAX_FINE && NT % 32 == 0, "line search needs at least 256 threads");
    constexpr int NW = LS_MAX_FINE / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int flags = 0;
    if (tid == 0) {
        sm.prof_last = clock64();
        // PairCoef (fast) evaluations OR their flags straight into sm.flags: the CALLER zeroes it (thread 0, next to
        // make_coef, before the barrier that precedes this call); the other block types collect flags per thread
        if (!std::is_same<Block, PairCoef>::value) sm.flags = 0;
    }

    // ---- coarse grid; candidate index n_coarse is theta = 0 (cos = 1, sin = 0 exactly as np.cos(0)) ----
    MinIdx mine;
    mine.v = INFINITY;
    mine.idx = 0x7fffffff;
    mine.v2 = INFINITY;
    for (int t = tid; t <= tab.n_coarse; t += NT) {
        // one evaluation site for both the grid points and theta = 0 (the pair cost is inlined: a call would spill the
        // caller's live registers to local memory around it, and those reloads miss L1 in the persistent kernels)
        const bool zero = (t == tab.n_coarse);
        const double2 cs = zero ? make_double2(1.0, 0.0) : reinterpret_cast<const double2 *>(tab.coarse_cs)[t];
        double v = pair_cost(b, cs.x, cs.y, i_in, j_in, flags, &sm.flags);
        if (zero) {
            sm.cost0 = v;
        } else {
            if (!(v == v)) v = INFINITY;            // a NaN never wins a '>' comparison
            MinIdx cand;
            cand.v = v;
            cand.idx = t;
            cand.v2 = INFINITY;
            mine = merge(mine, cand);
        }
    }
    LS_PROF(0)
    mine = warp_min<MARGINS>(mine);
    if (lane == 0) sm.warp_best[warp] = mine;
    ls_sync<NT>();
    MinIdx best = (lane < NT / 32) ? sm.warp_best[lane] : MinIdx{INFINITY, 0x7fffffff, INFINITY};
    best = warp_min<MARGINS>(best);     // every warp reduces the NT/32 partial results redundantly: no second barrier
    const double cost0 = sm.cost0;
    const bool improved = cost0 > best.v;
    const int row = improved ? best.idx : 0;
    const double start = improved ? best.v : cost0;
    const double coarse_margin = MARGINS ? fmin(fabs(cost0 - best.v), best.v2 - best.v) : INFINITY;
    LS_PROF(1)

    // ---- fine grid of that coarse point ---------------------------------------------------
    const int L = min(tab.fine_len[row], LS_MAX_FINE);
    const double2 *fcs = reinterpret_cast<const double2 *>(tab.fine_cs) + (size_t)row * tab.fine_stride;
    double2 my_cs = make_double2(1.0, 0.0);
    double my_thr = INFINITY;
    if (tid < LS_MAX_FINE) {
        double v = INFINITY;
        if (tid < L) {
            my_cs = fcs[tid];
            v = pair_cost(b, my_cs.x, my_cs.y, i_in, j_in, flags, &sm.flags);
        }
        my_thr = __dadd_rn(v, 1e-8);                // +inf (or NaN) beyond L: never accepted
        sm.fine_val[tid] = v;
        sm.fine_thr[tid] = my_thr;
        sm.fine_cs[tid] = my_cs;
    }
    if (flags) atomicOr(&sm.flags, flags);
    ls_sync<NT>();
    LS_PROF(2)

    // ---- the accept scan, by warp 0 alone (lane l holds the entries l, l + 32, ...): no CTA-wide barrier inside the
    // rounds, one warp-wide integer minimum per round.  The margin -- the smallest |cost - (new_cost + 1e-8)| over every
    // comparison the reference's sequential scan makes -- is collected on the way: entries between the last accepted one and
    // the next accepted one were compared against the running cost, the entries of a run of consecutive accepts each
    // against its predecessor. ------------------------------------------------------------------------------------------
    if (warp == 0) {
        double thr[NW], dprev[NW];
        unsigned am[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const int k = lane + 32 * j;
            thr[j] = sm.fine_thr[k];
            const double pv = sm.fine_val[k > 0 ? k - 1 : 0];
            const bool a = k > 0 && k < L && pv > thr[j];     // cost_{k-1} > cost_k + 1e-8
            am[j] = __ballot_sync(0xffffffffu, a);
            dprev[j] = fabs(pv - thr[j]);
        }
        int last = -1;          // last accepted index (uniform)
        double cur = start;     // running cost (uniform)
        double mg = INFINITY;
        for (;;) {
            // first k beyond `last` with cost_k + 1e-8 < cost (k < L implied: thr = +inf beyond)
            int mine = 0x7fffffff;
#pragma unroll
            for (int j = NW - 1; j >= 0; --j)
                if (lane + 32 * j > last && thr[j] < cur) mine = lane + 32 * j;
            const int first = (int)__reduce_min_sync(0xffffffffu, (unsigned)mine);
            // every entry in (last, first] (to the end of the grid if nothing is accepted any more) met the running cost
            if (MARGINS) {
                const int upto = first == 0x7fffffff ? L - 1 : first;
#pragma unroll
                for (int j = 0; j < NW; ++j) {
                    const int k = lane + 32 * j;
                    if (k > last && k <= upto) mg = fmin(mg, fabs(cur - thr[j]));
                }
            }
            if (first == 0x7fffffff) break;
            // extend over the run of consecutive accepts that follows: first zero bit of the mask above `first`
            // (bits at and beyond L are zero by construction, so the run always ends before L)
            int e = LS_MAX_FINE - 1;
            {
                int w = (first + 1) >> 5;
                unsigned zeros = w < NW ? ~am[0] : 0u;
#pragma unroll
                for (int q = 1; q < NW; ++q) zeros = (w == q) ? ~am[q] : zeros;
                if (w < NW) zeros &= ~0u << ((first + 1) & 31);
                while (w < NW && !zeros) {
                    ++w;
                    zeros = 0u;
#pragma unroll
                    for (int q = 1; q < NW; ++q) zeros = (w == q) ? ~am[q] : zeros;
                }
                if (w < NW) e = w * 32 + __ffs(zeros) - 2;                       // last set bit before the first zero
            }
            if (MARGINS) {
#pragma unroll
                for (int j = 0; j < NW; ++j) {                               // the run (first, e]: each against its predecessor
                    const int k = lane + 32 * j;
                    if (k > first && k <= e) mg = fmin(mg, dprev[j]);
                }
            }
            last = e;
            cur = sm.fine_val[e];
        }
        // the angle is known: thread 0 hands it to the caller before anybody waits at the barrier below
        if (lane == 0) decided(last >= 0, last >= 0 ? sm.fine_cs[last].x : 1.0, last >= 0 ? sm.fine_cs[last].y : 0.0);
        if (MARGINS) mg = warp_min_double(mg);
        if (lane == 0) {
            sm.scan_last = last;
            sm.scan_cur = cur;
            sm.scan_margin = mg;
        }
    }
    ls_sync<NT>();
    const int last = sm.scan_last;
    const double cur = sm.scan_cur, margin = sm.scan_margin;
    LS_PROF(3)
    LS_PROF(4)
    const int fidx = last;
    res.coarse_index = improved ? best.idx : -1;
    res.fine_index = fidx;
    res.flags = sm.flags & (QIO_B200_FLAG_WARN_NEGATIVE | QIO_B200_FLAG_RAISE_GT1);
    res.accepted = fidx >= 0 ? 1 : 0;
    res.cost = cur;
    res.cost0 = cost0;
    res.coarse_margin = coarse_margin;
    res.fine_margin = margin;
    if (fidx >= 0) {
        res.theta = tab.fine_theta[(size_t)row * tab.fine_stride + fidx];     // for the record only
        res.cos_theta = sm.fine_cs[fidx].x;
        res.sin_theta = sm.fine_cs[fidx].y;
    } else {
        res.theta = __longlong_as_double(0x7ff8000000000000LL);
        res.cos_theta = 1.0;
        res.sin_theta = 0.0;
    }
    res.reserved = 0.0;
    LS_PROF(5)
    if (FINAL_SYNC) ls_sync<NT>();   // smem may be reused by the caller
}

__device__ __forceinline__ void cta_linesearch(const PairBlock &b, bool i_in, bool j_in, const ThetaTablesDev &tab,
                                               LineSearchSmem &sm, qio_b200_ls_result &res) {
    cta_linesearch_t<LS_THREADS, PairBlock>(b, i_in, j_in, tab, sm, res);
}

}  // namespace qio
