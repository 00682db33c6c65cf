// a-7 (fast path)  The Jacobi cycle as ONE persistent cooperative kernel over a deferred-axis layout,
// with the line search of pair t+1 overlapped with the Givens update of pair t.
//
// State.  S holds the 2-RDM with axes 1, 2 in the CURRENT orbital basis and axes 0, 3 still in the basis
// of the last flush; W (n x n) is the rotation accumulated since then:
//     Gamma_cur[a,b,c,d] = sum_{a',d'} W[a,a'] W[d,d'] S[a',b,c,d'].
// A rank holds the slabs S[a0 : a0+na] of the leading (deferred) axis; gamma, U, W are replicated.
// Why.  A Givens rotation of axis 3 in the natural layout touches 2 n^3 isolated 8-byte entries (one
// 32-byte sector each; the materialised kernel measured 2.5-3x DRAM over-fetch), and axis 0 is the sharded
// axis.  Deferring exactly those two axes leaves only row-contiguous, shard-local work per accepted rotation:
// 4 n^3 doubles read + written = 64 n^3 B, all coalesced 16-byte accesses.  The 16-number pair block then is
//     B[al,be,ga,de] = sum_{a'} W[al,a'] ( S[a',be,ga,:] . W[de,:] )        al,be,ga,de in {i,j},
// a sum over the leading axis: each rank contributes its slabs, 256 doubles cross NVLink per step.
//
// Pipeline.  The pair order is a strict chain, but the block of pair t+1 AFTER rotation t is a linear image of
// the "super block" over the <= 4 orbitals D = {i_t, j_t, i_t+1, j_t+1} taken BEFORE rotation t:
//     SB[p,q,r,s] = sum_{a'} W[D_p,a'] ( S[a',D_q,D_r,:] . W[D_s,:] )        (<= 256 numbers)
//     B_{t+1} = T^(x4) SB,  T = 2 x |D| coefficients of (i_t+1, j_t+1) after the rotation (cos, sin of pair t).
// So as soon as theta_t is known, CTA 0 can search pair t+1 while all other CTAs stream the update of pair t, and the
// super block needed for pair t+2 is produced INSIDE that update: the few d-rows (a', x, y, :) with x, y in D_{t+1} are
// owned by dedicated warps which load the orbit of such a row under rotation t (1, 2 or 4 rows), rotate it in registers,
// write it back and take the dot products of the NEW values with the NEW rows of W on the fly.  One grid barrier per pair:
//   loop t:   [ CTA 0: gather SB_t partials (ranks: peer mailboxes); B_{t+1} <- SB_t, theta_t; line search t+1 ]
//          || [ CTAs 1..: super-block orbits of SB_{t+1} (update + dots), flat Givens update of everything else,
//                         gamma, W (ping-pong buffers), U ]
//             -- grid barrier --
// The flush (W applied to axes 0 and 3 with two DMMA GEMM passes) happens once, when the caller wants Gamma back.
#include <string.h>

#include <type_traits>

#define QIO_LS_TIMED 1      // this kernel reports the line-search stage timers in its summary
#include "linesearch.cuh"
#include "ll.cuh"
#include "sweep3.cuh"

namespace qio {

constexpr int SW_THREADS = 512;
constexpr int SW_WARPS = SW_THREADS / 32;      // 16 = one (q, r) row task per warp
constexpr int SB_DOUBLES = 256 + 32;           // super block + the 2 x 4 x 4 gamma entries
constexpr int MB_STRIDE = 512;                 // mailbox slot: 256 x (value, tag) 16-byte words, in doubles

struct SweepParams {
    double *S, *gamma, *U, *W;
    int n, a0, na;
    const int32_t *pairs;
    int P;
    const int32_t *inactive;
    ThetaTablesDev tab;
    qio_b200_ls_result *results;
    double *W2;                // second W buffer (ping-pong), n x n, in the workspace
    double *dbg;               // [2 * WS_MAX_CTAS] per-CTA arrival / release times of one mid-sweep barrier (ns)
    double *pbuf;              // [2][nsb + 1][SB_DOUBLES] (parity of the step that reads it)
    qio_b200_ls_result *cur;   // [2] published result of the pair being applied (parity of the step)
    unsigned *bar;             // [0] arrival counter (monotone), [1] generation
    double *summary;           // [32]
    int rank, nranks;
    unsigned epoch;            // launch epoch of the mailbox tags (ll.cuh: mailbox_tag)
    double *blocks_out;        // (P, 24) or NULL: the block each search ran on (decision audit)
    double *const *mailboxes;  // [nranks] peer-mapped mailbox bases, layout [2][nranks][MB_STRIDE]; NULL if nranks == 1
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Monotone-counter grid barrier (all CTAs co-resident: cooperative launch).  The arrival is an acq_rel atomic and
// the generation is published / polled with release / acquire accesses, so no separate fences are needed.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks, unsigned &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        unsigned prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
        if (prev == nblocks * epoch - 1) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(epoch) : "memory");
        } else {
            while (ld_acquire_u32(bar + 1) < epoch) {
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ double2 ldcg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

// (x, y) -> (c x + s y, c y - s x)
__device__ __forceinline__ void butterfly2(double *px, double *py, double c, double s) {
    const double2 x = ldcg2(px), y = ldcg2(py);
    *reinterpret_cast<double2 *>(px) = make_double2(c * x.x + s * y.x, c * x.y + s * y.y);
    *reinterpret_cast<double2 *>(py) = make_double2(c * y.x - s * x.x, c * y.y - s * x.y);
}
__device__ __forceinline__ void butterfly1(double *px, double *py, double c, double s) {
    const double x = __ldcg(px), y = __ldcg(py);
    *px = c * x + s * y;
    *py = c * y - s * x;
}

// Mixed-radix counter (a, r, v), r < nr, v < nv, advanced by a fixed stride without any division in the loop.
struct Counter3 {
    int a, r, v, nr, nv, da, dr, dv;
    __device__ __forceinline__ Counter3(long long start, long long stride, int nr_, int nv_) : nr(nr_), nv(nv_) {
        long long t = start / nv_;
        v = (int)(start - t * nv_);
        a = (int)(t / nr_);
        r = (int)(t - (long long)a * nr_);
        t = stride / nv_;
        dv = (int)(stride - t * nv_);
        da = (int)(t / nr_);
        dr = (int)(t - (long long)da * nr_);
    }
    __device__ __forceinline__ void step() {
        v += dv;
        r += dr;
        a += da;
        if (v >= nv) {
            v -= nv;
            ++r;
        }
        if (r >= nr) {
            r -= nr;
            ++a;
        }
    }
};

// Rotation (i, j, c, s) of axes 1 and 2 of the `na` local slabs.  V = 2 for even n (16-byte accesses).
// Work items are 16-byte words of d-rows, (slab, row, vector column) enumerated flat and grid-strided, so
// that consecutive threads touch consecutive words.
template <int V>
static __device__ __noinline__ void update_S(double *__restrict__ S, int n, int na, int i, int j, double c, double s,
                                         long long gtid, long long gthreads, int4 skip, bool skip_orbit4) {
    const long long n1 = n, n2 = n1 * n1, n3 = n2 * n1;
    const int rowv = n / V;          // vector elements per d-row
    // axis 1: rows (a', {i,j}, cc, :) for all (a', cc), cc in {i, j} excluded (they belong to the 4-orbit)
    // axis 2: rows (a', b, {i,j}, :) for all (a', b), b in {i, j} excluded
    // (`skip` lists up to four row indices whose orbits are owned by the super-block warps, -1 = none)
    // UNR butterflies (2 UNR independent 16-byte loads) are in flight per thread: the update is a latency x
    // bandwidth product (Little's law), one butterfly per thread at a time leaves most of L2 / HBM idle.
    constexpr int UNR = 4;
    using VT = typename std::conditional<V == 2, double2, double>::type;
#pragma unroll 1
    for (int axis = 0; axis < 2; ++axis) {
        const long long si = axis == 0 ? (long long)i * n2 : (long long)i * n1, sj = axis == 0 ? (long long)j * n2 : (long long)j * n1;
        const long long sr = axis == 0 ? n1 : n2;
        Counter3 k(gtid, gthreads, n, rowv);
        while (k.a < na) {
            long long off[UNR];          // element offset of the butterfly, -1 = nothing to do
            VT x[UNR], y[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const bool ok = k.a < na && !(k.r == i || k.r == j || k.r == skip.x || k.r == skip.y || k.r == skip.z || k.r == skip.w);
                off[u] = ok ? k.a * n3 + k.r * sr + (long long)k.v * V : -1;
                k.step();
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (off[u] >= 0) {
                    x[u] = __ldcg(reinterpret_cast<const VT *>(S + off[u] + si));
                    y[u] = __ldcg(reinterpret_cast<const VT *>(S + off[u] + sj));
                }
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (off[u] >= 0) {
                    VT *px = reinterpret_cast<VT *>(S + off[u] + si), *py = reinterpret_cast<VT *>(S + off[u] + sj);
                    if constexpr (V == 2) {
                        *px = make_double2(c * x[u].x + s * y[u].x, c * x[u].y + s * y[u].y);
                        *py = make_double2(c * y[u].x - s * x[u].x, c * y[u].y - s * x[u].y);
                    } else {
                        *px = c * x[u] + s * y[u];
                        *py = c * y[u] - s * x[u];
                    }
                }
        }
    }
    // the 4-row orbit (a', {i,j}, {i,j}, :): both axes
    if (skip_orbit4) return;
    for (Counter3 k(gtid, gthreads, 1, n); k.a < na; k.step()) {
        double *base = S + k.a * n3 + k.v;
        double *pii = base + (long long)i * n2 + (long long)i * n1, *pij = base + (long long)i * n2 + (long long)j * n1;
        double *pji = base + (long long)j * n2 + (long long)i * n1, *pjj = base + (long long)j * n2 + (long long)j * n1;
        const double vii = __ldcg(pii), vij = __ldcg(pij), vji = __ldcg(pji), vjj = __ldcg(pjj);
        const double t0 = c * vii + s * vij, t1 = c * vij - s * vii, t2 = c * vji + s * vjj, t3 = c * vjj - s * vji;
        *pii = c * t0 + s * t2;
        *pji = c * t2 - s * t0;
        *pij = c * t1 + s * t3;
        *pjj = c * t3 - s * t1;
    }
}

// Distinct orbitals of (i, j, i2, j2): D[0] = i, D[1] = j, then the new ones; pos[k] = index of entry k in D.
struct Distinct {
    int D[4];
    int pos[4];
    int nd;
};
__device__ __forceinline__ Distinct make_distinct(int i, int j, int i2, int j2) {
    Distinct d;
    d.D[0] = i;
    d.D[1] = j;
    d.D[2] = d.D[3] = i;
    d.pos[0] = 0;
    d.pos[1] = 1;
    d.nd = 2;
    const int extra[2] = {i2, j2};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int x = extra[k];
        int found = -1;
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (m < d.nd && d.D[m] == x && found < 0) found = m;
        if (found < 0) {
            found = d.nd;
            d.D[d.nd++] = x;
        }
        d.pos[2 + k] = found;
    }
    return d;
}

// ---- super-block orbits -------------------------------------------------------------------------------------
// One task = one orbit of d-rows (b, c) under the rotation being applied (1, 2 or 4 rows) that contains at least one
// row of the next super block, for one leading-axis slab.
struct SbTask {
    int nrows;          // 1, 2 or 4
    int kind;           // 0 single row, 1 axis-1 pair ((i,f),(j,f)), 2 axis-2 pair ((f,i),(f,j)), 3 the (i,j)x(i,j) orbit
    int x[4], y[4];     // (b, c) of the rows
    int nmem;           // rows of this orbit that belong to the super block (<= 4)
    int mslot[4], mq[4], mr[4];     // their slot in the orbit and their (q, r) position in the super block
};
constexpr int SB_MAX_TASKS = 16;

struct SbPlan {
    SbTask task[SB_MAX_TASKS];
    int ntasks;
    int4 skip;          // rows the flat update must leave to the tasks
    int skip_orbit4;
};

// Thread 0 of a CTA: orbits of the super block over dd (orbitals of pairs t+1, t+2) under rotation (i, j).
__device__ __forceinline__ void build_sb_plan(SbPlan &pl, const Distinct &dd, int i, int j, bool rotated) {
    int cls[4], nE = 0, F[4], Fq[4], nF = 0;
    for (int q = 0; q < dd.nd; ++q) {
        cls[q] = rotated ? (dd.D[q] == i ? 1 : dd.D[q] == j ? 2 : 0) : 0;
        if (cls[q]) ++nE;
        else {
            F[nF] = dd.D[q];
            Fq[nF++] = q;
        }
    }
    int nt = 0;
    for (int a = 0; a < nF; ++a)
        for (int b = 0; b < nF; ++b) {
            SbTask &t = pl.task[nt++];
            t.nrows = 1; t.kind = 0; t.x[0] = F[a]; t.y[0] = F[b];
            t.nmem = 1; t.mslot[0] = 0; t.mq[0] = Fq[a]; t.mr[0] = Fq[b];
        }
    pl.skip = make_int4(-1, -1, -1, -1);
    pl.skip_orbit4 = 0;
    if (nE > 0) {
        for (int a = 0; a < nF; ++a) {
            {   // rows (i, f), (j, f): rotation of axis 1
                SbTask &t = pl.task[nt++];
                t.nrows = 2; t.kind = 1; t.x[0] = i; t.y[0] = F[a]; t.x[1] = j; t.y[1] = F[a]; t.nmem = 0;
                for (int q = 0; q < dd.nd; ++q)
                    if (cls[q]) { t.mslot[t.nmem] = cls[q] - 1; t.mq[t.nmem] = q; t.mr[t.nmem] = Fq[a]; ++t.nmem; }
            }
            {   // rows (f, i), (f, j): rotation of axis 2
                SbTask &t = pl.task[nt++];
                t.nrows = 2; t.kind = 2; t.x[0] = F[a]; t.y[0] = i; t.x[1] = F[a]; t.y[1] = j; t.nmem = 0;
                for (int r = 0; r < dd.nd; ++r)
                    if (cls[r]) { t.mslot[t.nmem] = cls[r] - 1; t.mq[t.nmem] = Fq[a]; t.mr[t.nmem] = r; ++t.nmem; }
            }
        }
        SbTask &t = pl.task[nt++];
        t.nrows = 4; t.kind = 3; t.nmem = 0;
        t.x[0] = i; t.y[0] = i; t.x[1] = i; t.y[1] = j; t.x[2] = j; t.y[2] = i; t.x[3] = j; t.y[3] = j;
        for (int q = 0; q < dd.nd; ++q)
            for (int r = 0; r < dd.nd; ++r)
                if (cls[q] && cls[r]) {
                    t.mslot[t.nmem] = (cls[q] - 1) * 2 + (cls[r] - 1); t.mq[t.nmem] = q; t.mr[t.nmem] = r; ++t.nmem;
                }
        pl.skip = make_int4(nF > 0 ? F[0] : -1, nF > 1 ? F[1] : -1, nF > 2 ? F[2] : -1, nF > 3 ? F[3] : -1);
        pl.skip_orbit4 = 1;
    }
    pl.ntasks = nt;
}

// One warp, one task, one slab: rotate the orbit (if any), write it back, and accumulate the super-block contributions
// of its member rows into this warp's private accumulator.  Wn = the nd rows of W AFTER the rotation, in shared memory.
static __device__ __noinline__ void run_sb_task(const SweepParams &p, const SbTask &tk, int al, bool rotated, double c, double s,
                                            const double *Wn, int nd, double *sbw) {
    const int n = p.n, lane = threadIdx.x & 31;
    const long long n1 = n, n2 = n1 * n1, n3 = n2 * n1;
    double *row[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) row[k] = p.S + al * n3 + tk.x[k < tk.nrows ? k : 0] * n2 + tk.y[k < tk.nrows ? k : 0] * n1;
    double dot[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int q = 0; q < 4; ++q) dot[m][q] = 0.0;
    const bool write = rotated && tk.nrows > 1;
    // four 32-column chunks per trip: up to 16 independent loads per lane in flight (an orbit visit is latency-bound)
    for (int d0 = lane; d0 < n; d0 += 128) {
        double vv[4][4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int k = 0; k < 4; ++k) vv[ch][k] = (k < tk.nrows && d0 + 32 * ch < n) ? __ldcg(row[k] + d0 + 32 * ch) : 0.0;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const int d = d0 + 32 * ch;
            if (d >= n) break;
            double *v = vv[ch];
            if (write) {
                if (tk.kind == 3) {
                    const double t0 = c * v[0] + s * v[1], t1 = c * v[1] - s * v[0], t2 = c * v[2] + s * v[3], t3 = c * v[3] - s * v[2];
                    v[0] = c * t0 + s * t2;     // (i, i)
                    v[2] = c * t2 - s * t0;     // (j, i)
                    v[1] = c * t1 + s * t3;     // (i, j)
                    v[3] = c * t3 - s * t1;     // (j, j)
                } else {
                    const double x = v[0], y = v[1];
                    v[0] = c * x + s * y;
                    v[1] = c * y - s * x;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < tk.nrows) row[k][d] = v[k];
            }
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (m < tk.nmem) {
                    const int sl = tk.mslot[m];
                    const double val = sl == 0 ? v[0] : sl == 1 ? v[1] : sl == 2 ? v[2] : v[3];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q < nd) dot[m][q] += val * Wn[q * n + d];
                }
        }
    }
    const int pl = (lane >> 2) & 3, sl = lane & 3;          // lanes 0..15 own (p, s)
#pragma unroll
    for (int m = 0; m < 4; ++m)
        if (m < tk.nmem) {
            double ds = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q < nd) {
                    const double t = warp_sum(dot[m][q]);
                    if (sl == q) ds = t;
                }
            if (lane < 16 && pl < nd && sl < nd)
                sbw[((pl * 4 + tk.mq[m]) * 4 + tk.mr[m]) * 4 + sl] += Wn[pl * n + p.a0 + al] * ds;
        }
}

// B (24 doubles in smem) of the target pair from the super block: coefficient rows over D of the two targets.
struct Target {
    int idx[2][2];      // [target][term] index into D
    double coef[2][2];
    int cnt[2];
};
__device__ __forceinline__ Target make_target(const Distinct &dd, int which_i, int which_j, bool rotated, double c, double s) {
    // which_i / which_j: entry (0..3) of the (i, j, i2, j2) list that is the target orbital
    Target t;
    const int w[2] = {which_i, which_j};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int dpos = dd.pos[w[k]];
        if (rotated && dpos == 0) {            // the target is orbital i of the rotation: i' = c i + s j
            t.cnt[k] = 2; t.idx[k][0] = 0; t.coef[k][0] = c; t.idx[k][1] = 1; t.coef[k][1] = s;
        } else if (rotated && dpos == 1) {     // orbital j: j' = -s i + c j
            t.cnt[k] = 2; t.idx[k][0] = 0; t.coef[k][0] = -s; t.idx[k][1] = 1; t.coef[k][1] = c;
        } else {
            t.cnt[k] = 1; t.idx[k][0] = dpos; t.coef[k][0] = 1.0; t.idx[k][1] = dpos; t.coef[k][1] = 0.0;
        }
    }
    return t;
}
// Dense 2 x 4 coefficient matrix of the two target orbitals over D (zeros where absent), built by one thread.
__device__ __forceinline__ void target_matrix(const Target &t, double *T /*[2][4] smem*/) {
    for (int k = 0; k < 8; ++k) T[k] = 0.0;
    for (int k = 0; k < 2; ++k)
        for (int a = 0; a < t.cnt[k]; ++a) T[k * 4 + t.idx[k][a]] += t.coef[k][a];
}

// B = T^(x4) SB as four single-index contractions (256 -> 128 -> 64 -> 32 -> 16 numbers) by SEARCH_THREADS threads
// on the named barrier, plus the 2 x 2 x 2 gamma entries; then the polynomial coefficients of the pair cost.
constexpr int SEARCH_THREADS = LS_THREADS;
__device__ __forceinline__ void block_from_superblock(const double *sb /*[288] smem*/, const double *T /*[2][4]*/,
                                                      double *tmp /*[192] smem*/, double *blk /*[24]*/, PairCoef &pc) {
    const int o = threadIdx.x;
    double *y1 = tmp, *y2 = tmp + 128;      // y1[p][q][r][de] (128), y2[p][q][ga][de] (64); y3 reuses y1
    if (o < 128) {                          // contract s
        const int de = o & 1, pqr = o >> 1;
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s) acc += T[de * 4 + s] * sb[pqr * 4 + s];
        y1[o] = acc;
    }
    ls_sync<SEARCH_THREADS>();
    if (o < 64) {                           // contract r
        const int de = o & 1, ga = (o >> 1) & 1, pq = o >> 2;
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) acc += T[ga * 4 + r] * y1[(pq * 4 + r) * 2 + de];
        y2[o] = acc;
    } else if (o >= 64 && o < 72) {         // gamma entries: spin, al, be
        const int u = o - 64, spin = u >> 2, al = (u >> 1) & 1, be = u & 1;
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc += T[al * 4 + a] * T[be * 4 + c] * sb[256 + spin * 16 + a * 4 + c];
        blk[16 + u] = acc;
    }
    ls_sync<SEARCH_THREADS>();
    if (o < 32) {                           // contract q -> y3[p][be][ga][de] in y1
        const int gd = o & 3, be = (o >> 2) & 1, pp = o >> 3;
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += T[be * 4 + q] * y2[(pp * 4 + q) * 4 + gd];
        y1[o] = acc;
    }
    ls_sync<SEARCH_THREADS>();
    if (o < 16) {                           // contract p
        const int rest = o & 7, al = o >> 3;
        double acc = 0.0;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) acc += T[al * 4 + pp] * y1[pp * 8 + rest];
        blk[o] = acc;
    }
    ls_sync<SEARCH_THREADS>();
    if (o == 0) make_coef(blk, pc);
    ls_sync<SEARCH_THREADS>();
}

__global__ void __launch_bounds__(SW_THREADS, 1) sweep_kernel(SweepParams p) {
    __shared__ LineSearchSmemT<SEARCH_THREADS> ls;
    __shared__ double blk[QIO_B200_BLOCK_DOUBLES];
    __shared__ double sb[SB_DOUBLES];
    __shared__ PairCoef pc;
    __shared__ double2 coarse_cs_sm[320];
    __shared__ int32_t fine_len_sm[320];
    __shared__ SbPlan plan;
    __shared__ double Tmat[8];
    __shared__ double sbtmp[192];
    extern __shared__ double dyn_sm[];        // [4][n] rows of W after the rotation, then [SW_WARPS][256] accumulators
    const int n = p.n, tid = threadIdx.x, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    double *Wn = dyn_sm, *sbw = dyn_sm + 4 * (size_t)n;
    // super-block CTAs: CTAs 1..nsb (CTA 0 itself when the grid is a single CTA), apc slabs each
    const int nsb = G > 1 ? max(1, min(G - 1, (p.na + 2) / 3)) : 1;
    const int apc = (p.na + nsb - 1) / nsb;
    const int sb_idx = G > 1 ? cta - 1 : 0;                       // index among the super-block CTAs
    const bool is_sb_cta = (G == 1 || (cta >= 1 && cta <= nsb)) && sb_idx * apc < p.na;
    const size_t pbuf_par = (size_t)(nsb + 1) * SB_DOUBLES;       // doubles per parity of pbuf
    unsigned epoch = 0;
    int accepted_total = 0, flags_or = 0;
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1;
    const bool timing = ((cta == 0 || cta == 1 || cta == G - 3) && tid == 0);
#define SW_TICK(slot)              \
    if (timing) {                  \
        t1 = gtimer();             \
        tacc[slot] += t1 - t0;     \
        t0 = t1;                   \
    }

    if (tid < 8) ls.prof[tid] = 0;
    ThetaTablesDev tab = p.tab;
    if (cta == 0 && tab.n_coarse <= 320) {      // the search CTA keeps the coarse cos/sin table and the fine lengths in smem
        for (int t = tid; t < tab.n_coarse; t += SW_THREADS) {
            coarse_cs_sm[t] = reinterpret_cast<const double2 *>(tab.coarse_cs)[t];
            fine_len_sm[t] = tab.fine_len[t];
        }
        tab.coarse_cs = reinterpret_cast<const double *>(coarse_cs_sm);
        tab.fine_len = fine_len_sm;
    }
    // who does the small replicated updates
    const int cta_gamma = G > 1 ? G - 1 : 0, cta_wu = G > 2 ? G - 2 : (G > 1 ? 1 : 0);
    // update CTAs: all but CTA 0 (which searches), unless the grid is a single CTA
    const int upd_ctas = G > 1 ? G - 1 : 1, upd_cta = G > 1 ? cta - 1 : 0;
    const long long gtid = (long long)upd_cta * SW_THREADS + tid, gthreads = (long long)upd_ctas * SW_THREADS;
    double *Wbuf[2] = {p.W, p.W2};

    auto pair_at = [&](int t, int &i, int &j) {
        const int tt = t < p.P ? t : p.P - 1;
        i = p.pairs[2 * tt];
        j = p.pairs[2 * tt + 1];
    };

    // Super-block CTAs: partial super block over `dd` of the state AFTER rotation (i, j, c, s) -- the rotation of the
    // orbits it owns is applied here -- into pbuf[par].  Wcur = W before the rotation.
    auto phase_superblock = [&](const Distinct &dd, int i, int j, bool rotated, double c, double s, const double *Wcur,
                                int par) {
        for (int e = tid; e < dd.nd * n; e += SW_THREADS) {
            const int q = e / n, d = e - q * n, x = dd.D[q];
            double w;
            if (rotated && x == i) w = c * __ldcg(Wcur + (size_t)i * n + d) + s * __ldcg(Wcur + (size_t)j * n + d);
            else if (rotated && x == j) w = c * __ldcg(Wcur + (size_t)j * n + d) - s * __ldcg(Wcur + (size_t)i * n + d);
            else w = __ldcg(Wcur + (size_t)x * n + d);
            Wn[e] = w;
        }
        for (int e = tid; e < SW_WARPS * 256; e += SW_THREADS) sbw[e] = 0.0;
        if (tid == 0) build_sb_plan(plan, dd, i, j, rotated);
        __syncthreads();
        const int first = sb_idx * apc, nslab = min(apc, p.na - first);
        for (int slot = warp; slot < plan.ntasks * nslab; slot += SW_WARPS)
            run_sb_task(p, plan.task[slot % plan.ntasks], first + slot / plan.ntasks, rotated, c, s, Wn, dd.nd, sbw + warp * 256);
        __syncthreads();
        if (tid < 256) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < SW_WARPS; ++w) acc += sbw[w * 256 + tid];
            p.pbuf[par * pbuf_par + (size_t)sb_idx * SB_DOUBLES + tid] = acc;
        }
    };

    // the designated CTA: gamma entries over dd (after its own in-place update of gamma) into pbuf[par]
    auto publish_gamma_block = [&](const Distinct &dd, int par) {
        if (tid < 32) {
            const int spin = tid >> 4, pp = (tid >> 2) & 3, qq = tid & 3;
            double v = 0.0;
            if (pp < dd.nd && qq < dd.nd)
                v = __ldcg(p.gamma + (2 * (size_t)dd.D[pp] + spin) * (2 * (size_t)n) + 2 * dd.D[qq] + spin);
            p.pbuf[par * pbuf_par + (size_t)nsb * SB_DOUBLES + 256 + tid] = v;
        }
    };

    // CTA 0: sum the partials of parity `par` in a fixed order (CTAs, then ranks) into shared memory
    auto gather_superblock = [&](int t) {
        const double *src = p.pbuf + (t & 1) * pbuf_par;
        const int K = min(nsb, (p.na + apc - 1) / apc);
        if (tid < 256) {
            // 16 independent loads in flight per batch (the loads see loaded-L2 latency while the update streams);
            // fixed summation pattern: deterministic and identical on every rank
            double a16[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) a16[u] = 0.0;
            int c = 0;
            for (; c + 16 <= K; c += 16) {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = __ldcg(src + (size_t)(c + u) * SB_DOUBLES + tid);
#pragma unroll
                for (int u = 0; u < 16; ++u) a16[u] += v[u];
            }
            {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = (c + u < K) ? __ldcg(src + (size_t)(c + u) * SB_DOUBLES + tid) : 0.0;
#pragma unroll
                for (int u = 0; u < 16; ++u) a16[u] += v[u];
            }
#pragma unroll
            for (int w = 8; w > 0; w >>= 1)
#pragma unroll
                for (int u = 0; u < w; ++u) a16[u] += a16[u + w];
            sb[tid] = a16[0];
        } else if (tid < 288) {
            sb[tid] = __ldcg(src + (size_t)nsb * SB_DOUBLES + tid);
        }
        if (p.nranks > 1) {
            // Every 16-byte word carries the value and, in both 8-byte halves, the tag (ll.cuh): one NVLink store per
            // element, no fence, no separate flag -- the receiver spins on the word it needs.  Slots alternate with the
            // step parity; a slot is rewritten two steps later, after every reader has moved on.
            const int par = t & 1;
            const unsigned tag = mailbox_tag(p.epoch, t);
            ls_sync<SEARCH_THREADS>();
            if (tid < 256) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(sb[tid]);
                for (int dst = 0; dst < p.nranks; ++dst) {
                    unsigned long long *w = reinterpret_cast<unsigned long long *>(
                        p.mailboxes[dst] + ((size_t)par * p.nranks + p.rank) * MB_STRIDE) + 2 * tid;
                    ll_store(w, bits, tag);
                }
                double acc = 0.0;
                for (int r = 0; r < p.nranks; ++r) {
                    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(
                        p.mailboxes[p.rank] + ((size_t)par * p.nranks + r) * MB_STRIDE) + 2 * tid;
                    unsigned long long v;
                    while (!ll_load(w, v, tag)) {
                    }
                    acc += __longlong_as_double((long long)v);
                }
                sb[tid] = acc;
            }
        }
        ls_sync<SEARCH_THREADS>();
    };

    // CTA 0, threads < SEARCH_THREADS: line search of pair `t` (orbitals i, j) whose block is derived from the super
    // block in smem.  All synchronisation inside is on the named barrier of the line search.
    auto search_pair = [&](int t, int i, int j, const Target &tg) {
        long long c0 = 0;
        if (tid == 0) {
            c0 = clock64();
            target_matrix(tg, Tmat);
        }
        if (tid == 0) ls.flags = 0;         // the fast pair cost ORs its flags straight into it (see linesearch.cuh)
        const bool i_in = p.inactive[i] != 0, j_in = p.inactive[j] != 0;       // L2 latency overlaps the contraction
        ls_sync<SEARCH_THREADS>();
        block_from_superblock(sb, Tmat, sbtmp, blk, pc);
        if (p.blocks_out && tid < QIO_B200_BLOCK_DOUBLES) p.blocks_out[(size_t)t * QIO_B200_BLOCK_DOUBLES + tid] = blk[tid];
        if (tid == 0) {
            const long long c1 = clock64();
            ls.prof[6] += c1 - c0;
        }
        qio_b200_ls_result res;
        cta_linesearch_t<SEARCH_THREADS, PairCoef>(pc, i_in, j_in, tab, ls, res);
        if (tid == 0) {
            c0 = clock64();
            p.results[t] = res;
            p.cur[t & 1] = res;
            ls.prof[7] += clock64() - c0;
        }
        accepted_total += res.accepted;
        flags_or |= res.flags;
    };

    if (p.P > 0) {
        // ---- prologue: second W buffer, super block 0 (no rotation pending), search pair 0 ---------------
        if (timing) t0 = gtimer();
        for (long long e = (long long)cta * SW_THREADS + tid; e < (long long)n * n; e += (long long)G * SW_THREADS)
            p.W2[e] = __ldcg(p.W + e);
        {
            int i, j, i1, j1;
            pair_at(0, i, j);
            pair_at(1, i1, j1);
            const Distinct dd = make_distinct(i, j, i1, j1);
            if (is_sb_cta) phase_superblock(dd, i, j, false, 1.0, 0.0, Wbuf[0], 0);
            if (cta == cta_gamma) publish_gamma_block(dd, 0);
            grid_barrier(p.bar, G, epoch);
            if (cta == 0 && tid < SEARCH_THREADS) {
                gather_superblock(0);
                search_pair(0, i, j, make_target(dd, 0, 1, false, 1.0, 0.0));
            }
            grid_barrier(p.bar, G, epoch);
        }
        SW_TICK(0)

        for (int t = 0; t < p.P; ++t) {
            int i, j, i1, j1, i2, j2;
            pair_at(t, i, j);
            pair_at(t + 1, i1, j1);
            pair_at(t + 2, i2, j2);
            const bool rotated = __ldcg(&p.cur[t & 1].accepted) != 0;
            const double c = __ldcg(&p.cur[t & 1].cos_theta), s = __ldcg(&p.cur[t & 1].sin_theta);
            const bool need_sb = t + 2 < p.P;                 // SB_{t+1} feeds the search of pair t+2
            const Distinct dd1 = make_distinct(i1, j1, i2, j2);
            const double *Wcur = Wbuf[t & 1];
            double *Wnext = Wbuf[(t + 1) & 1];

            // ---- CTA 0: SB_t -> block of pair t+1 after rotation t -> line search -------------------------
            if (cta == 0 && tid < SEARCH_THREADS) {
                if (t > 0 && t + 1 < p.P) gather_superblock(t);
                SW_TICK(1)
                if (t + 1 < p.P) search_pair(t + 1, i1, j1, make_target(make_distinct(i, j, i1, j1), 2, 3, rotated, c, s));
                SW_TICK(2)
            }
            if (G == 1) __syncthreads();
            // ---- the other CTAs (or the only one): rotation t ---------------------------------------------
            if (G == 1 || cta != 0) {
                int4 skip = make_int4(-1, -1, -1, -1);
                bool skip4 = false;
                if (need_sb) {
                    if (is_sb_cta) {
                        phase_superblock(dd1, i, j, rotated, c, s, Wcur, (t + 1) & 1);
                        skip = plan.skip;
                        skip4 = plan.skip_orbit4 != 0;
                    } else if (rotated) {       // same plan, needed only for the rows to leave alone
                        int nE = 0, nF = 0, F[4] = {-1, -1, -1, -1};
                        for (int q = 0; q < dd1.nd; ++q) {
                            if (dd1.D[q] == i || dd1.D[q] == j) ++nE;
                            else F[nF++] = dd1.D[q];
                        }
                        if (nE > 0) {
                            skip = make_int4(F[0], F[1], F[2], F[3]);
                            skip4 = true;
                        }
                    }
                }
                if (cta == 1) { SW_TICK(4) }
                // flat update: by the CTAs that own no super-block orbits in this step (all update CTAs otherwise)
                // and not the two CTAs that carry the small replicated updates (they would be the stragglers)
                const int nsb_active = (need_sb && G > 1) ? min(nsb, (p.na + apc - 1) / apc) : 0;
                const int flat_ctas = upd_ctas - nsb_active - 2;
                const bool split = G > 8 && flat_ctas >= upd_ctas / 4;
                const long long f_threads = split ? (long long)flat_ctas * SW_THREADS : gthreads;
                const long long f_tid = split ? (long long)(upd_cta - nsb_active) * SW_THREADS + tid : gtid;
                if (rotated && f_tid >= 0 && f_tid < f_threads) {
                    if ((n & 1) == 0)
                        update_S<2>(p.S, n, p.na, i, j, c, s, f_tid, f_threads, skip, skip4);
                    else
                        update_S<1>(p.S, n, p.na, i, j, c, s, f_tid, f_threads, skip, skip4);
                }
                if (cta == 1 || cta == G - 3) { SW_TICK(5) }
            }
            if (cta == cta_gamma) {
                if (rotated) {
                    // gamma <- V gamma V^T on the same-spin blocks: columns, then rows
                    const size_t ld = 2 * (size_t)n;
                    for (int k = tid; k < 2 * n; k += SW_THREADS) {
                        const int a = k >> 1, spin = k & 1;
                        double *row = p.gamma + (2 * (size_t)a + spin) * ld;
                        butterfly1(row + 2 * i + spin, row + 2 * j + spin, c, s);
                    }
                    __syncthreads();
                    for (int k = tid; k < 2 * n; k += SW_THREADS) {
                        const int b = k >> 1, spin = k & 1;
                        butterfly1(p.gamma + (2 * (size_t)i + spin) * ld + 2 * b + spin,
                                   p.gamma + (2 * (size_t)j + spin) * ld + 2 * b + spin, c, s);
                    }
                    __syncthreads();
                }
                if (need_sb) publish_gamma_block(dd1, (t + 1) & 1);
            }
            if (cta == cta_wu) {
                // W (ping-pong: rows i, j rotated into the next buffer, plus the rows the previous step changed) and
                // U rows i, j in place (U <- V U, jacobi.py:230)
                int ip = -1, jp = -1;
                if (t > 0) pair_at(t - 1, ip, jp);
                for (int k = tid; k < n; k += SW_THREADS) {
                    const double x = __ldcg(Wcur + (size_t)i * n + k), y = __ldcg(Wcur + (size_t)j * n + k);
                    Wnext[(size_t)i * n + k] = rotated ? c * x + s * y : x;
                    Wnext[(size_t)j * n + k] = rotated ? c * y - s * x : y;
                    if (ip >= 0 && ip != i && ip != j) Wnext[(size_t)ip * n + k] = __ldcg(Wcur + (size_t)ip * n + k);
                    if (jp >= 0 && jp != i && jp != j) Wnext[(size_t)jp * n + k] = __ldcg(Wcur + (size_t)jp * n + k);
                    if (rotated) butterfly1(p.U + (size_t)i * n + k, p.U + (size_t)j * n + k, c, s);
                }
            }
            if (cta == 0) { SW_TICK(3) }
            if (t == p.P / 2 && tid == 0) p.dbg[cta] = (double)(gtimer() % 1000000000ull);
            grid_barrier(p.bar, G, epoch);
            if (t == p.P / 2 && tid == 0) p.dbg[1024 + cta] = (double)(gtimer() % 1000000000ull);
            if (cta == 0) { SW_TICK(6) } else { SW_TICK(7) }
        }
        // the caller's W must hold the final rotation: it is Wbuf[P & 1]
        if (p.P & 1)
            for (long long e = (long long)cta * SW_THREADS + tid; e < (long long)n * n; e += (long long)G * SW_THREADS)
                p.W[e] = __ldcg(p.W2 + e);
    }
    if (cta == 0 && tid == 0) {
        p.summary[0] = (double)accepted_total;
        p.summary[2] = (double)flags_or;
        // CTA 0: [4] prologue, [5] gather, [6] search, [7] tail, [8] barrier wait
        p.summary[4] = (double)tacc[0];
        p.summary[5] = (double)tacc[1];
        p.summary[6] = (double)tacc[2];
        p.summary[7] = (double)tacc[3];
        p.summary[8] = (double)tacc[6];
        for (int k = 0; k < 6; ++k) p.summary[10 + k] = (double)ls.prof[k];    // SM cycles inside the line search
        p.summary[20] = (double)ls.prof[6];      // block from super block + coefficients + pair / mask loads
        p.summary[21] = (double)ls.prof[7];      // result stores
    }
    if (cta == 1 && tid == 0) {
        // CTA 1 (a super-block CTA): [16] super-block orbits, [17] flat update (if any), [18] barrier wait
        p.summary[16] = (double)tacc[4];
        p.summary[17] = (double)tacc[5];
        p.summary[18] = (double)tacc[7];
    }
    if (G > 4 && cta == G - 3 && tid == 0) {
        // CTA G-3 (a flat-update CTA): [22] flat update, [23] barrier wait
        p.summary[22] = (double)tacc[5];
        p.summary[23] = (double)tacc[7];
    }
#undef SW_TICK
}

// Diagonal of the current 2-RDM through the pending rotation, local slabs only:
//   nn_k = sum_{a' local} W[k,a'] ( S[a',k,k,:] . W[k,:] );  one CTA per orbital in `inds`.
// gamma entries are written only when a0 == 0 so that shards combine by a sum.
__global__ void __launch_bounds__(256)
    deferred_diagonals_kernel(const double *__restrict__ gamma, const double *__restrict__ S,
                              const double *__restrict__ W, int n, int a0, int na,
                              const int32_t *__restrict__ inds, int m, double *__restrict__ diag) {
    __shared__ double part[8];
    const int t = blockIdx.x, k = inds[t], lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n1 = n, n2 = n1 * n1, n3 = n2 * n1;
    const double *wk = W + (size_t)k * n;
    double acc = 0.0;
    for (int al = warp; al < na; al += 8) {
        const double *row = S + al * n3 + k * n2 + k * n1;
        double dsum = 0.0;
        for (int d = lane; d < n; d += 32) dsum += row[d] * wk[d];
        dsum = warp_sum(dsum);
        acc += wk[a0 + al] * dsum;
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double nn = 0.0;
        for (int w = 0; w < 8; ++w) nn += part[w];
        const size_t ld = 2 * (size_t)n;
        diag[t] = a0 == 0 ? gamma[(2 * (size_t)k) * ld + 2 * k] : 0.0;
        diag[m + t] = a0 == 0 ? gamma[(2 * (size_t)k + 1) * ld + 2 * k + 1] : 0.0;
        diag[2 * m + t] = nn;
    }
}

__global__ void set_identity_kernel(double *W, int n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long long)n * n) W[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
}

__global__ void zero_offspin_if_kernel(double *gamma, int n, const double *summary) {
    if (summary[0] <= 0.0) return;
    const long long ld = 2LL * n, idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ld * ld) {
        const int r = (int)(idx / ld), c = (int)(idx % ld);
        if ((r & 1) != (c & 1)) gamma[idx] = 0.0;
    }
}

constexpr size_t WS_HEADER = 256 + 2 * sizeof(qio_b200_ls_result) + 96;
constexpr size_t WS_MAX_CTAS = 1024;     // one partial super block per CTA and parity; far above any SM count
static size_t sweep_ws_bytes(int n) {
    return WS_HEADER + sizeof(double) * (2 * (WS_MAX_CTAS + 1) * SB_DOUBLES + (size_t)n * n + 2 * WS_MAX_CTAS);
}

}  // namespace qio

// 3 = pipelined kernel without grid barriers (sweep3.cu) wherever the owned W columns fit in shared memory,
// 2 = barrier kernel of this file.  A caller forces one through qio_b200_sweep_options.engine.
static int sweep_engine(int n, int na) { return qio::sweep3_supported(n, na) ? 3 : 2; }

extern "C" int qio_b200_sweep_engine(int n, int na) { return sweep_engine(n, na); }

extern "C" size_t qio_b200_sweep_workspace_bytes(int n) {
    const size_t a = qio::sweep_ws_bytes(n), b = qio::sweep3_workspace_bytes(n);
    return a > b ? a : b;
}

extern "C" size_t qio_b200_mailbox_bytes(int nranks) {
    const size_t a = sizeof(double) * 2 * (size_t)(nranks > 0 ? nranks : 1) * qio::MB_STRIDE, b = qio::sweep3_mailbox_bytes(nranks);
    return a > b ? a : b;
}

extern "C" int qio_b200_sweep_reset(double *W, int n, void *stream) {
    using namespace qio;
    QIO_REQUIRE(W && n > 0, "sweep_reset: bad arguments");
    const long long total = (long long)n * n;
    set_identity_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(W, n);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_sweep_cycle(double *gamma, double *S, double *U, double *W, int n, int a0, int na,
                                    const int32_t *pairs, int P, const int32_t *inactive,
                                    const qio_b200_theta_tables *h_tables, qio_b200_ls_result *results,
                                    double *summary, void *workspace, const qio_b200_peers *h_peers,
                                    const qio_b200_sweep_options *h_opts, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n >= 2 && P >= 0, "sweep_cycle: bad sizes");
    QIO_REQUIRE(a0 >= 0 && na >= 0 && a0 + na <= n, "sweep_cycle: bad shard range");
    QIO_REQUIRE(gamma && U && W && inactive && h_tables && summary && workspace && (S || na == 0), "sweep_cycle: NULL pointer");
    QIO_REQUIRE(P == 0 || (pairs && results), "sweep_cycle: NULL pair list / results");
    QIO_REQUIRE(h_tables->n_coarse > 0 && h_tables->fine_stride > 0 && h_tables->fine_stride <= LS_MAX_FINE,
                "sweep_cycle: table sizes out of range");
    const int nranks = h_peers ? h_peers->nranks : 1;
    QIO_REQUIRE(nranks >= 1 && nranks <= 64, "sweep_cycle: bad rank count");
    QIO_REQUIRE(nranks == 1 || (h_peers->mailboxes && h_peers->rank >= 0 && h_peers->rank < nranks), "sweep_cycle: bad peers");
    QIO_REQUIRE(nranks > 1 || (a0 == 0 && na == n), "sweep_cycle: a partial shard needs peers");
    cudaStream_t st = as_stream(stream);
    qio_b200_sweep_options opts;
    memset(&opts, 0, sizeof(opts));
    if (h_opts) opts = *h_opts;
    QIO_REQUIRE(opts.engine == 0 || opts.engine == 2 || opts.engine == 3, "sweep_cycle: options.engine must be 0, 2 or 3");
    QIO_REQUIRE(nranks == 1 || (h_peers->epoch >= 1 && h_peers->epoch < (1u << (32 - LL_STEP_BITS))), "sweep_cycle: mailbox epoch must be in 1..4095");
    const int engine = opts.engine ? opts.engine : sweep_engine(n, na);
    if (engine == 3) {
        if (opts.timers)
            return sweep3_timed_launch(gamma, S, U, W, n, a0, na, pairs, P, inactive, h_tables, results, summary, workspace, h_peers, opts, st);
        return sweep3_launch(gamma, S, U, W, n, a0, na, pairs, P, inactive, h_tables, results, summary, workspace, h_peers, opts, st);
    }
    QIO_REQUIRE((long long)P + 1 < (1LL << LL_STEP_BITS), "sweep_cycle: too many pairs for the tag width");
    char *ws = reinterpret_cast<char *>(workspace);
    QIO_CUDA(cudaMemsetAsync(ws, 0, WS_HEADER, st));
    QIO_CUDA(cudaMemsetAsync(summary, 0, 32 * sizeof(double), st));
    SweepParams p;
    p.S = S;
    p.gamma = gamma;
    p.U = U;
    p.W = W;
    p.n = n;
    p.a0 = a0;
    p.na = na;
    p.pairs = pairs;
    p.P = P;
    p.inactive = inactive;
    p.tab = to_dev(*h_tables);
    p.results = results;
    p.bar = reinterpret_cast<unsigned *>(ws);
    p.cur = reinterpret_cast<qio_b200_ls_result *>(ws + 256);
    p.pbuf = reinterpret_cast<double *>(ws + WS_HEADER);
    p.W2 = p.pbuf + 2 * (WS_MAX_CTAS + 1) * SB_DOUBLES;
    p.dbg = p.W2 + (size_t)n * n;
    p.summary = summary;
    p.rank = h_peers ? h_peers->rank : 0;
    p.nranks = nranks;
    p.epoch = h_peers ? h_peers->epoch : 0u;
    p.blocks_out = opts.blocks_out;
    p.mailboxes = nranks > 1 ? reinterpret_cast<double *const *>(h_peers->mailboxes) : nullptr;
    int blocks_per_sm = 0;
    QIO_REQUIRE(n <= 4096, "sweep_cycle: n too large for the shared-memory W rows");
    const size_t dyn_smem = sizeof(double) * (4 * (size_t)n + (size_t)SW_WARPS * 256);
    QIO_CUDA(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    QIO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, sweep_kernel, SW_THREADS, dyn_smem));
    QIO_REQUIRE(blocks_per_sm >= 1, "sweep_cycle: kernel does not fit on an SM");
    const int grid = sm_count();
    QIO_REQUIRE(grid <= 1023, "sweep_cycle: workspace sized for at most 1023 CTAs");
    if (P > 0) {
        void *args[] = {&p};
        QIO_CUDA(cudaLaunchCooperativeKernel((const void *)sweep_kernel, dim3(grid), dim3(SW_THREADS), args, dyn_smem, st));
    }
    return QIO_B200_OK;
}

extern "C" int qio_b200_deferred_diagonals(const double *gamma, const double *S, const double *W, int n, int a0,
                                           int na, const int32_t *inds, int m, double *diag, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && m >= 0, "deferred_diagonals: bad sizes");
    QIO_REQUIRE(a0 >= 0 && na >= 0 && a0 + na <= n, "deferred_diagonals: bad shard range");
    if (m == 0) return QIO_B200_OK;
    QIO_REQUIRE(gamma && W && inds && diag && (S || na == 0), "deferred_diagonals: NULL pointer");
    deferred_diagonals_kernel<<<m, 256, 0, as_stream(stream)>>>(gamma, S, W, n, a0, na, inds, m, diag);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_zero_offspin_if_rotated(double *gamma, int n, const double *summary, void *stream) {
    using namespace qio;
    QIO_REQUIRE(gamma && summary && n > 0, "zero_offspin_if_rotated: bad arguments");
    const long long total = 4LL * n * n;
    zero_offspin_if_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(gamma, n, summary);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

// ---- peer-mapped mailboxes (CUDA IPC) ------------------------------------------------------------------
extern "C" int qio_b200_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle64) {
    using namespace qio;
    QIO_REQUIRE(ptr && handle64 && bytes > 0, "ipc_alloc: bad arguments");
    QIO_CUDA(cudaMalloc(ptr, bytes));
    QIO_CUDA(cudaMemset(*ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    QIO_CUDA(cudaIpcGetMemHandle(&h, *ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return QIO_B200_OK;
}

extern "C" int qio_b200_ipc_clear(void *ptr, size_t bytes, void *stream) {
    using namespace qio;
    QIO_REQUIRE(ptr && bytes > 0, "ipc_clear: bad arguments");
    QIO_CUDA(cudaMemsetAsync(ptr, 0, bytes, as_stream(stream)));
    return QIO_B200_OK;
}

extern "C" int qio_b200_ipc_open(const unsigned char *handle64, void **ptr) {
    using namespace qio;
    QIO_REQUIRE(ptr && handle64, "ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    QIO_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return QIO_B200_OK;
}

extern "C" int qio_b200_ipc_close(void *ptr) {
    using namespace qio;
    QIO_CUDA(cudaIpcCloseMemHandle(ptr));
    return QIO_B200_OK;
}

extern "C" int qio_b200_ipc_free(void *ptr) {
    using namespace qio;
    QIO_CUDA(cudaFree(ptr));
    return QIO_B200_OK;
}
