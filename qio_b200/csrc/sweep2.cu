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
// So as soon as theta_t is known, CTA 0 can search pair t+1 while all other CTAs stream the update of pair t:
//   loop t:   [ CTA 0: B_{t+1} <- SB_t, theta_t; line search t+1 ]  ||  [ CTAs 1..: Givens update t on S, gamma, W, U ]
//             -- grid barrier --
//             all CTAs: partial SB_{t+1} from the updated state (|D|^2 n rows, |D| dot products each)
//             -- grid barrier --   CTA 0 sums the partials in a fixed order (and across ranks through peer mailboxes)
// The flush (W applied to axes 0 and 3 with two DMMA GEMM passes) happens once, when the caller wants Gamma back.
#include <string.h>

#include "linesearch.cuh"

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
    double *pbuf;              // [K + 1][SB_DOUBLES]
    qio_b200_ls_result *cur;   // [2] published result of the pair being applied (parity of the step)
    unsigned *bar;             // [0] arrival counter (monotone), [1] generation
    double *summary;           // [16]
    int rank, nranks;
    unsigned long long epoch;  // distinguishes launches in the mailbox tags
    double *const *mailboxes;  // [nranks] peer-mapped mailbox bases, layout [2][nranks][MB_STRIDE]; NULL if nranks == 1
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Monotone-counter grid barrier (all CTAs co-resident: cooperative launch).
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks, unsigned &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        __threadfence();
        const unsigned prev = atomicAdd(bar, 1u);
        if (prev == nblocks * epoch - 1) {
            __threadfence();
            atomicExch(bar + 1, epoch);
        } else {
            while (ld_acquire_u32(bar + 1) < epoch) {
            }
        }
        __threadfence();
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
__device__ __forceinline__ void update_S(double *__restrict__ S, int n, int na, int i, int j, double c, double s,
                                         long long gtid, long long gthreads) {
    const long long n1 = n, n2 = n1 * n1, n3 = n2 * n1;
    const int rowv = n / V;          // vector elements per d-row
    // axis 1: rows (a', {i,j}, cc, :) for all (a', cc), cc in {i, j} excluded (they belong to the 4-orbit)
    for (Counter3 k(gtid, gthreads, n, rowv); k.a < na; k.step()) {
        if (k.r == i || k.r == j) continue;
        double *px = S + k.a * n3 + (long long)i * n2 + k.r * n1 + (long long)k.v * V;
        double *py = S + k.a * n3 + (long long)j * n2 + k.r * n1 + (long long)k.v * V;
        if (V == 2) butterfly2(px, py, c, s); else butterfly1(px, py, c, s);
    }
    // axis 2: rows (a', b, {i,j}, :) for all (a', b), b in {i, j} excluded
    for (Counter3 k(gtid, gthreads, n, rowv); k.a < na; k.step()) {
        if (k.r == i || k.r == j) continue;
        double *px = S + k.a * n3 + k.r * n2 + (long long)i * n1 + (long long)k.v * V;
        double *py = S + k.a * n3 + k.r * n2 + (long long)j * n1 + (long long)k.v * V;
        if (V == 2) butterfly2(px, py, c, s); else butterfly1(px, py, c, s);
    }
    // the 4-row orbit (a', {i,j}, {i,j}, :): both axes
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

// Partial super block over this CTA's leading-axis slabs -> out[256] (entry ((p*4+q)*4+r)*4+s, 0 if unused).
// The <= 4 rows of W live in shared memory for the phase; each warp owns one (q, r) row position and walks the
// CTA's slabs two at a time so that the row loads of both are in flight together.
__device__ __forceinline__ void superblock_partial(const SweepParams &p, const Distinct &dd, int cta, int apc,
                                                   double *wrow /* smem [4][n] */, double *out) {
    const int n = p.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n1 = n, n2 = n1 * n1, n3 = n2 * n1;
    for (int e = threadIdx.x; e < dd.nd * n; e += SW_THREADS) {
        const int s = e / n, d = e - s * n;
        wrow[e] = __ldcg(p.W + (size_t)dd.D[s] * n + d);
    }
    __syncthreads();
    const int q = warp >> 2, r = warp & 3;
    const int pl = (lane >> 2) & 3, sl = lane & 3;          // lanes 0..15 own (p, s)
    double acc = 0.0;
    if (q < dd.nd && r < dd.nd) {
        const int first = cta * apc, last = min(first + apc, p.na);
        for (int al = first; al < last; al += 2) {
            const bool two = al + 1 < last;
            const double *row0 = p.S + al * n3 + dd.D[q] * n2 + dd.D[r] * n1;
            const double *row1 = two ? row0 + n3 : row0;
            double d0[4] = {0.0, 0.0, 0.0, 0.0}, d1[4] = {0.0, 0.0, 0.0, 0.0};
            for (int d = lane; d < n; d += 32) {
                const double v0 = __ldcg(row0 + d), v1 = __ldcg(row1 + d);
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    if (s < dd.nd) {
                        const double w = wrow[s * n + d];
                        d0[s] += v0 * w;
                        d1[s] += v1 * w;
                    }
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                d0[s] = warp_sum(d0[s]);
                d1[s] = warp_sum(d1[s]);
            }
            if (lane < 16 && pl < dd.nd && sl < dd.nd) {
                const double e0 = sl == 0 ? d0[0] : sl == 1 ? d0[1] : sl == 2 ? d0[2] : d0[3];
                const double e1 = sl == 0 ? d1[0] : sl == 1 ? d1[1] : sl == 2 ? d1[2] : d1[3];
                acc += wrow[pl * n + p.a0 + al] * e0;
                if (two) acc += wrow[pl * n + p.a0 + al + 1] * e1;
            }
        }
    }
    if (lane < 16) out[((pl * 4 + q) * 4 + r) * 4 + sl] = acc;
    __syncthreads();
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
__device__ __forceinline__ void block_from_superblock(const double *sb /*[288] smem*/, const Target &t, double *blk) {
    const int o = threadIdx.x;
    if (o < 16) {
        const int al = (o >> 3) & 1, be = (o >> 2) & 1, ga = (o >> 1) & 1, de = o & 1;
        double acc = 0.0;
        for (int a = 0; a < t.cnt[al]; ++a)
            for (int b = 0; b < t.cnt[be]; ++b)
                for (int c = 0; c < t.cnt[ga]; ++c)
                    for (int d = 0; d < t.cnt[de]; ++d)
                        acc += t.coef[al][a] * t.coef[be][b] * t.coef[ga][c] * t.coef[de][d] *
                               sb[((t.idx[al][a] * 4 + t.idx[be][b]) * 4 + t.idx[ga][c]) * 4 + t.idx[de][d]];
        blk[o] = acc;
    } else if (o < 24) {
        const int u = o - 16, spin = u >> 2, al = (u >> 1) & 1, be = u & 1;
        double acc = 0.0;
        for (int a = 0; a < t.cnt[al]; ++a)
            for (int b = 0; b < t.cnt[be]; ++b)
                acc += t.coef[al][a] * t.coef[be][b] * sb[256 + spin * 16 + t.idx[al][a] * 4 + t.idx[be][b]];
        blk[o] = acc;
    }
}

__global__ void __launch_bounds__(SW_THREADS, 1) sweep_kernel(SweepParams p) {
    __shared__ LineSearchSmemT<SW_THREADS> ls;
    __shared__ double blk[QIO_B200_BLOCK_DOUBLES];
    __shared__ double sb[SB_DOUBLES];
    __shared__ PairCoef pc;
    __shared__ double2 coarse_cs_sm[320];
    __shared__ int32_t fine_len_sm[320];
    extern __shared__ double wrow_sm[];       // [4][n] rows of W for the super-block phase
    const int n = p.n, tid = threadIdx.x;
    const int G = gridDim.x, cta = blockIdx.x;
    const int apc = max(4, (p.na + G - 1) / G);                   // leading-axis slabs per CTA in the super-block phase
    const int K = (p.na + apc - 1) / apc;                         // CTAs with super-block work (K <= G)
    unsigned epoch = 0;
    int accepted_total = 0, flags_or = 0;
    unsigned long long tacc[6] = {0, 0, 0, 0, 0, 0}, t0 = 0, t1;
    const bool timing = (cta == 0 && tid == 0);
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

    auto pair_at = [&](int t, int &i, int &j) {
        const int tt = t < p.P ? t : p.P - 1;
        i = p.pairs[2 * tt];
        j = p.pairs[2 * tt + 1];
    };

    // super block of step t (orbitals of pairs t and t+1) from the current state
    auto phase_superblock = [&](int t) {
        int i, j, i2, j2;
        pair_at(t, i, j);
        pair_at(t + 1, i2, j2);
        const Distinct dd = make_distinct(i, j, i2, j2);
        if (cta < K) superblock_partial(p, dd, cta, apc, wrow_sm, p.pbuf + (size_t)cta * SB_DOUBLES);
        if (cta == (K < G ? K : 0) && tid < 32) {      // gamma entries over D (rank-replicated data)
            const int spin = tid >> 4, pp = (tid >> 2) & 3, qq = tid & 3;
            double v = 0.0;
            if (pp < dd.nd && qq < dd.nd)
                v = __ldcg(p.gamma + (2 * (size_t)dd.D[pp] + spin) * (2 * (size_t)n) + 2 * dd.D[qq] + spin);
            p.pbuf[(size_t)K * SB_DOUBLES + 256 + tid] = v;
        }
    };

    // CTA 0: sum the partials in a fixed order (CTAs, then ranks) into shared memory
    auto gather_superblock = [&](int t) {
        if (tid < 256) {
            double a8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};      // fixed pattern: deterministic on every rank
            int c = 0;
            for (; c + 8 <= K; c += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) a8[u] += __ldcg(p.pbuf + (size_t)(c + u) * SB_DOUBLES + tid);
            }
            for (int u = 0; c < K; ++c, ++u) a8[u] += __ldcg(p.pbuf + (size_t)c * SB_DOUBLES + tid);
            sb[tid] = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
        } else if (tid < 288) {
            sb[tid] = __ldcg(p.pbuf + (size_t)K * SB_DOUBLES + tid);
        }
        if (p.nranks > 1) {
            // Every 16-byte word carries (value, tag): one NVLink store per element, no fence, no separate flag --
            // the receiver spins on the tag of each word it needs (the "LL" idea).  Slots alternate with the step
            // parity; a slot is rewritten two steps later, after every reader has moved on (see the header comment).
            const int par = t & 1;
            const unsigned long long tag = (p.epoch << 32) | (unsigned long long)(t + 1);
            __syncthreads();
            if (tid < 256) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(sb[tid]);
                for (int dst = 0; dst < p.nranks; ++dst) {
                    unsigned long long *w = reinterpret_cast<unsigned long long *>(
                        p.mailboxes[dst] + ((size_t)par * p.nranks + p.rank) * MB_STRIDE) + 2 * tid;
                    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"(bits), "l"(tag) : "memory");
                }
                double acc = 0.0;
                for (int r = 0; r < p.nranks; ++r) {
                    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(
                        p.mailboxes[p.rank] + ((size_t)par * p.nranks + r) * MB_STRIDE) + 2 * tid;
                    unsigned long long v, g;
                    do {
                        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(g) : "l"(w) : "memory");
                    } while (g != tag);
                    acc += __longlong_as_double((long long)v);
                }
                sb[tid] = acc;
            }
        }
        __syncthreads();
    };

    // CTA 0: line search of pair `t` whose block is derived from the super block in smem
    auto search_pair = [&](int t, const Target &tg) {
        block_from_superblock(sb, tg, blk);
        __syncthreads();
        if (tid == 0) make_coef(blk, pc);
        __syncthreads();
        int i, j;
        pair_at(t, i, j);
        qio_b200_ls_result res;
        cta_linesearch_t<SW_THREADS, PairCoef>(pc, p.inactive[i] != 0, p.inactive[j] != 0, tab, ls, res);
        if (tid == 0) {
            p.results[t] = res;
            p.cur[t & 1] = res;
        }
        accepted_total += res.accepted;
        flags_or |= res.flags;
    };

    if (p.P > 0) {
        // ---- prologue: super block 0, search pair 0 -------------------------------------------------
        if (timing) t0 = gtimer();
        phase_superblock(0);
        grid_barrier(p.bar, G, epoch);
        if (cta == 0) {
            gather_superblock(0);
            int i, j, i2, j2;
            pair_at(0, i, j);
            pair_at(1, i2, j2);
            const Distinct dd = make_distinct(i, j, i2, j2);
            search_pair(0, make_target(dd, 0, 1, false, 1.0, 0.0));
        }
        grid_barrier(p.bar, G, epoch);
        SW_TICK(0)

        for (int t = 0; t < p.P; ++t) {
            int i, j, i2, j2;
            pair_at(t, i, j);
            pair_at(t + 1, i2, j2);
            const bool rotated = __ldcg(&p.cur[t & 1].accepted) != 0;
            const double c = __ldcg(&p.cur[t & 1].cos_theta), s = __ldcg(&p.cur[t & 1].sin_theta);

            // ---- CTA 0: search pair t+1 from SB_t and theta_t ---------------------------------------
            if (cta == 0 && t + 1 < p.P) {
                const Distinct dd = make_distinct(i, j, i2, j2);
                search_pair(t + 1, make_target(dd, 2, 3, rotated, c, s));
                SW_TICK(1)
            }
            // ---- the other CTAs (or the only one): apply rotation t ----------------------------------
            if ((G == 1 || cta != 0) && rotated) {
                if ((n & 1) == 0)
                    update_S<2>(p.S, n, p.na, i, j, c, s, gtid, gthreads);
                else
                    update_S<1>(p.S, n, p.na, i, j, c, s, gtid, gthreads);
            }
            if (rotated && cta == cta_gamma) {
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
            }
            if (rotated && cta == cta_wu) {
                // W and U rows i, j  (U <- V U, jacobi.py:230)
                for (int k = tid; k < 2 * n; k += SW_THREADS) {
                    double *M = (k < n) ? p.W : p.U;
                    const int col = (k < n) ? k : k - n;
                    butterfly1(M + (size_t)i * n + col, M + (size_t)j * n + col, c, s);
                }
            }
            SW_TICK(2)
            grid_barrier(p.bar, G, epoch);
            SW_TICK(3)
            if (t + 1 < p.P) {
                phase_superblock(t + 1);
                grid_barrier(p.bar, G, epoch);
                SW_TICK(4)
                if (cta == 0) gather_superblock(t + 1);
                SW_TICK(5)
            }
        }
    }
    if (cta == 0 && tid == 0) {
        p.summary[0] = (double)accepted_total;
        p.summary[2] = (double)flags_or;
        for (int k = 0; k < 6; ++k) p.summary[4 + k] = (double)tacc[k];
        for (int k = 0; k < 6; ++k) p.summary[10 + k] = (double)ls.prof[k];    // SM cycles inside the line search
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

static size_t sweep_ws_bytes(int n) {
    (void)n;
    const size_t K = 1024;       // one partial super block per CTA; far above any SM count
    return 256 + 2 * sizeof(qio_b200_ls_result) + 96 + sizeof(double) * (K + 1) * SB_DOUBLES;
}

}  // namespace qio

extern "C" size_t qio_b200_sweep_workspace_bytes(int n) { return qio::sweep_ws_bytes(n); }

extern "C" size_t qio_b200_mailbox_bytes(int nranks) {
    return sizeof(double) * 2 * (size_t)(nranks > 0 ? nranks : 1) * qio::MB_STRIDE;
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
                                    double *summary, void *workspace, const qio_b200_peers *h_peers, void *stream) {
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
    char *ws = reinterpret_cast<char *>(workspace);
    QIO_CUDA(cudaMemsetAsync(ws, 0, 256 + 2 * sizeof(qio_b200_ls_result) + 96, st));
    QIO_CUDA(cudaMemsetAsync(summary, 0, 16 * sizeof(double), st));
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
    p.pbuf = reinterpret_cast<double *>(ws + 256 + 2 * sizeof(qio_b200_ls_result) + 96);
    p.summary = summary;
    p.rank = h_peers ? h_peers->rank : 0;
    p.nranks = nranks;
    p.epoch = h_peers ? (unsigned long long)h_peers->epoch : 0ull;
    p.mailboxes = nranks > 1 ? reinterpret_cast<double *const *>(h_peers->mailboxes) : nullptr;
    int blocks_per_sm = 0;
    QIO_REQUIRE(n <= 4096, "sweep_cycle: n too large for the shared-memory W rows");
    if (sizeof(double) * 4 * (size_t)n > 32768)
        QIO_CUDA(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 4 * n)));
    QIO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, sweep_kernel, SW_THREADS, sizeof(double) * 4 * (size_t)n));
    QIO_REQUIRE(blocks_per_sm >= 1, "sweep_cycle: kernel does not fit on an SM");
    const int grid = sm_count();
    QIO_REQUIRE(grid <= 1023, "sweep_cycle: workspace sized for at most 1023 CTAs");
    if (P > 0) {
        void *args[] = {&p};
        QIO_CUDA(cudaLaunchCooperativeKernel((const void *)sweep_kernel, dim3(grid), dim3(SW_THREADS), args,
                                             sizeof(double) * 4 * (size_t)n, st));
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
