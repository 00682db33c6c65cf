// Device-side arithmetic shared by the pair kernels and the Jacobi sweep.
//
// Everything that feeds a DISCRETE decision of the reference (the rotated occupations, the
// one-orbital spectrum, the entropy sum and the angle scans of qio/grad/jacobi.py:15-60,117-158)
// is evaluated in the reference's own operation order with explicitly rounded adds/multiplies
// (__dadd_rn / __dmul_rn are never contracted into FMAs), and cos/sin come from host-built numpy
// tables, so the only difference to the CPU path is the last-bit behaviour of log().
#pragma once
#include "common.cuh"

namespace qio {

// A pair block in registers: G[m*8+n*4+p*2+q], gu[m*2+n], gd[m*2+n]  (I = (i, j)).
struct PairBlock {
    double G[16];
    double gu[4];
    double gd[4];
};

__device__ __forceinline__ void load_block(const double *__restrict__ rec, PairBlock &b) {
#pragma unroll
    for (int t = 0; t < 16; ++t) b.G[t] = rec[t];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        b.gu[t] = rec[16 + t];
        b.gd[t] = rec[20 + t];
    }
}

// Offset of block entry t (0..23) of pair (i, j) inside gamma (t >= 16) or Gamma (t < 16).
__device__ __forceinline__ size_t block_entry_offset(int t, int i, int j, int n) {
    if (t < 16) {
        const size_t a = (t & 8) ? j : i, b = (t & 4) ? j : i, c = (t & 2) ? j : i, d = (t & 1) ? j : i;
        return ((a * n + b) * n + c) * n + d;
    }
    const int u = t - 16, spin = u >> 2;
    const size_t r = (u & 2) ? j : i, c = (u & 1) ? j : i;
    return (2 * r + spin) * (size_t)(2 * n) + 2 * c + spin;
}

// -sum_{p>0} p ln p with numpy's small-array order ((0 + t0) + t1) + ...  (qio/entropy.py:21-22),
// plus the warn / raise conditions of qio/entropy.py:15-20 OR-ed into `flags`.
__device__ __forceinline__ double masked_entropy4(double s0, double s1, double s2, double s3, int &flags) {
    const double sp[4] = {s0, s1, s2, s3};
    double acc = 0.0;
    bool any_neg = false, big_neg = false, gt1 = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double p = sp[r];
        any_neg |= (p < 0.0);
        big_neg |= (p < -1e-6);
        gt1 |= (p > 1.0);
        if (p > 0.0) acc = __dadd_rn(acc, __dmul_rn(p, log(p)));
    }
    if (any_neg) {
        if (big_neg) flags |= QIO_B200_FLAG_WARN_NEGATIVE;
    } else if (gt1) {
        flags |= QIO_B200_FLAG_RAISE_GT1;
    }
    return -acc;
}

// Same record held in shared memory (broadcast reads) for register-tight kernels.
struct PairBlockRef {
    const double *G, *gu, *gd;
};

// Entropy of the rotated orbital with coefficients (r0, r1) on (i, j): qio/grad/jacobi.py:43-57.
template <typename Block>
__device__ __forceinline__ double rotated_orbital_entropy(const Block &b, double r0, double r1, int &flags) {
    const double r[2] = {r0, r1};
    double nu = 0.0, nd = 0.0, nn = 0.0;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const double rr = __dmul_rn(r[m], r[n]);
            nu = __dadd_rn(nu, __dmul_rn(rr, b.gu[m * 2 + n]));
            nd = __dadd_rn(nd, __dmul_rn(rr, b.gd[m * 2 + n]));
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    nn = __dadd_rn(nn, __dmul_rn(__dmul_rn(__dmul_rn(rr, r[p]), r[q]), b.G[m * 8 + n * 4 + p * 2 + q]));
        }
    const double s0 = __dadd_rn(__dsub_rn(__dsub_rn(1.0, nu), nd), nn);
    return masked_entropy4(s0, __dsub_rn(nu, nn), __dsub_rn(nd, nn), nn, flags);
}

// jacobi_cost(theta, i, j, ...) given cos/sin: S(i') if i inactive, plus S(j') if j inactive.
template <typename Block>
__device__ __forceinline__ double pair_cost(const Block &b, double c, double s, bool i_in, bool j_in, int &flags, int * /*sink*/ = nullptr) {
    double total = 0.0;
    if (i_in) total = __dadd_rn(total, rotated_orbital_entropy(b, c, s, flags));
    if (j_in) total = __dadd_rn(total, rotated_orbital_entropy(b, -s, c, flags));
    return total;
}

// ---- fast evaluation ------------------------------------------------------------------------------
// The rotated occupations are polynomials in (c, s):  nn_i = A0 c^4 + A1 c^3 s + A2 c^2 s^2 + A3 c s^3 + A4 s^4
// with A_k = sum of the block entries that carry k indices equal to j, and for orbital j (row (-s, c)) the
// same coefficients with c^2 <-> s^2, cs -> -cs.  Five coefficients per pair replace the 16-term sum per angle
// (depth ~3 FMAs instead of a 16-deep dependent add chain), and the entropy uses a branch-free log so that the
// 8 logarithms of one candidate overlap.  Values agree with the strict-order path to a few 1e-16; the angle a
// line search returns can differ only when two candidates are closer than that (reported as the margins).
struct PairCoef {
    double A[5];
    double gu[3];   // g_ii, g_ij + g_ji, g_jj  (up)
    double gd[3];   // (down)
    int sym;        // up and down blocks are bitwise equal (singlet): nu == nd, two spectrum entries coincide
};

__device__ __forceinline__ void make_coef(const double *rec /* 24-double block */, PairCoef &pc) {
    pc.A[0] = rec[0];
    pc.A[1] = ((rec[1] + rec[2]) + rec[4]) + rec[8];
    pc.A[2] = ((((rec[3] + rec[5]) + rec[6]) + rec[9]) + rec[10]) + rec[12];
    pc.A[3] = ((rec[7] + rec[11]) + rec[13]) + rec[14];
    pc.A[4] = rec[15];
    pc.gu[0] = rec[16];
    pc.gu[1] = rec[17] + rec[18];
    pc.gu[2] = rec[19];
    pc.gd[0] = rec[20];
    pc.gd[1] = rec[21] + rec[22];
    pc.gd[2] = rec[23];
    pc.sym = (pc.gu[0] == pc.gd[0] && pc.gu[1] == pc.gd[1] && pc.gu[2] == pc.gd[2]) ? 1 : 0;
}

// ln(x) for positive normal x, fdlibm's algorithm (error < 1 ulp; 1 ulp max against glibc on 2e7 samples),
// written N-wide: every step is applied to all N arguments before the next step, so that the N dependency
// chains (each ~45 operations deep) are interleaved in the instruction stream instead of running back to back.
// The division f / (2 + f) is an approximate single-precision reciprocal refined by two Newton steps.
template <int N>
__device__ __forceinline__ void log_pos_n(const double (&x)[N], double (&y)[N]) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    double f[N], d[N], r[N], dk[N];
#pragma unroll
    for (int e = 0; e < N; ++e) {
        int hi = __double2hiint(x[e]);
        const int lo = __double2loint(x[e]);
        int k = (hi >> 20) - 1023;
        hi &= 0x000fffff;
        const int adj = (hi + 0x95f64) & 0x100000;      // mantissa above sqrt(2): halve it, bump the exponent
        hi |= (adj ^ 0x3ff00000);
        k += adj >> 20;
        dk[e] = (double)k;
        f[e] = __hiloint2double(hi, lo) - 1.0;
        d[e] = 2.0 + f[e];
    }
#pragma unroll
    for (int e = 0; e < N; ++e) {
        float rf;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)d[e]));   // d in [1.7, 2.42]: ~23 good bits
        r[e] = (double)rf;
    }
    double t[N];
#pragma unroll
    for (int e = 0; e < N; ++e) t[e] = fma(-d[e], r[e], 1.0);
#pragma unroll
    for (int e = 0; e < N; ++e) r[e] = fma(r[e], t[e], r[e]);
#pragma unroll
    for (int e = 0; e < N; ++e) t[e] = fma(-d[e], r[e], 1.0);
#pragma unroll
    for (int e = 0; e < N; ++e) r[e] = fma(r[e], t[e], r[e]);
    double s[N], z[N], w[N], t1[N], t2[N], hfsq[N];
#pragma unroll
    for (int e = 0; e < N; ++e) {
        s[e] = f[e] * r[e];
        hfsq[e] = 0.5 * f[e] * f[e];
    }
#pragma unroll
    for (int e = 0; e < N; ++e) z[e] = s[e] * s[e];
#pragma unroll
    for (int e = 0; e < N; ++e) w[e] = z[e] * z[e];
#pragma unroll
    for (int e = 0; e < N; ++e) {
        t1[e] = fma(w[e], Lg6, Lg4);
        t2[e] = fma(w[e], Lg7, Lg5);
    }
#pragma unroll
    for (int e = 0; e < N; ++e) {
        t1[e] = fma(w[e], t1[e], Lg2);
        t2[e] = fma(w[e], t2[e], Lg3);
    }
#pragma unroll
    for (int e = 0; e < N; ++e) {
        t1[e] = w[e] * t1[e];
        t2[e] = fma(w[e], t2[e], Lg1);
    }
#pragma unroll
    for (int e = 0; e < N; ++e) t2[e] = fma(z[e], t2[e], t1[e]);          // R
#pragma unroll
    for (int e = 0; e < N; ++e) t1[e] = fma(s[e], hfsq[e] + t2[e], dk[e] * ln2_lo);
#pragma unroll
    for (int e = 0; e < N; ++e) y[e] = dk[e] * ln2_hi - ((hfsq[e] - t1[e]) - f[e]);
}

__device__ __forceinline__ double masked_entropy4_fast(double s0, double s1, double s2, double s3, int &flags) {
    const double sp[4] = {s0, s1, s2, s3};
    double arg[4], lg[4];
    bool any_neg = false, big_neg = false, gt1 = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double p = sp[r];
        any_neg |= (p < 0.0);
        big_neg |= (p < -1e-6);
        gt1 |= (p > 1.0);
        arg[r] = fmax(p, 1e-300);       // |p ln p| < 1e-297 below the clamp: contributes 0
    }
    log_pos_n<4>(arg, lg);
    double term[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) term[r] = (sp[r] > 1e-300) ? sp[r] * lg[r] : 0.0;
    flags |= any_neg ? (big_neg ? QIO_B200_FLAG_WARN_NEGATIVE : 0) : (gt1 ? QIO_B200_FLAG_RAISE_GT1 : 0);
    return -(((term[0] + term[1]) + term[2]) + term[3]);
}

// Deliberately NOT inlined: the line search calls it from three places and the persistent sweep kernel is
// instruction-fetch bound when the loop body outgrows the instruction cache.
// Spin-symmetric case: the spectrum is [s0, s1, s1, s3]; three logarithms, the same sum ((t0 + t1) + t1) + t3.
__device__ __forceinline__ double masked_entropy3_sym(double s0, double s1, double s3, int &flags) {
    const double sp[3] = {s0, s1, s3};
    double arg[3], lg[3];
    bool any_neg = false, big_neg = false, gt1 = false;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double p = sp[r];
        any_neg |= (p < 0.0);
        big_neg |= (p < -1e-6);
        gt1 |= (p > 1.0);
        arg[r] = fmax(p, 1e-300);
    }
    log_pos_n<3>(arg, lg);
    double term[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) term[r] = (sp[r] > 1e-300) ? sp[r] * lg[r] : 0.0;
    flags |= any_neg ? (big_neg ? QIO_B200_FLAG_WARN_NEGATIVE : 0) : (gt1 ? QIO_B200_FLAG_RAISE_GT1 : 0);
    return -(((term[0] + term[1]) + term[1]) + term[2]);
}

// Two variants (spin-symmetric: 2 x 3 logarithms; general: 2 x 4), both inlined into the line search (two sites).
// Flags (warn / raise conditions, rare) go straight to `sink` (a shared- or global-memory word): a by-reference int would
// live in local memory and cost every call a load and a store there.
__device__ __forceinline__ double pair_cost_sym(double A0, double A1, double A2, double A3, double A4, double g0, double g1,
                                                    double g2, double c, double s, bool i_in, bool j_in, int *sink) {
    const double c2 = c * c, s2 = s * s, cs = c * s, c2s2 = c2 * s2;
    const double nn_i = c2 * (A0 * c2 + A1 * cs) + A2 * c2s2 + s2 * (A3 * cs + A4 * s2);
    const double nn_j = s2 * (A0 * s2 - A1 * cs) + A2 * c2s2 + c2 * (A4 * c2 - A3 * cs);
    const double nu_i = g0 * c2 + g1 * cs + g2 * s2, nu_j = g0 * s2 - g1 * cs + g2 * c2;
    int fi = 0, fj = 0;       // nd == nu bit for bit
    const double Si = masked_entropy3_sym(1.0 - nu_i - nu_i + nn_i, nu_i - nn_i, nn_i, fi);
    const double Sj = masked_entropy3_sym(1.0 - nu_j - nu_j + nn_j, nu_j - nn_j, nn_j, fj);
    const int f = (i_in ? fi : 0) | (j_in ? fj : 0);
    if (f) atomicOr(sink, f);
    return (i_in ? Si : 0.0) + (j_in ? Sj : 0.0);
}

__device__ __forceinline__ double pair_cost_gen(const PairCoef &pc, double c, double s, bool i_in, bool j_in, int *sink) {
    const double c2 = c * c, s2 = s * s, cs = c * s, c2s2 = c2 * s2;
    const double nn_i = c2 * (pc.A[0] * c2 + pc.A[1] * cs) + pc.A[2] * c2s2 + s2 * (pc.A[3] * cs + pc.A[4] * s2);
    const double nn_j = s2 * (pc.A[0] * s2 - pc.A[1] * cs) + pc.A[2] * c2s2 + c2 * (pc.A[4] * c2 - pc.A[3] * cs);
    const double nu_i = pc.gu[0] * c2 + pc.gu[1] * cs + pc.gu[2] * s2, nu_j = pc.gu[0] * s2 - pc.gu[1] * cs + pc.gu[2] * c2;
    const double nd_i = pc.gd[0] * c2 + pc.gd[1] * cs + pc.gd[2] * s2, nd_j = pc.gd[0] * s2 - pc.gd[1] * cs + pc.gd[2] * c2;
    int fi = 0, fj = 0;
    const double Si = masked_entropy4_fast(1.0 - nu_i - nd_i + nn_i, nu_i - nn_i, nd_i - nn_i, nn_i, fi);
    const double Sj = masked_entropy4_fast(1.0 - nu_j - nd_j + nn_j, nu_j - nn_j, nd_j - nn_j, nn_j, fj);
    const int f = (i_in ? fi : 0) | (j_in ? fj : 0);
    if (f) atomicOr(sink, f);
    return (i_in ? Si : 0.0) + (j_in ? Sj : 0.0);
}

__device__ __forceinline__ double pair_cost(const PairCoef &pc, double c, double s, bool i_in, bool j_in, int & /*flags*/, int *sink) {
    if (pc.sym) return pair_cost_sym(pc.A[0], pc.A[1], pc.A[2], pc.A[3], pc.A[4], pc.gu[0], pc.gu[1], pc.gu[2], c, s, i_in, j_in, sink);
    return pair_cost_gen(pc, c, s, i_in, j_in, sink);
}

// Analytic d/dtheta and d2/dtheta2 of the pair cost at theta = 0 (qio/grad/gradient.py:13-120),
// literal formulas including the duplicated Gamma[i,j,j,j] term of gradient.py:40.
__device__ __forceinline__ void pair_grad_hess(const PairBlock &b, bool i_in, bool j_in, double &g1, double &g2) {
    // shorthand: G(m,n,p,q) with 0 = i, 1 = j
#define QG(m, n, p, q) b.G[(m) * 8 + (n) * 4 + (p) * 2 + (q)]
    const double dgi = 2.0 * b.gu[1];                // 2*gamma[2i,2j]; down copies up (gradient.py:18-21)
    const double diff = b.gu[0] - b.gu[3];           // gamma[2i,2i] - gamma[2j,2j]
    const double dG_i = QG(1, 0, 0, 0) + QG(0, 1, 0, 0) + QG(0, 0, 1, 0) + QG(0, 0, 0, 1);
    const double dG_j = -QG(0, 1, 1, 1) - QG(0, 1, 1, 1) - QG(1, 1, 0, 1) - QG(1, 1, 1, 0);
    const double mixed = 2.0 * (QG(1, 1, 0, 0) + QG(1, 0, 1, 0) + QG(1, 0, 0, 1) + QG(0, 1, 1, 0) + QG(0, 1, 0, 1) + QG(0, 0, 1, 1));
    g1 = 0.0;
    g2 = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (!(k == 0 ? i_in : j_in)) continue;
        const double nu = k == 0 ? b.gu[0] : b.gu[3];
        const double nd = k == 0 ? b.gd[0] : b.gd[3];
        const double nn = k == 0 ? QG(0, 0, 0, 0) : QG(1, 1, 1, 1);
        const double du = k == 0 ? dgi : -dgi, dd = du;
        const double ddu = k == 0 ? -2.0 * diff : 2.0 * diff, ddd = ddu;
        const double dG = k == 0 ? dG_i : dG_j;
        const double ddG = -4.0 * nn + mixed;
        const double sp[4] = {1.0 - nu - nd + nn, nu - nn, nd - nn, nn};
        const double dsp[4] = {-du - dd + dG, du - dG, dd - dG, dG};
        const double ddsp[4] = {-ddu - ddd + ddG, ddu - ddG, ddd - ddG, ddG};
        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (sp[r] > 0.0) {
                const double lp = log(sp[r]);
                a1 += lp * dsp[r];
                a2 += (1.0 / sp[r]) * (dsp[r] * dsp[r]);
                a3 += lp * ddsp[r];
            }
        }
        g1 += -a1;
        g2 += -a2 - a3;
    }
#undef QG
}

// ---- warp / block reductions used by the line search ------------------------------------------

struct MinIdx {
    double v;   // best value (NaN mapped to +inf by the caller)
    int idx;    // its first index
    double v2;  // second best value among OTHER indices
};

__device__ __forceinline__ MinIdx merge(const MinIdx &a, const MinIdx &b) {
    MinIdx r;
    const bool a_first = (a.v < b.v) || (a.v == b.v && a.idx < b.idx);
    if (a_first) {
        r.v = a.v;
        r.idx = a.idx;
        r.v2 = fmin(a.v2, b.v);
    } else {
        r.v = b.v;
        r.idx = b.idx;
        r.v2 = fmin(b.v2, a.v);
    }
    return r;
}

// Order-preserving map double -> unsigned 64-bit (all finite values and +-inf; NaN is mapped by the callers beforehand).
__device__ __forceinline__ unsigned long long dkey(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
// Warp-wide minimum of a 64-bit key with two hardware reductions (REDUX) instead of five shuffle rounds.
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    return ((unsigned long long)mhi << 32) | mlo;
}

// Same result as a merge() tree over the lanes: first index of the minimum, second best value among the OTHER indices.
// SECOND = false leaves the second-best value out (+inf): two hardware reductions fewer.
template <bool SECOND = true>
__device__ __forceinline__ MinIdx warp_min(MinIdx x) {
    const unsigned long long key = dkey(x.v), kmin = warp_min_u64(key);
    MinIdx r;
    r.v = dunkey(kmin);
    r.idx = (int)__reduce_min_sync(0xffffffffu, key == kmin ? (unsigned)x.idx : 0x7fffffffu);
    r.v2 = SECOND ? dunkey(warp_min_u64(dkey(x.idx == r.idx ? x.v2 : x.v))) : INFINITY;
    return r;
}

}  // namespace qio
