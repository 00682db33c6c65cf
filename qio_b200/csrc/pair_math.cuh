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

// Entropy of the rotated orbital with coefficients (r0, r1) on (i, j): qio/grad/jacobi.py:43-57.
__device__ __forceinline__ double rotated_orbital_entropy(const PairBlock &b, double r0, double r1, int &flags) {
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
__device__ __forceinline__ double pair_cost(const PairBlock &b, double c, double s, bool i_in, bool j_in, int &flags) {
    double total = 0.0;
    if (i_in) total = __dadd_rn(total, rotated_orbital_entropy(b, c, s, flags));
    if (j_in) total = __dadd_rn(total, rotated_orbital_entropy(b, -s, c, flags));
    return total;
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

__device__ __forceinline__ MinIdx warp_min(MinIdx x) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        MinIdx y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, off);
        y.idx = __shfl_xor_sync(0xffffffffu, x.idx, off);
        y.v2 = __shfl_xor_sync(0xffffffffu, x.v2, off);
        x = merge(x, y);
    }
    return x;
}

}  // namespace qio
