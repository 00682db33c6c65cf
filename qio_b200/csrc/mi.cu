// a-9 (EXTENSION, no reference counterpart)  two-orbital reduced density matrix, pair entropy.
//
// rho_ij is block diagonal in (N, S_z): blocks 1, 2, 2, 1, 4, 1, 2, 2, 1.  Every element is a linear form in the
// 24-double pair block (16 Gamma + 4 + 4 gamma entries) plus, for the 3- and 4-body expectation values that rdm1 / rdm2
// do not contain, an optional 9-number supplement per pair (mi_formulas.inc lists them; zeros when absent -> exact for
// two-electron states).  The table of linear forms is GENERATED (oracle/gen_mi_formulas.py) from a numerical operator
// expansion that is verified against brute-force partial traces, so no fermionic sign is hand-derived here.
// One thread per pair: assemble the 36 elements, symmetrise, eigenvalues by closed form (2 x 2) or cyclic Jacobi in
// registers (4 x 4), S_ij = -sum_{lambda > 0} lambda ln lambda.
#include "pair_math.cuh"
#include "mi_formulas.inc"

namespace qio {

static_assert(MI_NUM_SUPP == QIO_B200_MI_SUPP_DOUBLES, "header constant out of sync with the generated table");

__device__ __forceinline__ double xlogx(double p) { return p > 0.0 ? p * log(p) : 0.0; }

__device__ __forceinline__ void eig2(double a, double b, double d, double &l0, double &l1) {
    const double m = 0.5 * (a + d), h = 0.5 * (a - d), r = sqrt(h * h + b * b);
    l0 = m - r;
    l1 = m + r;
}

// cyclic Jacobi on a symmetric 4 x 4 matrix held in registers; eigenvalues in ascending order
__device__ __forceinline__ void eig4(double A[4][4], double lam[4]) {
#pragma unroll 1
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (p < q) off += A[p][q] * A[p][q];
        if (off < 1e-40) break;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q > p) {
                    const double apq = A[p][q];
                    if (fabs(apq) > 1e-300) {
                        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {      // columns p, q
                            const double akp = A[k][p], akq = A[k][q];
                            A[k][p] = c * akp - s * akq;
                            A[k][q] = s * akp + c * akq;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {      // rows p, q
                            const double apk = A[p][k], aqk = A[q][k];
                            A[p][k] = c * apk - s * aqk;
                            A[q][k] = s * apk + c * aqk;
                        }
                    }
                }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) lam[k] = A[k][k];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3 - a; ++b)
            if (lam[b] > lam[b + 1]) {
                const double t = lam[b];
                lam[b] = lam[b + 1];
                lam[b + 1] = t;
            }
}

__global__ void __launch_bounds__(128)
    pair_entropy_kernel(const double *__restrict__ blocks, const double *__restrict__ supp, int P,
                        double *__restrict__ s_pair, double *__restrict__ eig, int32_t *__restrict__ flags) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double rec[QIO_B200_BLOCK_DOUBLES], sp[MI_NUM_SUPP];
#pragma unroll
    for (int t = 0; t < QIO_B200_BLOCK_DOUBLES; ++t) rec[t] = blocks[(size_t)p * QIO_B200_BLOCK_DOUBLES + t];
#pragma unroll
    for (int t = 0; t < MI_NUM_SUPP; ++t) sp[t] = supp ? supp[(size_t)p * MI_NUM_SUPP + t] : 0.0;
    double el[MI_NUM_ELEMS];
    for (int e = 0; e < MI_NUM_ELEMS; ++e) {
        double v = 0.0;
        for (int t = MI_ELEM_START[e]; t < MI_ELEM_START[e + 1]; ++t) {
            const int kind = MI_TERM_KIND[t], a = MI_TERM_A[t];
            const double src = kind == 0 ? 1.0 : kind == 1 ? rec[a] : kind == 2 ? rec[a] - rec[MI_TERM_B[t]] : sp[a];
            v += (double)MI_TERM_COEF[t] * src;
        }
        el[e] = v;
    }
    double lam[16];
    int pos = 0, out = 0;
    for (int b = 0; b < 9; ++b) {
        const int m = MI_BLOCK_SIZE[b];
        if (m == 1) {
            lam[out] = el[pos];
        } else if (m == 2) {
            eig2(el[pos], 0.5 * (el[pos + 1] + el[pos + 2]), el[pos + 3], lam[out], lam[out + 1]);
        } else {
            double A[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) A[r][c] = 0.5 * (el[pos + r * 4 + c] + el[pos + c * 4 + r]);
            eig4(A, lam + out);
        }
        pos += m * m;
        out += m;
    }
    double S = 0.0, mn = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        S += xlogx(lam[k]);
        mn = fmin(mn, lam[k]);
        if (eig) eig[(size_t)p * 16 + k] = lam[k];
    }
    s_pair[p] = -S;
    if (flags) flags[p] = mn < -1e-6 ? QIO_B200_FLAG_WARN_NEGATIVE : 0;
}

// I[i, j] = I[j, i] = S_i + S_j - S_ij for the listed pairs (the matrix is zeroed first: zero diagonal, zero for pairs not listed)
__global__ void __launch_bounds__(256)
    mi_matrix_kernel(const double *__restrict__ s1, const double *__restrict__ s_pair, const int32_t *__restrict__ pairs, int P,
                     int n, double *__restrict__ I) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int i = pairs[2 * p], j = pairs[2 * p + 1];
    const double v = s1[i] + s1[j] - s_pair[p];
    I[(size_t)i * n + j] = v;
    I[(size_t)j * n + i] = v;
}

}  // namespace qio

extern "C" int qio_b200_mutual_information(const double *s1, const double *s_pair, const int32_t *pairs, int P, int n,
                                           double *I, void *stream) {
    using namespace qio;
    QIO_REQUIRE(P >= 0 && n > 0 && I, "mutual_information: bad arguments");
    QIO_CUDA(cudaMemsetAsync(I, 0, sizeof(double) * (size_t)n * n, as_stream(stream)));
    if (P == 0) return QIO_B200_OK;
    QIO_REQUIRE(s1 && s_pair && pairs, "mutual_information: NULL pointer");
    mi_matrix_kernel<<<(P + 255) / 256, 256, 0, as_stream(stream)>>>(s1, s_pair, pairs, P, n, I);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_pair_entropy(const double *blocks, const double *supp, int P, double *s_pair, double *eig,
                                     int32_t *flags, void *stream) {
    using namespace qio;
    QIO_REQUIRE(P >= 0, "pair_entropy: bad size");
    if (P == 0) return QIO_B200_OK;
    QIO_REQUIRE(blocks && s_pair, "pair_entropy: NULL pointer");
    pair_entropy_kernel<<<(P + 31) / 32, 32, 0, as_stream(stream)>>>(blocks, supp, P, s_pair, eig, flags);      // >= one CTA per SM from ~4,700 pairs on
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}
