// a-2 / a-3  one-orbital spectra and entropies (reference qio/entropy.py:4-47, jacobi.py:313-318).
#include "pair_math.cuh"

namespace qio {

__global__ void orbital_diagonals_kernel(const double *__restrict__ gamma, const double *__restrict__ Gamma,
                                         int n, int a0, int na, const int32_t *__restrict__ inds, int m,
                                         double *__restrict__ diag) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int k = inds[t];
    const size_t ld = 2 * (size_t)n;
    diag[t] = gamma[(2 * (size_t)k) * ld + 2 * k];
    diag[m + t] = gamma[(2 * (size_t)k + 1) * ld + 2 * k + 1];
    double nn = 0.0;
    if (k >= a0 && k < a0 + na) {
        const size_t n1 = n, kk = k, kl = k - a0;
        nn = Gamma[((kl * n1 + kk) * n1 + kk) * n1 + kk];
    }
    diag[2 * m + t] = nn;
}

constexpr int ENT_THREADS = 256;

// Single CTA: the result feeds a host-side convergence test, so the reduction order is fixed.
__global__ void __launch_bounds__(ENT_THREADS)
    orbital_entropies_kernel(const double *__restrict__ diag, int m, double *__restrict__ spec,
                             double *__restrict__ s_masked, double *__restrict__ s_unmasked,
                             double *__restrict__ summary) {
    __shared__ double sh_sum[ENT_THREADS], sh_min[ENT_THREADS], sh_max[ENT_THREADS];
    double tot = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int t = threadIdx.x; t < m; t += ENT_THREADS) {
        const double nu = diag[t], nd = diag[m + t], nn = diag[2 * m + t];
        const double s[4] = {__dadd_rn(__dsub_rn(__dsub_rn(1.0, nu), nd), nn), __dsub_rn(nu, nn), __dsub_rn(nd, nn), nn};
        int dummy = 0;
        const double e = masked_entropy4(s[0], s[1], s[2], s[3], dummy);
        double un = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (spec) spec[(size_t)r * m + t] = s[r];
            mn = fmin(mn, s[r]);
            mx = fmax(mx, s[r]);
            const double term = __dmul_rn(log(s[r]), s[r]);   // NaN for a zero or negative entry, as numpy
            un = (r == 0) ? term : __dadd_rn(un, term);
        }
        if (s_masked) s_masked[t] = e;
        if (s_unmasked) s_unmasked[t] = -un;
        tot += e;
    }
    sh_sum[threadIdx.x] = tot;
    sh_min[threadIdx.x] = mn;
    sh_max[threadIdx.x] = mx;
    __syncthreads();
    for (int off = ENT_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sh_sum[threadIdx.x] += sh_sum[threadIdx.x + off];
            sh_min[threadIdx.x] = fmin(sh_min[threadIdx.x], sh_min[threadIdx.x + off]);
            sh_max[threadIdx.x] = fmax(sh_max[threadIdx.x], sh_max[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && summary) {
        int flags = 0;
        // shannon() sees the whole (4, m) array at once: qio/entropy.py:15-20
        if (sh_min[0] < 0.0) {
            if (sh_min[0] < -1e-6) flags |= QIO_B200_FLAG_WARN_NEGATIVE;
        } else if (sh_max[0] > 1.0) {
            flags |= QIO_B200_FLAG_RAISE_GT1;
        }
        summary[0] = sh_sum[0];
        summary[1] = sh_min[0];
        summary[2] = sh_max[0];
        summary[3] = (double)flags;
    }
}

// shannon() of a flat array (qio/entropy.py:14-22): m < 8 is summed sequentially like numpy's
// small-array loop, larger arrays by a fixed-order tree.
__global__ void __launch_bounds__(ENT_THREADS)
    shannon_kernel(const double *__restrict__ p, long long m, double *__restrict__ summary) {
    __shared__ double sh_sum[ENT_THREADS], sh_min[ENT_THREADS], sh_max[ENT_THREADS];
    double tot = 0.0, mn = INFINITY, mx = -INFINITY;
    if (m < 8) {
        if (threadIdx.x == 0)
            for (long long t = 0; t < m; ++t) {
                const double v = p[t];
                mn = fmin(mn, v);
                mx = fmax(mx, v);
                if (v > 0.0) tot = __dadd_rn(tot, __dmul_rn(v, log(v)));
            }
    } else {
        for (long long t = threadIdx.x; t < m; t += ENT_THREADS) {
            const double v = p[t];
            mn = fmin(mn, v);
            mx = fmax(mx, v);
            if (v > 0.0) tot += v * log(v);
        }
    }
    sh_sum[threadIdx.x] = tot;
    sh_min[threadIdx.x] = mn;
    sh_max[threadIdx.x] = mx;
    __syncthreads();
    for (int off = ENT_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sh_sum[threadIdx.x] += sh_sum[threadIdx.x + off];
            sh_min[threadIdx.x] = fmin(sh_min[threadIdx.x], sh_min[threadIdx.x + off]);
            sh_max[threadIdx.x] = fmax(sh_max[threadIdx.x], sh_max[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int flags = 0;
        if (sh_min[0] < 0.0) {
            if (sh_min[0] < -1e-6) flags |= QIO_B200_FLAG_WARN_NEGATIVE;
        } else if (sh_max[0] > 1.0) {
            flags |= QIO_B200_FLAG_RAISE_GT1;
        }
        summary[0] = -sh_sum[0];
        summary[1] = sh_min[0];
        summary[2] = sh_max[0];
        summary[3] = (double)flags;
    }
}

}  // namespace qio

extern "C" int qio_b200_shannon(const double *p, long long m, double *summary, void *stream) {
    using namespace qio;
    QIO_REQUIRE(m >= 0 && summary && (p || m == 0), "shannon: bad arguments");
    shannon_kernel<<<1, ENT_THREADS, 0, as_stream(stream)>>>(p, m, summary);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_orbital_diagonals(const double *gamma, const double *Gamma, int n, int a0, int na,
                                          const int32_t *inds, int m, double *diag, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && m >= 0, "orbital_diagonals: bad sizes");
    QIO_REQUIRE(a0 >= 0 && na >= 0 && a0 + na <= n, "orbital_diagonals: bad shard range");
    if (m == 0) return QIO_B200_OK;
    QIO_REQUIRE(gamma && inds && diag && (Gamma || na == 0), "orbital_diagonals: NULL pointer");
    orbital_diagonals_kernel<<<(m + 127) / 128, 128, 0, as_stream(stream)>>>(gamma, Gamma, n, a0, na, inds, m, diag);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_orbital_entropies(const double *diag, int m, double *spec, double *s_masked,
                                          double *s_unmasked, double *summary, void *stream) {
    using namespace qio;
    QIO_REQUIRE(m >= 0, "orbital_entropies: bad size");
    QIO_REQUIRE(diag || m == 0, "orbital_entropies: NULL pointer");
    orbital_entropies_kernel<<<1, ENT_THREADS, 0, as_stream(stream)>>>(diag, m, spec, s_masked, s_unmasked, summary);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}
