// In-place two-orbital (Givens) rotation of the 2-RDM on all four axes and of the 1-RDM
// (reference qio/grad/jacobi.py:63-115, which runs a dense O(n^5) einsum for the same result).
//
// Only entries with at least one index in {i, j} change: n^4 - (n-2)^4 ~ 8 n^3 of them.  They
// decompose into disjoint orbits (replace every special index by i or j), so ONE launch can
// process all of them without ordering constraints:
//
//  * "bundles": the d-rows  Gamma[a, b, c, :]  with at least one of a, b, c in {i, j}.  A bundle is
//    the 2 / 4 / 8 rows obtained by flipping the special indices; one warp loads them with
//    coalesced (16-byte when n is even) accesses, rotates the d = i, j columns, applies the
//    butterflies over the special leading axes in registers and writes the rows back.
//    3(n-2)^2 + 3(n-2) + 1 bundles, ~6 n^3 doubles, all row-contiguous.
//  * "thin" work: for a, b, c all outside {i, j} only the two entries d = i and d = j rotate;
//    one thread per (a, b, c), two isolated 8-byte accesses each (2 (n-2)^3 doubles; each entry
//    is alone in its 32-byte sector -- this is the part the deferred-axis sweep layout removes).
//  * gamma: rows / columns i, j of the up-up and down-down blocks, one extra CTA.
#pragma once
#include "common.cuh"

namespace qio {

struct GivensParams {
    int i, j;
    double c, s;
    int active;   // 0 = do nothing (rotation rejected)
};

// map generic rank k in [0, n-2) to an orbital index skipping lo < hi
__device__ __forceinline__ int skip2(int k, int lo, int hi) {
    k += (k >= lo);
    k += (k >= hi);
    return k;
}

template <bool SA, bool SB, bool SC, int W>
__device__ __forceinline__ void rotate_bundle(double *__restrict__ G, int n, int i, int j, double c, double s,
                                              int a, int b, int cc, int lane) {
    constexpr int R = (SA ? 2 : 1) * (SB ? 2 : 1) * (SC ? 2 : 1);
    const size_t n1 = n;
    size_t off[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int bit = r;
        int rc = cc, rb = b, ra = a;
        if (SC) { rc = (bit & 1) ? j : i; bit >>= 1; }
        if (SB) { rb = (bit & 1) ? j : i; bit >>= 1; }
        if (SA) { ra = (bit & 1) ? j : i; }
        off[r] = (((size_t)ra * n1 + rb) * n1 + rc) * n1;
    }
    double xi[R], xj[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        xi[r] = G[off[r] + i];
        xj[r] = G[off[r] + j];
    }
    __syncwarp();   // every lane holds the old d = i, j entries before any lane overwrites them
    for (int d0 = lane * W; d0 < n; d0 += 32 * W) {
        double v[R][W];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (W == 2) {
                const double2 t = *reinterpret_cast<const double2 *>(G + off[r] + d0);
                v[r][0] = t.x;
                v[r][W - 1] = t.y;
            } else {
                v[r][0] = G[off[r] + d0];
            }
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int d = d0 + w;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (d == i) v[r][w] = c * xi[r] + s * xj[r];
                if (d == j) v[r][w] = c * xj[r] - s * xi[r];
            }
            // butterflies over the special leading axes: (x, y) -> (c x + s y, -s x + c y)
#pragma unroll
            for (int stride = 1; stride < R; stride <<= 1)
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!(r & stride)) {
                        const double x = v[r][w], y = v[r | stride][w];
                        v[r][w] = c * x + s * y;
                        v[r | stride][w] = c * y - s * x;
                    }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (W == 2)
                *reinterpret_cast<double2 *>(G + off[r] + d0) = make_double2(v[r][0], v[r][W - 1]);
            else
                G[off[r] + d0] = v[r][0];
        }
    }
}

template <int W>
__device__ __forceinline__ void bundle_dispatch(double *__restrict__ G, int n, int i, int j, double c, double s,
                                                long long w, int lane) {
    const long long m = n - 2, m2 = m * m;
    const int lo = min(i, j), hi = max(i, j);
    if (w < 3 * m2) {
        const int type = (int)(w / m2);
        const long long r = w % m2;
        const int k1 = skip2((int)(r / m), lo, hi), k2 = skip2((int)(r % m), lo, hi);
        if (type == 0) rotate_bundle<true, false, false, W>(G, n, i, j, c, s, 0, k1, k2, lane);
        else if (type == 1) rotate_bundle<false, true, false, W>(G, n, i, j, c, s, k1, 0, k2, lane);
        else rotate_bundle<false, false, true, W>(G, n, i, j, c, s, k1, k2, 0, lane);
        return;
    }
    w -= 3 * m2;
    if (w < 3 * m) {
        const int type = (int)(w / m);
        const int k = skip2((int)(w % m), lo, hi);
        if (type == 0) rotate_bundle<true, true, false, W>(G, n, i, j, c, s, 0, 0, k, lane);
        else if (type == 1) rotate_bundle<true, false, true, W>(G, n, i, j, c, s, 0, k, 0, lane);
        else rotate_bundle<false, true, true, W>(G, n, i, j, c, s, k, 0, 0, lane);
        return;
    }
    rotate_bundle<true, true, true, W>(G, n, i, j, c, s, 0, 0, 0, lane);
}

__device__ __forceinline__ long long givens_num_bundles(int n) {
    const long long m = n - 2;
    return 3 * m * m + 3 * m + 1;
}

// thin work item t in [0, (n-2)^3)
__device__ __forceinline__ void rotate_thin(double *__restrict__ G, int n, int i, int j, double c, double s, long long t) {
    const long long m = n - 2;
    const int lo = min(i, j), hi = max(i, j);
    const int kc = skip2((int)(t % m), lo, hi);
    const long long t2 = t / m;
    const int kb = skip2((int)(t2 % m), lo, hi), ka = skip2((int)(t2 / m), lo, hi);
    double *row = G + (((size_t)ka * n + kb) * n + kc) * (size_t)n;
    const double x = row[i], y = row[j];
    row[i] = c * x + s * y;
    row[j] = c * y - s * x;
}

// gamma <- V gamma V^T on the up-up and down-down blocks; one CTA.  `zero_offspin` clears the
// up-down / down-up entries (what the reference's fresh gamma_ holds, jacobi.py:87).
__device__ __forceinline__ void rotate_gamma_cta(double *__restrict__ gamma, int n, int i, int j, double c, double s,
                                                 bool zero_offspin) {
    const size_t ld = 2 * (size_t)n;
    // columns: T = gamma V^T
    for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
        const int a = t >> 1, spin = t & 1;
        double *row = gamma + (2 * (size_t)a + spin) * ld;
        const double x = row[2 * i + spin], y = row[2 * j + spin];
        row[2 * i + spin] = c * x + s * y;
        row[2 * j + spin] = c * y - s * x;
    }
    __threadfence_block();
    __syncthreads();
    // rows: gamma' = V T
    for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
        const int b = t >> 1, spin = t & 1;
        double *ri = gamma + (2 * (size_t)i + spin) * ld + 2 * b + spin;
        double *rj = gamma + (2 * (size_t)j + spin) * ld + 2 * b + spin;
        const double x = *ri, y = *rj;
        *ri = c * x + s * y;
        *rj = c * y - s * x;
    }
    if (zero_offspin) {
        const long long total = (long long)ld * ld;
        for (long long e = threadIdx.x; e < total; e += blockDim.x) {
            const int r = (int)(e / ld), cidx = (int)(e % ld);
            if ((r & 1) != (cidx & 1)) gamma[e] = 0.0;
        }
    }
}

constexpr int GIVENS_THREADS = 256;

struct GivensGrid {
    unsigned bundle_blocks, thin_blocks;
    unsigned total() const { return bundle_blocks + thin_blocks + 1; }
};

static inline GivensGrid givens_grid(int n) {
    const long long m = n - 2;
    const long long bundles = 3 * m * m + 3 * m + 1;
    const long long thin = m * m * m;
    GivensGrid g;
    g.bundle_blocks = (unsigned)((bundles + GIVENS_THREADS / 32 - 1) / (GIVENS_THREADS / 32));
    g.thin_blocks = (unsigned)((thin + GIVENS_THREADS - 1) / GIVENS_THREADS);
    return g;
}

// body shared by the host-parameter and device-parameter kernels
__device__ __forceinline__ void givens_body(double *__restrict__ gamma, double *__restrict__ G, int n,
                                            const GivensParams &p, unsigned bundle_blocks, unsigned thin_blocks,
                                            bool zero_offspin) {
    if (!p.active) return;
    const unsigned bid = blockIdx.x;
    if (bid < bundle_blocks) {
        const long long w = (long long)bid * (GIVENS_THREADS / 32) + (threadIdx.x >> 5);
        if (w >= givens_num_bundles(n)) return;
        if ((n & 1) == 0)
            bundle_dispatch<2>(G, n, p.i, p.j, p.c, p.s, w, threadIdx.x & 31);
        else
            bundle_dispatch<1>(G, n, p.i, p.j, p.c, p.s, w, threadIdx.x & 31);
    } else if (bid < bundle_blocks + thin_blocks) {
        const long long m = n - 2;
        const long long t = (long long)(bid - bundle_blocks) * GIVENS_THREADS + threadIdx.x;
        if (t < m * m * m) rotate_thin(G, n, p.i, p.j, p.c, p.s, t);
    } else if (gamma) {
        rotate_gamma_cta(gamma, n, p.i, p.j, p.c, p.s, zero_offspin);
    }
}

}  // namespace qio
