// Pair kernels: block gather, all-pairs gradient / Hessian (a-8), candidate costs (a-4) and the
// batched angle line search (a-5).  Reference: qio/grad/gradient.py:13-146, qio/grad/jacobi.py:15-158.
//
// Data flow.  A pair (i, j) only ever reads Gamma[{i,j}^4] (16 doubles, each in its own 32-byte
// sector of the n^4 tensor) and 2 x 4 entries of gamma.  The gather kernel pulls those 24 numbers
// per pair into a dense (P, 24) table -- one thread per entry, so a warp has 32 independent
// sector loads in flight -- and every downstream kernel streams that table with coalesced,
// 16-byte vectorised loads staged through shared memory.  In the sharded layout each rank fills
// the entries whose leading index it owns and the tables are summed (exact: x + 0).
#include "linesearch.cuh"

namespace qio {

// ---- gather -------------------------------------------------------------------------------
__global__ void pair_blocks_kernel(const double *__restrict__ gamma, const double *__restrict__ Gamma, int n,
                                   int a0, int na, const int32_t *__restrict__ pairs, int P,
                                   double *__restrict__ blocks) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)P * QIO_B200_BLOCK_DOUBLES) return;
    const int p = (int)(idx / QIO_B200_BLOCK_DOUBLES), t = (int)(idx % QIO_B200_BLOCK_DOUBLES);
    const int i = pairs[2 * p], j = pairs[2 * p + 1];
    double v = 0.0;
    if (t < 16) {
        const int a = (t & 8) ? j : i;
        if (a >= a0 && a < a0 + na) {
            const size_t n3 = (size_t)n * n * n;
            v = __ldg(Gamma + block_entry_offset(t, i, j, n) - (size_t)a0 * n3);
        }
    } else if (i >= a0 && i < a0 + na) {
        v = __ldg(gamma + block_entry_offset(t, i, j, n));
    }
    blocks[idx] = v;
}

// Stage `count` consecutive records of the table into shared memory with double2 loads.
constexpr int REC_PAD = QIO_B200_BLOCK_DOUBLES + 1;   // 25: conflict-free per-thread record reads

__device__ __forceinline__ void stage_records(const double *__restrict__ blocks, long long first, int count,
                                              double *sm /* [count][REC_PAD] */) {
    const double2 *src = reinterpret_cast<const double2 *>(blocks + first * QIO_B200_BLOCK_DOUBLES);
    const int n2 = count * (QIO_B200_BLOCK_DOUBLES / 2);
    for (int e = threadIdx.x; e < n2; e += blockDim.x) {
        const double2 v = src[e];
        const int rec = (2 * e) / QIO_B200_BLOCK_DOUBLES, off = (2 * e) % QIO_B200_BLOCK_DOUBLES;
        sm[rec * REC_PAD + off] = v.x;
        sm[rec * REC_PAD + off + 1] = v.y;
    }
}

// ---- all-pairs gradient / Hessian ------------------------------------------------------------
constexpr int GH_THREADS = 32;        // one pair per thread; small CTAs so that 4,950 pairs already give more CTAs than SMs

__device__ __forceinline__ void pair_from_linear(long long k, int &i, int &j) {
    // k = i(i-1)/2 + j, 0 <= j < i   (row-major lower triangle, the order of gradient.py:128-130)
    long long ii = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)k)) * 0.5);
    while (ii * (ii - 1) / 2 > k) --ii;
    while ((ii + 1) * ii / 2 <= k) ++ii;
    i = (int)ii;
    j = (int)(k - ii * (ii - 1) / 2);
}

__global__ void __launch_bounds__(GH_THREADS)
    fqi_grad_hess_kernel(const double *__restrict__ blocks, int n, long long P, const int32_t *__restrict__ inactive,
                         double *__restrict__ dX, double *__restrict__ ddX) {
    __shared__ double sm[GH_THREADS * REC_PAD];
    const long long first = (long long)blockIdx.x * GH_THREADS;
    const int count = (int)min((long long)GH_THREADS, P - first);
    stage_records(blocks, first, count, sm);
    __syncthreads();
    if ((int)threadIdx.x >= count) return;
    PairBlock b;
    load_block(sm + threadIdx.x * REC_PAD, b);
    int i, j;
    pair_from_linear(first + threadIdx.x, i, j);
    double g1, g2;
    pair_grad_hess(b, inactive[i] != 0, inactive[j] != 0, g1, g2);
    dX[(size_t)i * n + j] = g1;
    dX[(size_t)j * n + i] = -g1;
    ddX[(size_t)i * n + j] = g2;
    ddX[(size_t)j * n + i] = g2;
}

// Explicit pair list (any order of i, j): g1[p], g2[p].
__global__ void __launch_bounds__(GH_THREADS)
    pair_grad_hess_kernel(const double *__restrict__ blocks, const int32_t *__restrict__ pairs, int P,
                          const int32_t *__restrict__ inactive, double *__restrict__ g1, double *__restrict__ g2) {
    __shared__ double sm[GH_THREADS * REC_PAD];
    const long long first = (long long)blockIdx.x * GH_THREADS;
    const int count = (int)min((long long)GH_THREADS, (long long)P - first);
    stage_records(blocks, first, count, sm);
    __syncthreads();
    if ((int)threadIdx.x >= count) return;
    PairBlock b;
    load_block(sm + threadIdx.x * REC_PAD, b);
    const long long p = first + threadIdx.x;
    double a1, a2;
    pair_grad_hess(b, inactive[pairs[2 * p]] != 0, inactive[pairs[2 * p + 1]] != 0, a1, a2);
    g1[p] = a1;
    g2[p] = a2;
}

// Fused single-GPU variant: 32 pairs per CTA, one thread per (pair, entry) for the gather.
constexpr int GHF_PAIRS = 32;

__global__ void __launch_bounds__(GHF_PAIRS *QIO_B200_BLOCK_DOUBLES)
    fqi_grad_hess_fused_kernel(const double *__restrict__ gamma, const double *__restrict__ Gamma, int n, long long P,
                               const int32_t *__restrict__ inactive, double *__restrict__ dX,
                               double *__restrict__ ddX) {
    __shared__ double sm[GHF_PAIRS * REC_PAD];
    const int t = threadIdx.x % QIO_B200_BLOCK_DOUBLES, pl = threadIdx.x / QIO_B200_BLOCK_DOUBLES;
    const long long k = (long long)blockIdx.x * GHF_PAIRS + pl;
    int i = 1, j = 0;
    if (k < P) {
        pair_from_linear(k, i, j);
        const size_t off = block_entry_offset(t, i, j, n);
        sm[pl * REC_PAD + t] = (t < 16) ? __ldg(Gamma + off) : __ldg(gamma + off);
    }
    __syncthreads();
    if (threadIdx.x >= GHF_PAIRS) return;
    const long long kk = (long long)blockIdx.x * GHF_PAIRS + threadIdx.x;
    if (kk >= P) return;
    pair_from_linear(kk, i, j);
    PairBlock b;
    load_block(sm + threadIdx.x * REC_PAD, b);
    double g1, g2;
    pair_grad_hess(b, inactive[i] != 0, inactive[j] != 0, g1, g2);
    dX[(size_t)i * n + j] = g1;
    dX[(size_t)j * n + i] = -g1;
    ddX[(size_t)i * n + j] = g2;
    ddX[(size_t)j * n + i] = g2;
}

// ---- explicit candidates (jacobi_cost) --------------------------------------------------------
__global__ void pair_cost_kernel(const double *__restrict__ blocks, const int32_t *__restrict__ cand_block,
                                 const int32_t *__restrict__ cand_pair, const double *__restrict__ cand_cs, int C,
                                 const int32_t *__restrict__ inactive, double *__restrict__ cost,
                                 int32_t *__restrict__ flags) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C) return;
    PairBlock b;
    load_block(blocks + (size_t)cand_block[t] * QIO_B200_BLOCK_DOUBLES, b);
    int f = 0;
    const int i = cand_pair[2 * t], j = cand_pair[2 * t + 1];
    cost[t] = pair_cost(b, cand_cs[2 * t], cand_cs[2 * t + 1], inactive[i] != 0, inactive[j] != 0, f);
    if (flags) flags[t] = f;
}

// ---- batched line search: one CTA per pair -------------------------------------------------------
template <bool STRICT>
__global__ void __launch_bounds__(LS_THREADS)
    pair_linesearch_kernel(const double *__restrict__ blocks, const int32_t *__restrict__ pairs, int P,
                           const int32_t *__restrict__ inactive, ThetaTablesDev tab,
                           qio_b200_ls_result *__restrict__ results) {
    __shared__ LineSearchSmem sm;
    __shared__ double rec[QIO_B200_BLOCK_DOUBLES];
    __shared__ PairCoef pc;
    __shared__ double2 coarse_cs_sm[320];
    if (tab.n_coarse <= 320) {      // the coarse cos/sin table is reused for every pair of this CTA
        for (int t = threadIdx.x; t < tab.n_coarse; t += LS_THREADS)
            coarse_cs_sm[t] = reinterpret_cast<const double2 *>(tab.coarse_cs)[t];
        tab.coarse_cs = reinterpret_cast<const double *>(coarse_cs_sm);
    }
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        if (threadIdx.x < QIO_B200_BLOCK_DOUBLES)
            rec[threadIdx.x] = blocks[(size_t)p * QIO_B200_BLOCK_DOUBLES + threadIdx.x];
        __syncthreads();
        const int i = pairs[2 * p], j = pairs[2 * p + 1];
        qio_b200_ls_result res;
        if (STRICT) {
            PairBlock b;
            load_block(rec, b);
            cta_linesearch(b, inactive[i] != 0, inactive[j] != 0, tab, sm, res);
        } else {
            if (threadIdx.x == 0) {
                make_coef(rec, pc);
                sm.flags = 0;
            }
            __syncthreads();
            cta_linesearch_t<LS_THREADS, PairCoef>(pc, inactive[i] != 0, inactive[j] != 0, tab, sm, res);
        }
        if (threadIdx.x == 0) results[p] = res;
    }
}

}  // namespace qio

// =================================================================================================
extern "C" int qio_b200_pair_blocks(const double *gamma, const double *Gamma, int n, int a0, int na,
                                    const int32_t *pairs, int P, double *blocks, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && P >= 0, "pair_blocks: bad sizes");
    QIO_REQUIRE(a0 >= 0 && na >= 0 && a0 + na <= n, "pair_blocks: bad shard range");
    if (P == 0) return QIO_B200_OK;
    QIO_REQUIRE(gamma && pairs && blocks && (Gamma || na == 0), "pair_blocks: NULL pointer");
    const long long total = (long long)P * QIO_B200_BLOCK_DOUBLES;
    pair_blocks_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(gamma, Gamma, n, a0, na, pairs, P, blocks);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_fqi_grad_hess(const double *blocks, int n, const int32_t *inactive, double *dX, double *ddX,
                                      void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && inactive && dX && ddX, "fqi_grad_hess: bad arguments");
    cudaStream_t st = as_stream(stream);
    QIO_CUDA(cudaMemsetAsync(dX, 0, sizeof(double) * n * n, st));
    QIO_CUDA(cudaMemsetAsync(ddX, 0, sizeof(double) * n * n, st));
    const long long P = (long long)n * (n - 1) / 2;
    if (P == 0) return QIO_B200_OK;
    QIO_REQUIRE(blocks, "fqi_grad_hess: blocks is NULL");
    fqi_grad_hess_kernel<<<(unsigned)((P + GH_THREADS - 1) / GH_THREADS), GH_THREADS, 0, st>>>(blocks, n, P, inactive, dX, ddX);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_pair_grad_hess(const double *blocks, const int32_t *pairs, int P, const int32_t *inactive,
                                       double *g1, double *g2, void *stream) {
    using namespace qio;
    QIO_REQUIRE(P >= 0, "pair_grad_hess: bad size");
    if (P == 0) return QIO_B200_OK;
    QIO_REQUIRE(blocks && pairs && inactive && g1 && g2, "pair_grad_hess: NULL pointer");
    pair_grad_hess_kernel<<<(P + GH_THREADS - 1) / GH_THREADS, GH_THREADS, 0, as_stream(stream)>>>(blocks, pairs, P, inactive, g1, g2);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_fqi_grad_hess_fused(const double *gamma, const double *Gamma, int n, const int32_t *inactive,
                                            double *dX, double *ddX, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && gamma && Gamma && inactive && dX && ddX, "fqi_grad_hess_fused: bad arguments");
    cudaStream_t st = as_stream(stream);
    QIO_CUDA(cudaMemsetAsync(dX, 0, sizeof(double) * n * n, st));
    QIO_CUDA(cudaMemsetAsync(ddX, 0, sizeof(double) * n * n, st));
    const long long P = (long long)n * (n - 1) / 2;
    if (P == 0) return QIO_B200_OK;
    fqi_grad_hess_fused_kernel<<<(unsigned)((P + GHF_PAIRS - 1) / GHF_PAIRS), GHF_PAIRS * QIO_B200_BLOCK_DOUBLES, 0, st>>>(
        gamma, Gamma, n, P, inactive, dX, ddX);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_pair_cost(const double *blocks, const int32_t *cand_block, const int32_t *cand_pair,
                                  const double *cand_cs, int C, const int32_t *inactive, double *cost, int32_t *flags,
                                  void *stream) {
    using namespace qio;
    QIO_REQUIRE(C >= 0, "pair_cost: bad size");
    if (C == 0) return QIO_B200_OK;
    QIO_REQUIRE(blocks && cand_block && cand_pair && cand_cs && inactive && cost, "pair_cost: NULL pointer");
    pair_cost_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(blocks, cand_block, cand_pair, cand_cs, C, inactive, cost, flags);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

extern "C" int qio_b200_pair_linesearch(const double *blocks, const int32_t *pairs, int P, const int32_t *inactive,
                                        const qio_b200_theta_tables *h_tables, qio_b200_ls_result *results,
                                        int strict_order, void *stream) {
    using namespace qio;
    QIO_REQUIRE(P >= 0, "pair_linesearch: bad size");
    if (P == 0) return QIO_B200_OK;
    QIO_REQUIRE(blocks && pairs && inactive && h_tables && results, "pair_linesearch: NULL pointer");
    QIO_REQUIRE(h_tables->n_coarse > 0 && h_tables->fine_stride > 0 && h_tables->fine_stride <= LS_MAX_FINE,
                "pair_linesearch: table sizes out of range");
    // persistent grid: a multiple of the SM count (6 CTAs of 320 threads fit per SM)
    const int grid = min(P, sm_count() * 6);
    if (strict_order)
        pair_linesearch_kernel<true><<<grid, LS_THREADS, 0, as_stream(stream)>>>(blocks, pairs, P, inactive, to_dev(*h_tables), results);
    else
        pair_linesearch_kernel<false><<<grid, LS_THREADS, 0, as_stream(stream)>>>(blocks, pairs, P, inactive, to_dev(*h_tables), results);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}
