// a-7  one Jacobi cycle (reference qio/grad/jacobi.py:211-233) entirely on the device.
//
// The pair sequence is a strict dependency chain (consecutive pairs share an orbital), so the
// steps run in order on one stream; what is removed is every host round trip: per pair
//   search kernel (1 CTA)  : gathers the 24-number block from the CURRENT RDMs, runs the angle
//                            line search, records the result, rotates rows i, j of U when accepted
//                            and publishes (i, j, c, s, accepted) in device memory;
//   Givens kernel (full grid): reads that record and applies the rotation in place (or exits).
// The host enqueues all 2P launches back to back and reads `results` once at the end.
#include "givens.cuh"
#include "linesearch.cuh"

namespace qio {

struct SweepScratch {
    GivensParams params;
    int accepted_count;
    int flags_or;
    int pad[2];
};

__global__ void sweep_init_kernel(SweepScratch *sc) {
    sc->params.active = 0;
    sc->accepted_count = 0;
    sc->flags_or = 0;
}

__global__ void __launch_bounds__(LS_THREADS)
    sweep_search_kernel(const double *__restrict__ gamma, const double *__restrict__ Gamma, double *__restrict__ U,
                        int n, const int32_t *__restrict__ pairs, int step, const int32_t *__restrict__ inactive,
                        ThetaTablesDev tab, qio_b200_ls_result *__restrict__ results, SweepScratch *__restrict__ sc) {
    __shared__ LineSearchSmem sm;
    __shared__ double rec[QIO_B200_BLOCK_DOUBLES];
    const int i = pairs[2 * step], j = pairs[2 * step + 1];
    if (threadIdx.x < QIO_B200_BLOCK_DOUBLES) {
        const size_t off = block_entry_offset(threadIdx.x, i, j, n);
        rec[threadIdx.x] = (threadIdx.x < 16) ? Gamma[off] : gamma[off];
    }
    __shared__ PairCoef pc;
    __syncthreads();
    if (threadIdx.x == 0) {
        make_coef(rec, pc);
        sm.flags = 0;
    }
    __syncthreads();
    qio_b200_ls_result res;
    cta_linesearch_t<LS_THREADS, PairCoef>(pc, inactive[i] != 0, inactive[j] != 0, tab, sm, res);
    if (threadIdx.x == 0) {
        results[step] = res;
        sc->params.i = i;
        sc->params.j = j;
        sc->params.c = res.cos_theta;
        sc->params.s = res.sin_theta;
        sc->params.active = res.accepted;
        sc->accepted_count += res.accepted;
        sc->flags_or |= res.flags;
    }
    if (res.accepted) {   // U <- V(t, i, j) U : rows i and j   (jacobi.py:230)
        const double c = res.cos_theta, s = res.sin_theta;
        double *ui = U + (size_t)i * n, *uj = U + (size_t)j * n;
        for (int k = threadIdx.x; k < n; k += LS_THREADS) {
            const double x = ui[k], y = uj[k];
            ui[k] = c * x + s * y;
            uj[k] = c * y - s * x;
        }
    }
}

__global__ void __launch_bounds__(GIVENS_THREADS)
    sweep_givens_kernel(double *__restrict__ gamma, double *__restrict__ G, int n, const SweepScratch *__restrict__ sc,
                        unsigned bundle_blocks, unsigned thin_blocks) {
    const GivensParams p = sc->params;
    givens_body(gamma, G, n, p, bundle_blocks, thin_blocks, false);
}

// get_cost_fqi over the inactive mask (entropy.py:24-47) + bookkeeping; one CTA.
__global__ void __launch_bounds__(256)
    sweep_finish_kernel(double *__restrict__ gamma, const double *__restrict__ Gamma, int n,
                        const int32_t *__restrict__ inactive, const SweepScratch *__restrict__ sc,
                        double *__restrict__ summary) {
    __shared__ double sh[256];
    const size_t ld = 2 * (size_t)n, n1 = n;
    double tot = 0.0;
    int flags = 0;
    for (int k = threadIdx.x; k < n; k += 256) {
        if (!inactive[k]) continue;
        const double nu = gamma[(2 * (size_t)k) * ld + 2 * k], nd = gamma[(2 * (size_t)k + 1) * ld + 2 * k + 1];
        const double nn = Gamma[((k * n1 + k) * n1 + k) * n1 + k];
        tot += masked_entropy4(__dadd_rn(__dsub_rn(__dsub_rn(1.0, nu), nd), nn), __dsub_rn(nu, nn), __dsub_rn(nd, nn), nn, flags);
    }
    sh[threadIdx.x] = tot;
    const int any_flags = __syncthreads_or(flags);
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    const int accepted = sc->accepted_count;
    if (threadIdx.x == 0) {
        summary[0] = (double)accepted;
        summary[1] = sh[0];
        summary[2] = (double)(sc->flags_or | any_flags);
        summary[3] = 0.0;
    }
    if (accepted > 0) {   // the reference's jacobi_transform returns zeros outside the same-spin blocks
        const long long total = (long long)ld * ld;
        for (long long e = threadIdx.x; e < total; e += 256) {
            const int r = (int)(e / ld), c = (int)(e % ld);
            if ((r & 1) != (c & 1)) gamma[e] = 0.0;
        }
    }
}

}  // namespace qio

extern "C" size_t qio_b200_jacobi_workspace_bytes(int n, int P) {
    (void)n;
    (void)P;
    return 256;
}

extern "C" int qio_b200_jacobi_cycle(double *gamma, double *Gamma, double *U, int n, const int32_t *pairs, int P,
                                     const int32_t *inactive, const qio_b200_theta_tables *h_tables,
                                     qio_b200_ls_result *results, double *summary, void *workspace, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n >= 2 && P >= 0, "jacobi_cycle: bad sizes");
    QIO_REQUIRE(gamma && Gamma && U && inactive && h_tables && summary && workspace, "jacobi_cycle: NULL pointer");
    QIO_REQUIRE(P == 0 || (pairs && results), "jacobi_cycle: NULL pair list / results");
    QIO_REQUIRE(h_tables->n_coarse > 0 && h_tables->fine_stride > 0 && h_tables->fine_stride <= LS_MAX_FINE,
                "jacobi_cycle: table sizes out of range");
    cudaStream_t st = as_stream(stream);
    SweepScratch *sc = reinterpret_cast<SweepScratch *>(workspace);
    const ThetaTablesDev tab = to_dev(*h_tables);
    const GivensGrid g = givens_grid(n);
    sweep_init_kernel<<<1, 1, 0, st>>>(sc);
    QIO_LAUNCH_CHECK();
    for (int step = 0; step < P; ++step) {
        sweep_search_kernel<<<1, LS_THREADS, 0, st>>>(gamma, Gamma, U, n, pairs, step, inactive, tab, results, sc);
        sweep_givens_kernel<<<g.total(), GIVENS_THREADS, 0, st>>>(gamma, Gamma, n, sc, g.bundle_blocks, g.thin_blocks);
    }
    QIO_LAUNCH_CHECK();
    sweep_finish_kernel<<<1, 256, 0, st>>>(gamma, Gamma, n, inactive, sc, summary);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}
