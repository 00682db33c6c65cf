// a-6  jacobi_transform as an in-place Givens update (see givens.cuh for the decomposition).
#include "givens.cuh"

namespace qio {

__global__ void __launch_bounds__(GIVENS_THREADS)
    givens_apply_kernel(double *__restrict__ gamma, double *__restrict__ G, int n, GivensParams p,
                        unsigned bundle_blocks, unsigned thin_blocks, int zero_offspin) {
    givens_body(gamma, G, n, p, bundle_blocks, thin_blocks, zero_offspin != 0);
}

}  // namespace qio

extern "C" int qio_b200_givens_apply(double *gamma, double *Gamma, int n, int i, int j, double c, double s,
                                     int zero_offspin, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n >= 2, "givens_apply: n must be >= 2");
    QIO_REQUIRE(i >= 0 && i < n && j >= 0 && j < n && i != j, "givens_apply: bad orbital pair");
    QIO_REQUIRE(Gamma, "givens_apply: Gamma is NULL");
    GivensParams p;
    p.i = i;
    p.j = j;
    p.c = c;
    p.s = s;
    p.active = 1;
    const GivensGrid g = givens_grid(n);
    givens_apply_kernel<<<g.total(), GIVENS_THREADS, 0, as_stream(stream)>>>(gamma, Gamma, n, p, g.bundle_blocks,
                                                                            g.thin_blocks, zero_offspin);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}
