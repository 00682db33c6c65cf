// a-1  prep_rdm12 (reference qio/qio.py:152-174) on the device.
//
// For a fixed leading index a and second output index b the operation is a "matrix plus twice
// its transpose" on the n x n matrix  Y[p][q] = dm2[a][p][b][q]  (row pitch n^2, rows contiguous):
//     Gamma[a][b][c][d] = (2 * Y[d][c] + Y[c][d]) / 6 .
// One CTA owns a pair of mirrored 32 x 32 tiles (I, J) / (J, I) of Y: both are read once with
// row-contiguous (256 B) accesses, transposed through padded shared memory, and both output
// tiles are written once, again row-contiguous.  HBM traffic = 16 n^4 bytes, the algorithmic
// minimum.  Bit-exact with numpy: 2*x is exact, one rounding in the sum,
// one IEEE division.
#include "common.cuh"

namespace qio {

constexpr int TILE = 32;
constexpr int ROWS_PER_PASS = 8;

__global__ void __launch_bounds__(TILE *ROWS_PER_PASS)
    prep_rdm2_kernel(const double *__restrict__ dm2, double *__restrict__ Gamma, int n, int tiles) {
    __shared__ double tA[TILE][TILE + 1];
    __shared__ double tB[TILE][TILE + 1];

    // decode the tile pair I <= J from the linear index
    int p = blockIdx.x, I = 0;
    while (p >= tiles - I) {
        p -= tiles - I;
        ++I;
    }
    const int J = I + p;
    const int b = blockIdx.y;
    const size_t a = blockIdx.z;
    const size_t n1 = (size_t)n, n2 = n1 * n1, n3 = n2 * n1;
    const double *Y = dm2 + a * n3 + (size_t)b * n1;   // Y[p][q] = Y[p * n2 + q]
    double *out = Gamma + a * n3 + (size_t)b * n2;     // out[c][d] = out[c * n + d]

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i0 = I * TILE, j0 = J * TILE;
    const bool diag = (I == J);

#pragma unroll
    for (int r = ty; r < TILE; r += ROWS_PER_PASS) {
        const int rowA = i0 + r, colA = j0 + tx;
        tA[r][tx] = (rowA < n && colA < n) ? Y[(size_t)rowA * n2 + colA] : 0.0;
        if (!diag) {
            const int rowB = j0 + r, colB = i0 + tx;
            tB[r][tx] = (rowB < n && colB < n) ? Y[(size_t)rowB * n2 + colB] : 0.0;
        }
    }
    __syncthreads();

#pragma unroll
    for (int r = ty; r < TILE; r += ROWS_PER_PASS) {
        {   // output tile (I, J): c = i0 + r, d = j0 + tx ; Y[d][c] lives in tile (J, I)
            const int c = i0 + r, d = j0 + tx;
            if (c < n && d < n) {
                const double yt = diag ? tA[tx][r] : tB[tx][r];
                out[(size_t)c * n1 + d] = __ddiv_rn(__dadd_rn(2.0 * yt, tA[r][tx]), 6.0);
            }
        }
        if (!diag) {   // output tile (J, I): c = j0 + r, d = i0 + tx
            const int c = j0 + r, d = i0 + tx;
            if (c < n && d < n) out[(size_t)c * n1 + d] = __ddiv_rn(__dadd_rn(2.0 * tA[tx][r], tB[r][tx]), 6.0);
        }
    }
}

__global__ void prep_rdm1_kernel(const double *__restrict__ dm1, double *__restrict__ gamma, int n) {
    const int m = 2 * n;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * m) return;
    const int r = (int)(idx / m), c = (int)(idx % m);
    gamma[idx] = ((r & 1) == (c & 1)) ? __ddiv_rn(dm1[(size_t)(r >> 1) * n + (c >> 1)], 2.0) : 0.0;
}

}  // namespace qio

extern "C" int qio_b200_prep_rdm12(const double *dm1, const double *dm2, double *gamma, double *Gamma,
                                   int n, int na, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && n <= 65535, "prep_rdm12: n out of range");
    QIO_REQUIRE(na >= 0 && na <= n, "prep_rdm12: na out of range");
    QIO_REQUIRE((dm1 == nullptr) == (gamma == nullptr), "prep_rdm12: dm1 and gamma must both be given or both be NULL");
    cudaStream_t st = as_stream(stream);
    if (dm1) {
        const long long total = 4LL * n * n;
        prep_rdm1_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dm1, gamma, n);
        QIO_LAUNCH_CHECK();
    }
    if (na > 0) {
        QIO_REQUIRE(dm2 && Gamma, "prep_rdm12: dm2 / Gamma is NULL");
        QIO_REQUIRE(dm2 != Gamma, "prep_rdm12: in-place operation is not supported");
        const int tiles = (n + TILE - 1) / TILE;
        dim3 grid(tiles * (tiles + 1) / 2, n, na);
        prep_rdm2_kernel<<<grid, dim3(TILE, ROWS_PER_PASS), 0, st>>>(dm2, Gamma, n, tiles);
        QIO_LAUNCH_CHECK();
    }
    return QIO_B200_OK;
}
