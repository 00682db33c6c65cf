// a-1  prep_rdm12 (reference qio/qio.py:152-174) on the device.
//
// For a fixed leading index a and second output index b the operation is a "matrix plus twice
// its transpose" on the n x n matrix  Y[p][q] = dm2[a][p][b][q]  (row pitch n^2, rows contiguous):
//     Gamma[a][b][c][d] = (2 * Y[d][c] + Y[c][d]) / 6 .
// One CTA owns a pair of mirrored T x T tiles (I, J) / (J, I) of Y (T = 52) and keeps them in shared memory (pitch T + 1:
// the transposed reads of a half-warp hit 16 different bank pairs): both are read once with row-contiguous 16-byte loads
// and both output tiles are written once, again row-contiguous; five CTAs per SM overlap each other's load and store phases.  HBM traffic = 16 n^4 bytes, the algorithmic minimum.  Bit-exact with numpy:
// 2*x is exact, one rounding in the sum, one IEEE division.
#include "common.cuh"

namespace qio {

constexpr int PREP_THREADS = 256;
constexpr int PREP_T = 52;         // tile edge: 416-byte row pieces (13 sectors, sector-aligned), two tiles = 44 KB -> 5 CTAs per SM

// e / d for e * d < 2^32 with a precomputed magic = ceil(2^32 / d) (d >= 2; magic 0 stands for d = 1)
__device__ __forceinline__ unsigned div_magic(int d) { return d <= 1 ? 0u : (unsigned)((0x100000000ull + d - 1) / d); }
__device__ __forceinline__ int fast_div(int e, unsigned magic) { return magic ? (int)__umulhi((unsigned)e, magic) : e; }

template <bool VEC>
__global__ void __launch_bounds__(PREP_THREADS)
    prep_rdm2_kernel(const double *__restrict__ dm2, double *__restrict__ Gamma, int n, int T, int tiles) {
    extern __shared__ double prep_sm[];
    const int pitch = T + 1;
    double *tA = prep_sm, *tB = prep_sm + (size_t)T * pitch;

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

    const int tid = threadIdx.x;
    const int i0 = I * T, j0 = J * T;
    const int hi = min(T, n - i0), hj = min(T, n - j0);      // valid extent of the two index ranges
    const bool diag = (I == J);

    // tile A = Y[i0.., j0..] (hi x hj), tile B = Y[j0.., i0..] (hj x hi)
    auto load_tile = [&](double *t, int r0, int c0, int nr, int nc) {
        if (VEC) {      // n, T even: 16-byte loads, scalar stores into the odd-pitch tile
            const int ncv = nc >> 1;
            const unsigned magic = div_magic(ncv);
#pragma unroll 4
            for (int e = tid; e < nr * ncv; e += PREP_THREADS) {
                const int r = fast_div(e, magic), cv = e - r * ncv;
                const double2 v = __ldg(reinterpret_cast<const double2 *>(Y + (size_t)(r0 + r) * n2 + c0) + cv);
                t[r * pitch + 2 * cv] = v.x;
                t[r * pitch + 2 * cv + 1] = v.y;
            }
        } else {
            const unsigned magic = div_magic(nc);
#pragma unroll 4
            for (int e = tid; e < nr * nc; e += PREP_THREADS) {
                const int r = fast_div(e, magic), c = e - r * nc;
                t[r * pitch + c] = __ldg(Y + (size_t)(r0 + r) * n2 + c0 + c);
            }
        }
    };
    load_tile(tA, i0, j0, hi, hj);
    if (!diag) load_tile(tB, j0, i0, hj, hi);
    __syncthreads();

    // output tile (I, J): c = i0 + r, d = j0 + x;  Y[d][c] lives in tile (J, I) (= tile A itself on the diagonal)
    const double *tT = diag ? tA : tB;
    {
        const unsigned magic = div_magic(hj);
#pragma unroll 4
        for (int e = tid; e < hi * hj; e += PREP_THREADS) {
            const int r = fast_div(e, magic), x = e - r * hj;
            out[(size_t)(i0 + r) * n1 + j0 + x] = __ddiv_rn(__dadd_rn(2.0 * tT[x * pitch + r], tA[r * pitch + x]), 6.0);
        }
    }
    if (!diag) {    // output tile (J, I): c = j0 + r, d = i0 + x
        const unsigned magic = div_magic(hi);
#pragma unroll 4
        for (int e = tid; e < hj * hi; e += PREP_THREADS) {
            const int r = fast_div(e, magic), x = e - r * hi;
            out[(size_t)(j0 + r) * n1 + i0 + x] = __ddiv_rn(__dadd_rn(2.0 * tA[x * pitch + r], tB[r * pitch + x]), 6.0);
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
        const int T = n <= PREP_T ? n : PREP_T;
        const int tiles = (n + T - 1) / T;
        const size_t smem = sizeof(double) * (size_t)(tiles > 1 ? 2 : 1) * T * (T + 1);
        const bool vec = (n % 2 == 0) && (T % 2 == 0) && ((uintptr_t)dm2 % 16 == 0);
        dim3 grid(tiles * (tiles + 1) / 2, n, na);
        if (vec) {
            QIO_CUDA(cudaFuncSetAttribute(prep_rdm2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            prep_rdm2_kernel<true><<<grid, PREP_THREADS, smem, st>>>(dm2, Gamma, n, T, tiles);
        } else {
            QIO_CUDA(cudaFuncSetAttribute(prep_rdm2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            prep_rdm2_kernel<false><<<grid, PREP_THREADS, smem, st>>>(dm2, Gamma, n, T, tiles);
        }
        QIO_LAUNCH_CHECK();
    }
    return QIO_B200_OK;
}
