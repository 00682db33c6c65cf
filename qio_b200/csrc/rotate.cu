// Orbital rotation of the RDMs (reference qio/grad/jacobi.py:200-202, qio/grad/gradient.py:212-214).
//
// The 4-index rotation  Gamma'[ijkl] = sum U[ia] U[jb] U[kc] U[ld] Gamma[abcd]  is done as four
// single-axis passes; each pass contracts the LEADING axis and appends the new index at the end,
//     out[(b,c,d), i] = sum_a in[a, (b,c,d)] * U[i, a],                         (MODE_CYCLIC)
// so the axes rotate cyclically and after four passes the tensor is back in its original order.
// Every pass is a tall GEMM  C[M = n^3, N = n] = A B  with B[k][i] = U[i][k]:  2 n^5 flop per pass over
// 16 n^4 bytes, i.e. n/8 flop/byte -- above the B200 FP64 balance point for n >= ~50, hence the one
// path of this library that runs on the FP64 tensor pipe (mma.sync m8n8k4 f64 -> DMMA; tcgen05 has no
// FP64 kind).  Two more operand arrangements serve the flush of the deferred-axis sweep state:
//     out[i, m] = sum_a U[i, a] in[a, m]      leading axis, order kept        (MODE_LEAD_KEEP)
//     out[m, i] = sum_d in[m, d] U[i, d]      last axis                       (MODE_LAST)
//
// Tiling: CTA = 128 (M) x 16*NT8 (N), 8 warps as 4 (M) x 2 (N), warp tile 32 x 8*NT8, K in slabs
// of 16 through a 4-stage cp.async pipeline.  Shared-memory pitches (132 / 20 doubles) make all
// fragment loads bank-conflict free.

#include "common.cuh"

namespace qio {

// rotate_tma.cu: persistent TMA-fed kernel for MODE_CYCLIC / MODE_LEAD_KEEP; returns 1 if it ran, 0 if not eligible
int rotate_axis_tma(const double *U, const double *in, double *out, int n, int kdim, long long ldu, long long rows,
                    long long ld_in, long long ld_out, bool lead_keep, int batch, long long bs_in, long long bs_out, cudaStream_t st);

constexpr int RA_BM = 128;
constexpr int RA_BK = 16;
constexpr int RA_STAGES = 4;
constexpr int RA_THREADS = 256;
constexpr int RA_APITCH = RA_BM + 4;   // doubles, A stored [k][m]
constexpr int RA_BPITCH = RA_BK + 4;   // doubles, B stored [n][k]; also A stored [m][k] in MODE_LAST

constexpr int MODE_CYCLIC = 0, MODE_LEAD_KEEP = 1, MODE_LAST = 2;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = pred ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// in:  MODE_CYCLIC / MODE_LEAD_KEEP: [n][ld_in] (K rows of M contiguous columns)
//      MODE_LAST: [rows][ld_in] (M rows of K = n contiguous columns, ld_in >= n)
// out: MODE_CYCLIC / MODE_LAST: [rows][ld_out] (ld_out >= n);  MODE_LEAD_KEEP: [n][ld_out] (ld_out >= rows)
template <int NT8, bool ALIGNED, int MODE>
__global__ void __launch_bounds__(RA_THREADS, 1)
    rotate_axis_kernel(const double *__restrict__ U, const double *__restrict__ in, double *__restrict__ out, int n,
                       int kdim, long long ldu, long long rows, long long ld_in, long long ld_out) {
    constexpr int BN = 16 * NT8;
    constexpr int A_STAGE = (MODE == MODE_LAST) ? RA_BM * RA_BPITCH : RA_BK * RA_APITCH;
    extern __shared__ __align__(16) double smem[];
    double *sA = smem;                              // [STAGES][A_STAGE]
    double *sB = smem + RA_STAGES * A_STAGE;        // [STAGES][BN][BPITCH]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * (8 * NT8);
    const long long m0 = (long long)blockIdx.x * RA_BM;
    const int n0 = blockIdx.y * BN;
    const int ktiles = (kdim + RA_BK - 1) / RA_BK;      // n = number of outputs (rows of U), kdim = contraction length

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * RA_BK;
        double *a = sA + stage * A_STAGE;
        double *b = sB + stage * BN * RA_BPITCH;
        if (MODE == MODE_LAST) {
            // A tile: 128 rows (m) x 16 contiguous k
            if (ALIGNED) {
                for (int e = tid; e < RA_BM * (RA_BK / 2); e += RA_THREADS) {
                    const int mr = e / (RA_BK / 2), kc = (e % (RA_BK / 2)) * 2;
                    const bool ok = (m0 + mr < rows) && (k0 + kc < kdim);
                    const double *src = ok ? in + (m0 + mr) * ld_in + k0 + kc : in;
                    cp_async16(a + mr * RA_BPITCH + kc, src, ok);
                }
            } else {
                for (int e = tid; e < RA_BM * RA_BK; e += RA_THREADS) {
                    const int mr = e / RA_BK, kc = e % RA_BK;
                    const bool ok = (m0 + mr < rows) && (k0 + kc < kdim);
                    const double *src = ok ? in + (m0 + mr) * ld_in + k0 + kc : in;
                    cp_async8(a + mr * RA_BPITCH + kc, src, ok);
                }
            }
        } else if (ALIGNED) {
            // A tile: 16 rows (k) x 128 contiguous m
            for (int e = tid; e < RA_BK * (RA_BM / 2); e += RA_THREADS) {
                const int kr = e / (RA_BM / 2), mc = (e % (RA_BM / 2)) * 2;
                const bool ok = (k0 + kr < kdim) && (m0 + mc < rows);
                const double *src = ok ? in + (long long)(k0 + kr) * ld_in + m0 + mc : in;
                cp_async16(a + kr * RA_APITCH + mc, src, ok);
            }
        } else {
            for (int e = tid; e < RA_BK * RA_BM; e += RA_THREADS) {
                const int kr = e / RA_BM, mc = e % RA_BM;
                const bool ok = (k0 + kr < kdim) && (m0 + mc < rows);
                const double *src = ok ? in + (long long)(k0 + kr) * ld_in + m0 + mc : in;
                cp_async8(a + kr * RA_APITCH + mc, src, ok);
            }
        }
        // B tile: BN rows of U x 16 contiguous k
        if (ALIGNED) {
            for (int e = tid; e < BN * (RA_BK / 2); e += RA_THREADS) {
                const int nr = e / (RA_BK / 2), kc = (e % (RA_BK / 2)) * 2;
                const bool ok = (n0 + nr < n) && (k0 + kc < kdim);
                const double *src = ok ? U + (long long)(n0 + nr) * ldu + k0 + kc : U;
                cp_async16(b + nr * RA_BPITCH + kc, src, ok);
            }
        } else {
            for (int e = tid; e < BN * RA_BK; e += RA_THREADS) {
                const int nr = e / RA_BK, kc = e % RA_BK;
                const bool ok = (n0 + nr < n) && (k0 + kc < kdim);
                const double *src = ok ? U + (long long)(n0 + nr) * ldu + k0 + kc : U;
                cp_async8(b + nr * RA_BPITCH + kc, src, ok);
            }
        }
    };

    double acc[4][NT8][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
    for (int s = 0; s < RA_STAGES - 1; ++s) {
        if (s < ktiles) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<RA_STAGES - 2>();
        __syncthreads();
        {   // prefetch the slab that reuses the stage consumed in the previous iteration
            const int nk = kt + RA_STAGES - 1;
            if (nk < ktiles) load_stage(nk % RA_STAGES, nk);
            cp_async_commit();
        }
        const double *a = sA + (kt % RA_STAGES) * A_STAGE;
        const double *b = sB + (kt % RA_STAGES) * BN * RA_BPITCH;
#pragma unroll
        for (int kk = 0; kk < RA_BK; kk += 4) {
            double af[4], bf[NT8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
                af[mi] = (MODE == MODE_LAST) ? a[(wm + 8 * mi + g) * RA_BPITCH + kk + t]
                                             : a[(kk + t) * RA_APITCH + wm + 8 * mi + g];
#pragma unroll
            for (int ni = 0; ni < NT8; ++ni) bf[ni] = b[(wn + 8 * ni + g) * RA_BPITCH + kk + t];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < NT8; ++ni) dmma8x8x4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C fragment (row g, cols 2t, 2t+1) of each 8x8 tile
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const long long m = m0 + wm + 8 * mi + g;
        if (m >= rows) continue;
#pragma unroll
        for (int ni = 0; ni < NT8; ++ni) {
            const int col = n0 + wn + 8 * ni + 2 * t;
            if (MODE == MODE_LEAD_KEEP) {
                if (col < n) out[(long long)col * ld_out + m] = acc[mi][ni][0];
                if (col + 1 < n) out[(long long)(col + 1) * ld_out + m] = acc[mi][ni][1];
            } else {
                double *orow = out + m * ld_out;
                if (ALIGNED && col + 1 < n) {
                    *reinterpret_cast<double2 *>(orow + col) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                } else {
                    if (col < n) orow[col] = acc[mi][ni][0];
                    if (col + 1 < n) orow[col + 1] = acc[mi][ni][1];
                }
            }
        }
    }
}

template <int NT8, int MODE>
static int launch_rotate_axis(const double *U, const double *in, double *out, int n, int kdim, long long ldu,
                              long long rows, long long ld_in, long long ld_out, cudaStream_t st) {
    constexpr int BN = 16 * NT8;
    constexpr int A_STAGE = (MODE == MODE_LAST) ? RA_BM * RA_BPITCH : RA_BK * RA_APITCH;
    const size_t smem = sizeof(double) * (RA_STAGES * A_STAGE + RA_STAGES * BN * RA_BPITCH);
    const bool aligned = (n % 2 == 0) && (kdim % 2 == 0) && (ldu % 2 == 0) && (ld_in % 2 == 0) && (ld_out % 2 == 0) && (rows % 2 == 0) &&
                         ((uintptr_t)in % 16 == 0) && ((uintptr_t)U % 16 == 0) && ((uintptr_t)out % 16 == 0);
    dim3 grid((unsigned)((rows + RA_BM - 1) / RA_BM), (unsigned)((n + BN - 1) / BN));
    if (aligned) {
        QIO_CUDA(cudaFuncSetAttribute(rotate_axis_kernel<NT8, true, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rotate_axis_kernel<NT8, true, MODE><<<grid, RA_THREADS, smem, st>>>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out);
    } else {
        QIO_CUDA(cudaFuncSetAttribute(rotate_axis_kernel<NT8, false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rotate_axis_kernel<NT8, false, MODE><<<grid, RA_THREADS, smem, st>>>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out);
    }
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

template <int MODE>
static int dispatch_rotate_axis(const double *U, const double *in, double *out, int n, int kdim, long long ldu,
                                long long rows, long long ld_in, long long ld_out, cudaStream_t st) {
    if (MODE != MODE_LAST) {
        const int rc = rotate_axis_tma(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, MODE == MODE_LEAD_KEEP, 1, 0, 0, st);
        if (rc != 0) return rc < 0 ? rc : QIO_B200_OK;
    }
    // N tile = 16*NT8 columns: fewest padded columns, ties to the wider tile
    int best = 1;
    long long best_pad = -1;
    for (int nt = 1; nt <= 7; ++nt) {
        const long long tiles = (n + 16 * nt - 1) / (16 * nt);
        const long long padded = tiles * 16 * nt;
        if (best_pad < 0 || padded <= best_pad) {
            best_pad = padded;
            best = nt;
        }
    }
    switch (best) {
        case 1: return launch_rotate_axis<1, MODE>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, st);
        case 2: return launch_rotate_axis<2, MODE>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, st);
        case 3: return launch_rotate_axis<3, MODE>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, st);
        case 4: return launch_rotate_axis<4, MODE>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, st);
        case 5: return launch_rotate_axis<5, MODE>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, st);
        case 6: return launch_rotate_axis<6, MODE>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, st);
        default: return launch_rotate_axis<7, MODE>(U, in, out, n, kdim, ldu, rows, ld_in, ld_out, st);
    }
}

// gamma_out = (U x 1_2) gamma_in (U x 1_2)^T ; one CTA per output row, no workspace.
__global__ void __launch_bounds__(256)
    rotate_rdm1_kernel(const double *__restrict__ U, const double *__restrict__ gin, double *__restrict__ gout, int n) {
    extern __shared__ double trow[];   // [2n]  T[r, :] = sum_a U[i, a] gamma[2a + sigma, :]
    const int r = blockIdx.x, i = r >> 1, sigma = r & 1, ld = 2 * n;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) {
        double acc = 0.0;
        for (int a = 0; a < n; ++a) acc += U[(size_t)i * n + a] * gin[(size_t)(2 * a + sigma) * ld + c];
        trow[c] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ld; c += blockDim.x) {
        const int j = c >> 1, tau = c & 1;
        double acc = 0.0;
        for (int b = 0; b < n; ++b) acc += trow[2 * b + tau] * U[(size_t)j * n + b];
        gout[(size_t)r * ld + c] = acc;
    }
}

}  // namespace qio

extern "C" int qio_b200_rotate_axis(const double *U, const double *in, double *out, int n, long long rows,
                                    long long ld_in, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && rows >= 0 && ld_in >= rows, "rotate_axis: bad sizes");
    if (rows == 0) return QIO_B200_OK;
    QIO_REQUIRE(U && in && out && in != out, "rotate_axis: bad pointers");
    return dispatch_rotate_axis<MODE_CYCLIC>(U, in, out, n, n, n, rows, ld_in, n, as_stream(stream));
}

extern "C" int qio_b200_rotate_axis_batched(const double *U, const double *in, double *out, int n, long long rows,
                                            long long ld_in, int batch, long long batch_stride_in, long long batch_stride_out,
                                            void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && rows >= 0 && ld_in >= rows && batch >= 0, "rotate_axis_batched: bad sizes");
    if (rows == 0 || batch == 0) return QIO_B200_OK;
    QIO_REQUIRE(U && in && out && in != out, "rotate_axis_batched: bad pointers");
    QIO_REQUIRE(batch == 1 || (batch_stride_in >= ld_in * (long long)(n - 1) + rows && batch_stride_out >= rows * (long long)n),
                "rotate_axis_batched: batch strides overlap");
    const int rc = rotate_axis_tma(U, in, out, n, n, n, rows, ld_in, n, false, batch, batch_stride_in, batch_stride_out, as_stream(stream));
    if (rc != 0) return rc < 0 ? rc : QIO_B200_OK;
    for (int b = 0; b < batch; ++b) {       // shapes the TMA kernel does not take: one launch per problem
        const int r = dispatch_rotate_axis<MODE_CYCLIC>(U, in + (long long)b * batch_stride_in, out + (long long)b * batch_stride_out,
                                                        n, n, n, rows, ld_in, n, as_stream(stream));
        if (r != QIO_B200_OK) return r;
    }
    return QIO_B200_OK;
}

extern "C" int qio_b200_contract_axis(const double *U, const double *in, double *out, int n, long long rows,
                                      int last_axis, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && rows >= 0, "contract_axis: bad sizes");
    if (rows == 0) return QIO_B200_OK;
    QIO_REQUIRE(U && in && out && in != out, "contract_axis: bad pointers");
    if (last_axis) return dispatch_rotate_axis<MODE_LAST>(U, in, out, n, n, n, rows, n, n, as_stream(stream));
    return dispatch_rotate_axis<MODE_LEAD_KEEP>(U, in, out, n, n, n, rows, rows, rows, as_stream(stream));
}

extern "C" int qio_b200_contract_axis_partial(const double *U, long long ldu, const double *in, double *out, int n,
                                              int kdim, long long rows, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && kdim >= 0 && rows >= 0 && ldu >= kdim, "contract_axis_partial: bad sizes");
    if (rows == 0) return QIO_B200_OK;
    QIO_REQUIRE(U && out && (in || kdim == 0) && in != out, "contract_axis_partial: bad pointers");
    if (kdim == 0) {
        QIO_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)n * rows, as_stream(stream)));
        return QIO_B200_OK;
    }
    return dispatch_rotate_axis<MODE_LEAD_KEEP>(U, in, out, n, kdim, ldu, rows, rows, rows, as_stream(stream));
}

extern "C" int qio_b200_contract_axis_cyclic(const double *U, long long ldu, const double *in, double *out, int n_out,
                                             int kdim, long long rows, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n_out > 0 && kdim > 0 && rows >= 0 && ldu >= kdim, "contract_axis_cyclic: bad sizes");
    if (rows == 0) return QIO_B200_OK;
    QIO_REQUIRE(U && in && out && in != out, "contract_axis_cyclic: bad pointers");
    return dispatch_rotate_axis<MODE_CYCLIC>(U, in, out, n_out, kdim, ldu, rows, rows, n_out, as_stream(stream));
}

extern "C" int qio_b200_rotate_rdm2(const double *U, double *Gamma, double *work, int n, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && U && Gamma && work && Gamma != work, "rotate_rdm2: bad arguments");
    const long long n3 = (long long)n * n * n;
    double *src = Gamma, *dst = work;
    for (int pass = 0; pass < 4; ++pass) {
        const int rc = qio_b200_rotate_axis(U, src, dst, n, n3, n3, stream);
        if (rc != QIO_B200_OK) return rc;
        double *tmp = src;
        src = dst;
        dst = tmp;
    }
    return QIO_B200_OK;   // four swaps: the result is back in Gamma
}

extern "C" int qio_b200_rotate_rdm1(const double *U, const double *gamma_in, double *gamma_out, int n, void *stream) {
    using namespace qio;
    QIO_REQUIRE(n > 0 && U && gamma_in && gamma_out && gamma_in != gamma_out, "rotate_rdm1: bad arguments");
    rotate_rdm1_kernel<<<2 * n, 256, sizeof(double) * 2 * n, as_stream(stream)>>>(U, gamma_in, gamma_out, n);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}
