// a-1  prep_rdm12 (reference qio/qio.py:152-174) on the device.
//
// For a fixed leading index a and second output index b the operation is a "matrix plus twice
// its transpose" on the n x n matrix  Y[p][q] = dm2[a][p][b][q]  (row pitch n^2, rows contiguous):
//     Gamma[a][b][c][d] = (2 * Y[d][c] + Y[c][d]) / 6 .
// Two kernels.  prep_tma_kernel (even n): persistent, tiles move by TMA in both directions, the combination is done IN
// PLACE in shared memory (see below).  prep_rdm2_kernel (any n; odd n, unaligned pointers): the cp-style kernel described
// next.
// One CTA owns a pair of mirrored T x T tiles (I, J) / (J, I) of Y (T = 52) and keeps them in shared memory (pitch T + 1:
// the transposed reads of a half-warp hit 16 different bank pairs): both are read once with row-contiguous 16-byte loads
// and both output tiles are written once, again row-contiguous; five CTAs per SM overlap each other's load and store phases.  HBM traffic = 16 n^4 bytes, the algorithmic minimum.  Bit-exact with numpy:
// 2*x is exact, one rounding in the sum, one IEEE division.
#include <algorithm>

#include "tma.cuh"

namespace qio {

// ---- x / 6, correctly rounded, in three FP64 operations ----------------------------------------------------------------
// Markstein's division step: with y = RN(1/6), q0 = RN(x y) is within one ulp of x/6, r = x - 6 q0 is exact (one FMA), and
// RN(q0 + r y) is the correctly rounded quotient -- the same value as the IEEE division numpy performs, for a quarter of its
// instructions (the IEEE sequence is ~20 FP64 instructions and was a third of the old kernel's issue slots).  Inputs outside
// [2^-900, 2^900] (where the residual could underflow or the product overflow), infinities and NaNs take the IEEE division.
// Checked against exact rational arithmetic on the host and against __ddiv_rn on 4e9 device samples (tests/test_gpu_large.py).
__device__ __forceinline__ double div6_rn(double x) {
    const double ax = fabs(x);
    if (ax < 0x1p900 && (ax > 0x1p-900 || ax == 0.0)) {
        const double y = 0x1.5555555555555p-3;      // RN(1/6)
        const double q0 = x * y;
        const double r = fma(-6.0, q0, x);
        return r == 0.0 ? q0 : fma(r, y, q0);
    }
    return __ddiv_rn(x, 6.0);
}

// ---- TMA kernel ----------------------------------------------------------------------------------------------------------
// Work item = (a, b, tile pair I <= J) with 64 x 64 tiles of Y[p][q] = dm2[a][p][b][q]:  tile A = Y[I, J], tile B = Y[J, I].
//     out[I, J][r][x] = (2 B[x][r] + A[r][x]) / 6,      out[J, I][x][r] = (2 A[r][x] + B[x][r]) / 6
// use the SAME two numbers: each thread loads 2 x 2 blocks of A and of B (one 16-byte shared-memory access per row), forms
// both results and writes them back over the inputs -- the tiles then leave by TMA from where they arrived.  No staging
// buffer, no index arithmetic for global memory at all (the tensor maps do it), no load -> barrier -> store phases: a
// loader thread fills each of the three stages (2 x 32 KB) the moment a storer thread has released it, so loads run ahead of
// the combination and the stores behind it.
// Tiles are stored as column chunks of 16 doubles (128-byte rows, SWIZZLE_128B), so the transposed reads of B are spread
// over the banks by the swizzle.  Rows / columns beyond n are zero-filled on the way in and clipped on the way out.
constexpr int PT_T = 64;                       // tile edge
constexpr int PT_CH = 16;                      // doubles per chunk row (128 bytes)
constexpr int PT_TILE_DOUBLES = PT_T * PT_T;   // 32 KB
constexpr int PT_STAGES = 3;
constexpr int PT_PREFETCH = 8;
#ifndef PT_MAX_N
#define PT_MAX_N (2 * PT_T)                    // largest n the TMA kernel is used for (see qio_b200_prep_rdm12)
#endif                 // items whose tiles are requested into L2 ahead of the shared-memory stages
constexpr int PT_CONSUMERS = 256;
constexpr int PT_THREADS = PT_CONSUMERS + 64;      // + a loader warp and a storer warp (one thread each)

// element (row, col) of a tile stored as [col chunk][row][16 doubles], 16-byte units XOR-swizzled with the row (128B mode)
__device__ __forceinline__ int pt_off(int row, int col) {
    return (col >> 4) * (PT_T * PT_CH) + row * PT_CH + ((((col & 15) >> 1) ^ (row & 7)) << 1) + (col & 1);
}

// tensor maps by the number of 16-row groups of the box (1..4): a tile of k chunks is moved with boxes of 16 k rows, so that
// neither a load nor -- fatally -- a store touches rows of the neighbouring tile
struct PrepMaps {
    CUtensorMap in[PT_T / PT_CH], out[PT_T / PT_CH];
};

__global__ void __launch_bounds__(PT_THREADS, 1)
    prep_tma_kernel(const __grid_constant__ PrepMaps maps, int n, int na, int tiles, long long items) {
    extern __shared__ __align__(1024) unsigned char prep_dyn[];
    double *stage0 = reinterpret_cast<double *>(prep_dyn);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(stage0 + (size_t)PT_STAGES * 2 * PT_TILE_DOUBLES);
    unsigned long long *done = full + PT_STAGES, *freed = done + PT_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int npairs = tiles * (tiles + 1) / 2;
    // tile t covers the 16-column chunks [t base + min(t, rem), ...) -- base or base + 1 of them (at most 4): balanced, so that no
    // tile pair degenerates into a sliver (n = 200: 64 + 48 + 48 + 40 instead of 64 + 64 + 64 + 8)
    const int nchunks = (n + PT_CH - 1) / PT_CH, tbase = nchunks / tiles, trem = nchunks - tbase * tiles;
    auto tile_start = [&](int t) { return (t * tbase + min(t, trem)) * PT_CH; };
    auto tile_chunks = [&](int t) { return tbase + (t < trem ? 1 : 0); };

    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(done + s, PT_CONSUMERS / 32);
            mbar_init(freed + s, 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // work items in the order (a, b, tile pair); every role walks its own copy of the iterator with 32-bit steps (no 64-bit
    // divisions in the loops)
    struct Item {
        int p, b, a, I, J;      // pair index, second index, leading index, tiles
    };
    auto tile_of = [&](Item &t) {
        int p = t.p, I = 0;
        while (p >= tiles - I) {
            p -= tiles - I;
            ++I;
        }
        t.I = I;
        t.J = I + p;
    };
    auto first_item = [&](long long item) {
        Item t;
        const long long ab = item / npairs;
        t.p = (int)(item - ab * npairs);
        t.a = (int)(ab / n);
        t.b = (int)(ab - (long long)t.a * n);
        tile_of(t);
        return t;
    };
    auto advance = [&](Item &t, unsigned steps) {        // steps < 2^20
        const unsigned q = (unsigned)t.p + steps, db = q / (unsigned)npairs;
        t.p = (int)(q - db * (unsigned)npairs);
        const unsigned bb = (unsigned)t.b + db, da = bb / (unsigned)n;
        t.b = (int)(bb - da * (unsigned)n);
        t.a += (int)da;
        tile_of(t);
    };

    if (warp == PT_CONSUMERS / 32) {
        // ---- loader thread: fills every stage as soon as the storer has released it (loads run ahead of the combination) ----
        if (lane == 0) {
            // The row pieces of a tile are 128 bytes at a pitch of n^2 doubles: a DRAM-unfriendly pattern with a long effective
            // latency.  Shared memory (3 stages) cannot hold enough of it in flight, so the tiles of the next PT_PREFETCH items
            // are requested into L2 first (TMA prefetch: no shared memory, no tracking); the stage loads then hit L2.
            auto prefetch_item = [&](const Item &t) {
                if (t.a >= na) return;
                const int i0 = tile_start(t.I), j0 = tile_start(t.J), ci = tile_chunks(t.I), cj = tile_chunks(t.J);
                for (int ch = 0; ch < cj; ++ch) tma_prefetch_4d(&maps.in[ci - 1], j0 + ch * PT_CH, t.b, i0, t.a);
                if (t.I != t.J)
                    for (int ch = 0; ch < ci; ++ch) tma_prefetch_4d(&maps.in[cj - 1], i0 + ch * PT_CH, t.b, j0, t.a);
            };
            Item cur = first_item(blockIdx.x), ahead = cur;
            for (int d = 0; d < PT_PREFETCH; ++d) {
                prefetch_item(ahead);
                advance(ahead, gridDim.x);
            }
            long long it = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
                const int s = (int)(it % PT_STAGES);
                prefetch_item(ahead);
                advance(ahead, gridDim.x);
                if (it >= PT_STAGES) mbar_wait(freed + s, (unsigned)(((it / PT_STAGES) - 1) & 1));
                const int a = cur.a, b = cur.b, I = cur.I, J = cur.J;
                double *tA = stage0 + (size_t)s * 2 * PT_TILE_DOUBLES, *tB = tA + PT_TILE_DOUBLES;
                const int i0 = tile_start(I), j0 = tile_start(J), ci = tile_chunks(I), cj = tile_chunks(J);
                // A = Y[i0.., j0..]: cj boxes {16 q, 1 b, 16 ci p, 1 a};  B = Y[j0.., i0..]: ci boxes of 16 cj rows
                const int ra = ci * PT_CH, rb_ = cj * PT_CH;
#ifdef PT_EXP_NOLOAD
                mbar_arrive(full + s);
                if (false)
#else
                mbar_expect_tx(full + s, (unsigned)((I != J ? 2 : 1) * ci * cj * PT_CH * PT_CH * sizeof(double)));
#endif
                for (int ch = 0; ch < cj; ++ch)
                    tma_load_4d(tA + (size_t)ch * ra * PT_CH, &maps.in[ci - 1], j0 + ch * PT_CH, b, i0, a, full + s);
#ifndef PT_EXP_NOLOAD
                if (I != J)
                    for (int ch = 0; ch < ci; ++ch)
                        tma_load_4d(tB + (size_t)ch * rb_ * PT_CH, &maps.in[cj - 1], i0 + ch * PT_CH, b, j0, a, full + s);
#endif
                advance(cur, gridDim.x);
            }
        }
        return;
    }
    if (warp == PT_CONSUMERS / 32 + 1) {
        // ---- storer thread: the combined tiles leave from where they arrived; the stage is released once the TMA unit has
        // read it ------------------------------------------------------------------------------------------------------
        if (lane == 0) {
            long long it = 0;
            Item cur = first_item(blockIdx.x);
            for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
                const int s = (int)(it % PT_STAGES);
                mbar_wait(done + s, (unsigned)((it / PT_STAGES) & 1));
                const int a = cur.a, b = cur.b, I = cur.I, J = cur.J;
                advance(cur, gridDim.x);
                const double *tA = stage0 + (size_t)s * 2 * PT_TILE_DOUBLES, *tB = tA + PT_TILE_DOUBLES;
                const int i0 = tile_start(I), j0 = tile_start(J), ci = tile_chunks(I), cj = tile_chunks(J);
                const int ab = a * n + b;
                const int ra = ci * PT_CH, rb_ = cj * PT_CH;
#ifndef PT_EXP_NOSTORE      // (profiling switch: -DPT_EXP_NOSTORE / -DPT_EXP_NOLOAD build the kernel without its stores / loads)
                for (int ch = 0; ch < cj; ++ch)          // out[I, J]: rows i0.., columns j0..
                    tma_store_3d(&maps.out[ci - 1], j0 + ch * PT_CH, i0, ab, tA + (size_t)ch * ra * PT_CH);
                if (I != J)
                    for (int ch = 0; ch < ci; ++ch)      // out[J, I]: rows j0.., columns i0..
                        tma_store_3d(&maps.out[cj - 1], i0 + ch * PT_CH, j0, ab, tB + (size_t)ch * rb_ * PT_CH);
#endif
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(freed + s);
            }
            tma_store_wait<0>();            // shared memory must outlive the last stores
        }
        return;
    }

    // ---- consumers: combine in place.  Warp w takes the block rows rb = w, w + 8, ..., lane l the block column xb = l (a
    // tile has at most 32 x 32 blocks of 2 x 2), so every shared-memory offset is a per-lane constant plus a row term.
    long long it = 0;
    Item cur = first_item(blockIdx.x);
    // A[r][x], x = 2 l: chunk l >> 3, 16-byte unit (l & 7) ^ (r & 7);   B[x][r], x = 2 l: row term x * 16, unit (rb & 7) ^ (x & 7)
    const int a_unit = lane & 7;
    const int b_row0 = (2 * lane) * PT_CH, b_row1 = (2 * lane + 1) * PT_CH, b_x0 = (2 * lane) & 7, b_x1 = (2 * lane + 1) & 7;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int s = (int)(it % PT_STAGES);
        const int I = cur.I, J = cur.J;
        advance(cur, gridDim.x);
        double *tA = stage0 + (size_t)s * 2 * PT_TILE_DOUBLES, *tB = (I == J) ? tA : tA + PT_TILE_DOUBLES;
        // 2 x 2 blocks per direction; chunk c of tile A starts at c * (rows of A) * 16 doubles, of tile B at c * (rows of B) * 16
        const int ra = tile_chunks(I) * PT_CH, rbx = tile_chunks(J) * PT_CH;
        const int hi = min(ra, n - tile_start(I)) >> 1, hj = min(rbx, n - tile_start(J)) >> 1;
        const int a_lane = (lane >> 3) * (ra * PT_CH);
        mbar_wait(full + s, (unsigned)((it / PT_STAGES) & 1));
        if (lane < hj)
            for (int rb = warp; rb < hi; rb += PT_CONSUMERS / 32) {
                if (I == J && rb > lane) continue;          // the mirrored block does both
                const int r = 2 * rb;
                double2 *pa0 = reinterpret_cast<double2 *>(tA + a_lane + r * PT_CH + ((a_unit ^ (r & 7)) << 1));
                double2 *pa1 = reinterpret_cast<double2 *>(tA + a_lane + (r + 1) * PT_CH + ((a_unit ^ ((r + 1) & 7)) << 1));
                const int b_chunk = (rb >> 3) * (rbx * PT_CH), b_unit = rb & 7;
                double2 *pb0 = reinterpret_cast<double2 *>(tB + b_chunk + b_row0 + ((b_unit ^ b_x0) << 1));
                double2 *pb1 = reinterpret_cast<double2 *>(tB + b_chunk + b_row1 + ((b_unit ^ b_x1) << 1));
                const double2 a0 = *pa0, a1 = *pa1, b0 = *pb0, b1 = *pb1;       // a_dr = A[r + dr][x, x + 1], b_dx = B[x + dx][r, r + 1]
                // out[I, J][r + dr][x + dx] = (2 B[x + dx][r + dr] + A[r + dr][x + dx]) / 6, and the mirror image
                double v[8] = {fma(2.0, b0.x, a0.x), fma(2.0, b1.x, a0.y), fma(2.0, b0.y, a1.x), fma(2.0, b1.y, a1.y),
                               fma(2.0, a0.x, b0.x), fma(2.0, a1.x, b0.y), fma(2.0, a0.y, b1.x), fma(2.0, a1.y, b1.y)};
                // one range test for the eight sums (exponent field within [2^-900, 2^900), or zero): the three-operation
                // division for all of them, or -- values no RDM holds -- the IEEE division for all of them
                bool fast = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned hw = (unsigned)__double2hiint(v[k]) & 0x7fffffffu;
                    fast &= ((hw >> 20) - 123u < 1800u) || ((hw | (unsigned)__double2loint(v[k])) == 0u);
                }
                if (fast) {
                    const double y = 0x1.5555555555555p-3;      // RN(1/6)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double q0 = v[k] * y, rr = fma(-6.0, q0, v[k]);
                        v[k] = rr == 0.0 ? q0 : fma(rr, y, q0);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = __ddiv_rn(v[k], 6.0);
                }
                *pb0 = make_double2(v[4], v[5]);
                *pb1 = make_double2(v[6], v[7]);
                *pa0 = make_double2(v[0], v[1]);      // on the diagonal block of a diagonal tile pa == pb: the A form is written last (both agree there)
                *pa1 = make_double2(v[2], v[3]);
            }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(done + s);
    }
}

// ---- the cp-style kernel (any n) ---------------------------------------------------------------------------------------
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

// Returns 1 if the TMA kernel took the job, 0 if it cannot (the caller uses prep_rdm2_kernel), < 0 on error.
static int prep_tma_launch(const double *dm2, double *Gamma, int n, int na, cudaStream_t st) {
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return 0;
    const cuuint64_t n1 = (cuuint64_t)n;
    PrepMaps maps;
    for (int k = 1; k <= PT_T / PT_CH; ++k) {
        {   // dm2[a][p][b][q]: dimensions (q, b, p, a), box = 16 q x 1 b x 16 k p x 1 a
            const cuuint64_t gdim[4] = {n1, n1, n1, (cuuint64_t)na};
            const cuuint64_t gstride[3] = {n1 * 8, n1 * n1 * 8, n1 * n1 * n1 * 8};
            const cuuint32_t box[4] = {PT_CH, 1, (cuuint32_t)(PT_CH * k), 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            if (encode(&maps.in[k - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(dm2), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return 0;
        }
        {   // Gamma[a][b][c][d]: dimensions (d, c, ab), box = 16 d x 16 k c x 1
            const cuuint64_t gdim[3] = {n1, n1, (cuuint64_t)na * n1};
            const cuuint64_t gstride[2] = {n1 * 8, n1 * n1 * 8};
            const cuuint32_t box[3] = {PT_CH, (cuuint32_t)(PT_CH * k), 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            if (encode(&maps.out[k - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, Gamma, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return 0;
        }
    }
    const int tiles = ((n + PT_CH - 1) / PT_CH + PT_T / PT_CH - 1) / (PT_T / PT_CH);     // at most four 16-column chunks per tile
    const long long items = (long long)na * n * (tiles * (tiles + 1) / 2);
    const size_t smem = sizeof(double) * (size_t)PT_STAGES * 2 * PT_TILE_DOUBLES + 3 * PT_STAGES * sizeof(unsigned long long);
    QIO_CUDA(cudaFuncSetAttribute(prep_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<long long>(items, sm_count());
    prep_tma_kernel<<<grid, PT_THREADS, smem, st>>>(maps, n, na, tiles, items);
    QIO_LAUNCH_CHECK();
    return 1;
}

// self-test of div6_rn against the IEEE division: `count` pseudo-random doubles (sign, 2046 exponents, 52 mantissa bits all
// random), mismatches counted
__global__ void div6_selftest_kernel(unsigned long long count, unsigned long long seed, unsigned long long *bad) {
    unsigned long long local = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;      // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const double x = __longlong_as_double((long long)z);
        const double q = div6_rn(x), ref = __ddiv_rn(x, 6.0);
        if (__double_as_longlong(q) != __double_as_longlong(ref) && !(q != q && ref != ref)) ++local;
        // and the same mantissa in the range RDM entries live in
        const double x2 = __longlong_as_double((long long)((z & 0x800fffffffffffffull) | ((unsigned long long)(1023 - 40 + (z >> 52) % 48) << 52)));
        const double q2 = div6_rn(x2), ref2 = __ddiv_rn(x2, 6.0);
        if (__double_as_longlong(q2) != __double_as_longlong(ref2)) ++local;
    }
    if (local) atomicAdd(bad, local);
}

}  // namespace qio

extern "C" int qio_b200_selftest_div6(unsigned long long count, unsigned long long seed, unsigned long long *bad, void *stream) {
    using namespace qio;
    QIO_REQUIRE(bad != nullptr, "selftest_div6: NULL pointer");
    QIO_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), as_stream(stream)));
    div6_selftest_kernel<<<4 * sm_count(), 256, 0, as_stream(stream)>>>(count, seed, bad);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}

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
        // The TMA kernel wins where a matrix is at most 2 x 2 tiles (n <= 128: 0.40 ms at n = 100 against 0.42); beyond, its
        // stores -- 128-byte row pieces per chunk box, the same rows revisited by up to 13 boxes -- drain at 2.2 TB/s and the
        // cp-style kernel with its 416-byte row pieces is faster (n = 200: 6.6 against 7.3 ms; profiles/r2_prep_summary.md).
        if (n % 2 == 0 && n >= 16 && n <= PT_MAX_N && (uintptr_t)dm2 % 16 == 0 && (uintptr_t)Gamma % 16 == 0 &&
            (long long)na * n < 0x7fffffffLL) {
            const int rc = prep_tma_launch(dm2, Gamma, n, na, st);
            if (rc != 0) return rc < 0 ? rc : QIO_B200_OK;
        }
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
