// One axis of the 4-index rotation as a PERSISTENT, TMA-fed FP64 tensor-core GEMM (reference qio/grad/jacobi.py:90,
// qio/grad/gradient.py:214: einsum 'ia,jb,kc,ld,abcd->ijkl', four single-axis passes).
//
//     C[M = rows, N = n] = A B,   A[m, k] = in[k * ld_in + m]  (K = kdim rows of M contiguous columns),   B[k, i] = U[i * ldu + k]
//     MODE_CYCLIC:    out[m * n + i]       (the contracted leading axis re-appears as the last one)
//     MODE_LEAD_KEEP: out[i * ld_out + m]  (order kept; the flush of the deferred sweep state)
//
// tcgen05.mma has no FP64 kind, so the tensor path is warp-level mma.sync m8n8k4 f64 (SASS DMMA.8x8x4) -- what this kernel
// adds over rotate.cu is everything around it:
//   * persistent: one CTA per SM loops over 128-row M tiles; the K slabs of consecutive tiles stream through one ring of
//     shared-memory stages without ever draining the pipeline (a tile has only n / 8 = 13..25 slabs);
//   * the A slabs come in by TMA (cp.async.bulk.tensor.3d, one instruction per 8 KB stage, issued by a producer warp and
//     tracked by mbarriers); the tensor is described as [16][K][M / 16] so that one box = 8 k-rows x 128 m-columns lands as
//     sixteen-double lines in SWIZZLE_128B order, and an MMA k-group takes the even (then the odd) rows of a slab: the
//     fragment loads of a half-warp then hit all 32 banks exactly once.  Out-of-range rows / columns are zero-filled by
//     the TMA unit, no predication in the kernel;
//   * B (the rotation matrix, <= 104 columns per CTA) is resident in shared memory for the whole kernel, k-major with a
//     pitch = 2 mod 16 doubles (conflict-free for the same k-groups);
//   * the N tiles are split 7 + 6 between the two warp columns (warps w and w + 4 share a scheduler), so that n = 100 is
//     padded to 104 columns instead of 112; when the upper half of the last K slab is padding (K = 1..4 mod 8) that slab runs
//     one k-group instead of two, as a second fully unrolled copy of the slab body (a run-time bound on the k-group loop
//     cost 4-7 %, more than it saved);
//   * no division on the hot path: the ring position / phase and the (batch, M tile) decomposition are running counters (a
//     64-bit q % stages per slab was 10 % of the stall samples, the B prologue another 5 %: the producer now starts
//     streaming while the consumers load B eight elements in flight per thread);
//   * batched: `batch` independent problems of the same shape (the slabs of a leading-index shard, qio_b200_rotate_axis_batched)
//     are one launch -- the tensor map has the batch as its fourth dimension and the persistent tile loop runs over
//     (batch, M tile).
#include <type_traits>

#include "tma.cuh"

namespace qio {

constexpr int RT_BM = 128;            // rows of an M tile
constexpr int RT_KS = 8;              // k-rows per stage
constexpr int RT_STAGE_DOUBLES = RT_KS * RT_BM;      // 8 KB
constexpr int RT_NCOL = 2;            // warp columns (N direction).  Three columns (12 warps, 5 + 4 + 4 tiles, three warps per
                                      // scheduler against the "wait" stalls of the DMMA chain) measured +1 % at n = 100 and -5 % at
                                      // n = 200: not kept (profiles/r2_final_summary.md)
constexpr int RT_CONSUMERS = 128 * RT_NCOL;     // 8 warps: 4 (M) x 2 (N)
constexpr int RT_THREADS = RT_CONSUMERS + 32;
constexpr int RT_MAX_NT = 7;          // n8-tiles of the widest warp column
constexpr int RT_MAX_TILES = 13;      // n8-tiles per CTA: 7 + 6
constexpr int RT_MAX_COLS = 8 * RT_MAX_TILES;          // 104 columns per CTA
constexpr int RT_MAX_STAGES = 8;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// grid = (CTAs along M (persistent), column blocks of <= 104);  dynamic shared memory:
//   [stages][RT_STAGE_DOUBLES] A ring (1024-byte aligned) | Bs[kpad][np] | full[stages], empty[stages] mbarriers
template <bool LEAD_KEEP>
__global__ void __launch_bounds__(RT_THREADS, 1)
    rotate_tma_kernel(const __grid_constant__ CUtensorMap tmA, const double *__restrict__ U, double *__restrict__ out, int n,
                      int kdim, long long ldu, long long rows, long long ld_out, int cols_per_cta, int np, int kpad, int stages,
                      int batch, long long bs_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);
    double *sB = sA + (size_t)stages * RT_STAGE_DOUBLES;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sB + (size_t)kpad * np);
    unsigned long long *empty = full + RT_MAX_STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.y * cols_per_cta, ncols = min(cols_per_cta, n - n0);
    const int nslabs = kpad / RT_KS;
    const long long mtiles_b = (rows + RT_BM - 1) / RT_BM, mtiles = mtiles_b * batch;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, RT_CONSUMERS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == RT_CONSUMERS / 32) {
        // ---- producer warp: one lane streams the slabs of all tiles of this CTA through the ring (it starts while the
        // consumers are still loading B).  Stage index and phase are running counters: a 64-bit q % stages per slab in every
        // warp was 10 % of the kernel's stall samples (profiles/r2_final_summary.md) ---------------------------------------
        if (lane == 0) {
            int s = 0;
            unsigned ph = 1;            // parity of the "empty" completion to wait for: none in the first round
            bool first_round = true;
            int b = (int)(blockIdx.x / mtiles_b), mt = (int)(blockIdx.x - b * mtiles_b);       // tile = b * mtiles_b + mt
            for (long long tile = blockIdx.x; tile < mtiles; tile += gridDim.x) {
                for (int ks = 0; ks < nslabs; ++ks) {
                    if (!first_round) mbar_wait(empty + s, ph);
                    mbar_expect_tx(full + s, RT_STAGE_DOUBLES * sizeof(double));
                    tma_load_4d(sA + (size_t)s * RT_STAGE_DOUBLES, &tmA, 0, ks * RT_KS, mt * (RT_BM / 16), b, full + s);
                    if (++s == stages) {
                        s = 0;
                        ph ^= 1u;
                        first_round = false;
                    }
                }
                for (mt += (int)gridDim.x; mt >= mtiles_b; mt -= (int)mtiles_b) ++b;
            }
        }
        return;
    }

    // B: sB[k][c] = U[n0 + c][k], zero beyond (kdim, ncols); eight loads in flight per consumer thread
    for (int base = 0; base < kpad * np; base += 8 * RT_CONSUMERS) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = base + u * RT_CONSUMERS + tid, c = e / kpad, k = e - c * kpad;       // k fastest: coalesced reads of a row of U
            v[u] = (e < kpad * np && c < ncols && k < kdim) ? U[(long long)(n0 + c) * ldu + k] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = base + u * RT_CONSUMERS + tid, c = e / kpad, k = e - c * kpad;
            if (e < kpad * np) sB[(size_t)k * np + c] = v[u];
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(RT_CONSUMERS) : "memory");

    // ---- consumers: 8 warps as 4 (M) x 2 (N) -------------------------------------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 3) * 32, wcol = warp >> 2;
    // the n8-tiles are dealt to the warp columns as evenly as possible (13 -> 7 + 6), the wider ones first
    const int tiles_n = (ncols + 7) / 8;
    const int nt = (tiles_n + RT_NCOL - 1 - wcol) / RT_NCOL;
    int wn = 0;
    for (int c = 0; c < wcol; ++c) wn += 8 * ((tiles_n + RT_NCOL - 1 - c) / RT_NCOL);
    // per-lane constants of the swizzled A fragment: m = wm + 8 mi + g -> 16-double line chunk and position inside
    int a_line[4], a_unit[4], a_low[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int mm = wm + 8 * mi + g;
        a_line[mi] = (mm >> 4) * RT_KS;        // first line of that m-chunk
        a_unit[mi] = (mm & 15) >> 1;
        a_low[mi] = mm & 1;
    }
    int s = 0;
    unsigned ph = 0;                    // parity of the "full" completion to wait for
    // kdim = 1..4 mod 8: rows 4..7 of the last slab are padding (zero-filled by the TMA unit, zero in B)
    const bool half_last = (kdim % RT_KS) != 0 && (kdim % RT_KS) <= RT_KS / 2;
    int tb = (int)(blockIdx.x / mtiles_b), tmt = (int)(blockIdx.x - tb * mtiles_b);       // tile = tb * mtiles_b + tmt (no division per tile)
    for (long long tile = blockIdx.x; tile < mtiles; tile += gridDim.x) {
        double acc[4][RT_MAX_NT][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < RT_MAX_NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        // one slab = two k-groups of 4 rows; all fragments of the slab are loaded before its DMMAs (one exposed shared-memory
        // latency per slab).  When the upper half of the last slab is padding that slab runs one k-group only -- as a second,
        // fully unrolled copy of the body (a run-time bound on the k-group loop cost more than it saved).
        auto slab = [&](auto ngrp_c, int ks) {
            constexpr int NG = decltype(ngrp_c)::value;
            mbar_wait(full + s, ph);
            const double *a = sA + (size_t)s * RT_STAGE_DOUBLES;
            const double *b = sB + (size_t)(ks * RT_KS) * np + wn + g;
            double af[NG][4], bf[NG][RT_MAX_NT];
#pragma unroll
            for (int grp = 0; grp < NG; ++grp) {
                // k-groups of a full slab: rows {0, 2, 4, 6}, then {1, 3, 5, 7} (conflict-free fragment loads); the single
                // k-group of a half-empty last slab: rows 0..3 (two-way conflicts, once per tile)
                const int r = NG == 2 ? 2 * t + grp : t;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
                    af[grp][mi] = a[(a_line[mi] + r) * 16 + ((a_unit[mi] ^ r) << 1) + a_low[mi]];
#pragma unroll
                for (int ni = 0; ni < RT_MAX_NT; ++ni) bf[grp][ni] = ni < nt ? b[(size_t)r * np + 8 * ni] : 0.0;
            }
#pragma unroll
            for (int grp = 0; grp < NG; ++grp)
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < RT_MAX_NT; ++ni)
                        if (ni < nt) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[grp][mi], bf[grp][ni]);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            if (++s == stages) {
                s = 0;
                ph ^= 1u;
            }
        };
        const int nfull = half_last ? nslabs - 1 : nslabs;
        // (starting the second warp column a few slabs late, so that the epilogues of the two warps of a scheduler do not
        // coincide, changed nothing: 0.697 -> 0.700 ms per n = 100 pass)
        for (int ks = 0; ks < nfull; ++ks) slab(std::integral_constant<int, 2>{}, ks);
        if (half_last) slab(std::integral_constant<int, 1>{}, nslabs - 1);
        // epilogue: C fragment (row g, cols 2t, 2t+1) of each 8x8 tile
        const long long m0 = (long long)tmt * RT_BM;
        double *outb = out + (long long)tb * bs_out;
        const int colw = n0 + wn + 2 * t, cend = n0 + ncols;
        if (LEAD_KEEP) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const long long m = m0 + wm + 8 * mi + g;
                if (m >= rows) continue;
                double *o = outb + (long long)colw * ld_out + m;
#pragma unroll
                for (int ni = 0; ni < RT_MAX_NT; ++ni) {
                    if (ni >= nt) continue;
                    const int col = colw + 8 * ni;
                    if (col < cend) o[(long long)(8 * ni) * ld_out] = acc[mi][ni][0];
                    if (col + 1 < cend) o[(long long)(8 * ni + 1) * ld_out] = acc[mi][ni][1];
                }
            }
        } else {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const long long m = m0 + wm + 8 * mi + g;
                if (m >= rows) continue;
                double *o = outb + m * ld_out + colw;
#pragma unroll
                for (int ni = 0; ni < RT_MAX_NT; ++ni) {
                    if (ni >= nt) continue;
                    const int col = colw + 8 * ni;
                    if (col + 1 < cend) {
                        *reinterpret_cast<double2 *>(o + 8 * ni) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                    } else if (col < cend) {
                        o[8 * ni] = acc[mi][ni][0];
                    }
                }
            }
        }
        for (tmt += (int)gridDim.x; tmt >= mtiles_b; tmt -= (int)mtiles_b) ++tb;
    }
}

static EncodeTiledFn lookup_encode_tiled() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        return reinterpret_cast<EncodeTiledFn>(p);
    return nullptr;
}
EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = lookup_encode_tiled();      // driver entry point: thread-safe one-time lookup
    return fn;
}

// Returns 1 if the TMA kernel took the job, 0 if the shape is not eligible (the caller falls back to rotate.cu), < 0 on error.
int rotate_axis_tma(const double *U, const double *in, double *out, int n, int kdim, long long ldu, long long rows,
                    long long ld_in, long long ld_out, bool lead_keep, int batch, long long bs_in, long long bs_out, cudaStream_t st) {
    if (rows % 16 != 0 || ld_in % 2 != 0 || ((uintptr_t)in % 16) != 0 || rows / 16 > 0x7fffffffLL || kdim < 1) return 0;
    if (batch < 1 || (batch > 1 && (bs_in % 2 != 0 || bs_out % 2 != 0))) return 0;
    if (!lead_keep && (ld_out % 2 != 0 || ((uintptr_t)out % 16) != 0 || n % 2 != 0)) return 0;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return 0;
    const int kpad = (kdim + RT_KS - 1) / RT_KS * RT_KS;      // rows of B in shared memory (the TMA unit zero-fills A beyond kdim)
    const int ny = (n + RT_MAX_COLS - 1) / RT_MAX_COLS;
    const int cols_per_cta = ((n + ny - 1) / ny + 7) / 8 * 8;      // balanced column blocks, multiples of 8
    int np = cols_per_cta;
    while (np % 16 != 2) ++np;
    int dev = 0, smem_max = 0;
    QIO_CUDA(cudaGetDevice(&dev));
    QIO_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t fixed = sizeof(double) * (size_t)kpad * np + 2 * RT_MAX_STAGES * sizeof(unsigned long long) + 1024;
    if (fixed + 3 * RT_STAGE_DOUBLES * sizeof(double) > (size_t)smem_max) return 0;
    int stages = (int)(((size_t)smem_max - fixed) / (RT_STAGE_DOUBLES * sizeof(double)));
    stages = stages > RT_MAX_STAGES ? RT_MAX_STAGES : stages;
    const size_t smem = (size_t)stages * RT_STAGE_DOUBLES * sizeof(double) + fixed - 1024;

    CUtensorMap tm;
    const cuuint64_t gdim[4] = {16, (cuuint64_t)kdim, (cuuint64_t)(rows / 16), (cuuint64_t)batch};
    const cuuint64_t gstride[3] = {(cuuint64_t)ld_in * sizeof(double), 16 * sizeof(double),
                                   (cuuint64_t)(batch > 1 ? bs_in : ld_in * (long long)kdim) * sizeof(double)};
    const cuuint32_t box[4] = {16, RT_KS, RT_BM / 16, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(in), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return 0;
    const long long mtiles = (rows + RT_BM - 1) / RT_BM * batch;
    const int gx = (int)std::min<long long>(mtiles, std::max(1, sm_count() / ny));
    dim3 grid(gx, ny);
    if (lead_keep) {
        QIO_CUDA(cudaFuncSetAttribute(rotate_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rotate_tma_kernel<true><<<grid, RT_THREADS, smem, st>>>(tm, U, out, n, kdim, ldu, rows, ld_out, cols_per_cta, np, kpad, stages, batch, bs_out);
    } else {
        QIO_CUDA(cudaFuncSetAttribute(rotate_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rotate_tma_kernel<false><<<grid, RT_THREADS, smem, st>>>(tm, U, out, n, kdim, ldu, rows, ld_out, cols_per_cta, np, kpad, stages, batch, bs_out);
    }
    QIO_LAUNCH_CHECK();
    return 1;
}

}  // namespace qio
