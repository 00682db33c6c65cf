// a-7 (pipelined path)  The Jacobi cycle as ONE persistent kernel WITHOUT grid barriers.
//
// Same deferred-axis state as sweep2.cu:  Gamma_cur[a,b,c,d] = sum_{a',d'} W[a,a'] W[d,d'] S[a',b,c,d'],  axes 1, 2 of S in
// the current basis, axes 0 (sharded) and 3 (contiguous) deferred through W.  What changes is who waits for whom.
//
//  * Ownership.  A Givens rotation of axes 1, 2 never mixes the leading index a' or the last index d'.  Every worker CTA
//    therefore OWNS a fixed range of (slab, d') columns of S (a few hundred bytes of every d-row) and is the only CTA that
//    ever reads or writes it: consecutive rotations need no grid-wide ordering at all, a worker only needs the angle.
//  * Lookahead.  The block of pair k after rotations 0..k-1 is a linear image of the "super block" P_k over the <= 6
//    orbitals of pairs k-2, k-1, k taken from the state after rotation k-3:
//        P_k[p,q,r,s] = sum_{a'} W[D_p,a'] ( S[a',D_q,D_r,:] . W[D_s,:] ),     B_k = T^(x4) P_k,
//    T = the two target orbitals after rotations k-2 and k-1 in terms of the orbitals before them (2 x |D|, built from the
//    two angles the search CTA has just found).  So the search CTA needs from the rest of the machine only data that is
//    three rotations old: it runs search after search back to back, and the workers, the reduction and (with peers) the
//    NVLink exchange have a full search period of slack.
//  * Roles.  CTA 0 searches (320 threads) and prefetches the next super block (192 threads).  CTAs 1..4 reduce the workers'
//    partial super blocks in a fixed order and publish the total (to every rank's mailbox when sharded).  CTA 5 keeps the
//    small replicated state (gamma, U, W) and publishes the gamma part of the super block.  CTAs 6.. are workers: wait for
//    angle w (warp 15 polls it one step ahead and prepares the step's plan) -> rotate the rows of the next super block
//    (P_{w+3}) inside their columns, take the dot products with their columns of W (private global copy, the rows a step
//    needs cached in shared memory and rotated locally) -> post the partial -> stream the rest of rotation w.
//  * Dormant orbitals: rows of orbitals that have not yet appeared in any super block are skipped by the update and caught
//    up with the accumulated rotation when they wake (see catch_up).
//  * Signalling is one-directional and tagged: angle records, partials and totals are 16-byte words written with a single
//    store and polled by the consumer; each 8-byte half carries 32 data bits and the tag, so a torn observation is
//    rejected (ll.cuh).  Slot counts follow from how far the parties can drift apart (see the constants).  Every wait has
//    a time-out that aborts the whole kernel (summary[3]); until the first super block has arrived a longer one applies
//    (launch skew between ranks).  All signalling words of the workspace are cleared by the launch; only the peer
//    mailboxes survive a launch and carry the caller's epoch in their tags.
//
// The launch is cooperative only to guarantee that all CTAs are co-resident.  n > 1024, or shapes where a worker would own
// more than four row pieces, stay on the barrier kernel of sweep2.cu (see qio_b200_sweep_cycle).
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "linesearch.cuh"
#include "ll.cuh"
#include "sweep3.cuh"

// The file is compiled twice: as is (no phase timers: the default kernel) and through sweep3_timed.cu with QIO_S3_TIMED
// (namespace s3t): the phase timers that feed summary[4..28] cost registers in a kernel that has none to spare (about 5 % of
// the cycle), so the timed build is only launched when QIO_B200_SWEEP_TIMERS is set (bench.py does that for its phase table).
#ifndef S3_NS
#define S3_NS s3
#define S3_LAUNCH sweep3_launch
#endif

namespace qio {
namespace S3_NS {

constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;
constexpr int SEARCH_THREADS = LS_THREADS;          // 320: warps 0..9 of CTA 0
constexpr int FETCH_THREADS = THREADS - LS_THREADS; // 192: warps 10..15 of CTA 0
constexpr int MAXM = 6;                             // orbitals of three consecutive pairs
constexpr int MAXE = 8;                             // ... plus the two of the rotation being applied
constexpr int SB_MAX = MAXM * MAXM * MAXM * MAXM;   // 1296
constexpr int GM_MAX = 2 * MAXM * MAXM;             // 72 gamma entries (spin, p, q)
constexpr int TOT_WORDS = SB_MAX + GM_MAX;
constexpr int LOOK = 3;          // P_k is taken from the state after rotation k - LOOK
constexpr int SLOTS = 4;         // partial / total slots: a worker writes P_{w+3} only after the reducer finished P_w
constexpr int MB_SLOTS = 8;      // mailbox slots: rank A may post P_{j+2} while rank B still reads P_{j-2}
constexpr int RING = 8;          // angle records: the search CTA is at most LOOK records ahead of the slowest reader
constexpr int MAX_PIECES = 4;
constexpr int TILE_LEN = 96;     // owned columns handled per super-block tile
constexpr int TILE_PITCH = 98;   // row pitch of the tile: 196 words = 4 mod 32, the 8 rows a warp reads fall on different banks
constexpr int DT_STRIDE = 16;     // ints per entry of the per-pair table
constexpr int WORK_THREADS = THREADS - 32;   // a worker's warps 0..14 compute, warp 15 plans the next step (polls the angle)
// The two phases of a worker step touch disjoint rows (the flat phase leaves the rows of the tile alone), so they run
// CONCURRENTLY on two groups of warps: the first Params::tile_threads working threads the latency-bound tile phase (rows of the
// next super block: load, rotate, dot products, post), the others the bandwidth-bound flat phase (everything else of the
// rotation).  A step then lasts max(tile, flat) instead of tile + flat.  Both groups load asynchronously: the tile group with
// cp.async into the tile (one round trip whatever its size), the flat group with cp.async into private staging slots, two
// batches of four butterflies in flight per thread (flat_update_staged).  The requests of the two groups share the SM's memory
// pipeline, and the tile loads -- on the dependency chain -- queue behind the flat ones: deeper flat batches are SLOWER
// (n = 100, one GPU, 128 tile threads: registers-only 41.4 ms per cycle, staged 3 / 4 / 6 per batch 42.3 / 42.7 / 43.6), what
// pays is the larger tile group the cheaper flat phase affords (staged 4: 128 / 160 / 192 / 224 / 256 / 288 / 320 tile threads
// -> 42.7 / 41.7 / 39.8 / 39.0 / 39.1 / 40.4 / 42.8 ms; n = 200: 1133 registers-only -> 1109 ms at 224).
// The split is chosen per launch: 224 tile threads, 256 where a rank's share of the rotation is small (sharded runs).
constexpr int MIN_TILE_THREADS = 128;
constexpr int MAXN3 = 1024;      // largest n of the pipelined kernel (rank / activation tables in shared memory)
constexpr int CU_COLS = 32;       // columns per catch-up chunk (one per lane)
constexpr int CU_PITCH = 36;      // pitch of the staged catch-up vectors (conflict-free B fragments)
// Reducer CTAs (Params::nred): 4 on one rank, 8 when sharded.  A reducer thread polls its share of an entry's column of 142
// partials eight at a time; with 4 CTAs that is three dependent round trips through L2 after the last partial arrived, with 8
// CTAs (16 sub-sums per entry) two -- on the dependency chain of the sharded runs, where four fewer worker CTAs cost nothing.
constexpr int NRED_MAX = 8;

struct Params {
    double *S, *gamma, *U, *W;
    int n, a0, na;
    const int32_t *pairs;
    int P;
    const int32_t *inactive;
    ThetaTablesDev tab;
    qio_b200_ls_result *results;
    double *summary;
    unsigned long long *ring;   // [RING][4]: tagged words of cos and sin; tag = 2 (k + 1) + accepted
    unsigned *abort_flag;       // [0] != 0: some wait timed out, everybody leaves; [1] != 0: the first super block has arrived
    unsigned long long *pbuf;   // [SLOTS][nworkers][SB_MAX][2] tagged partial entries, tag = k + 1 (cleared by the launch)
    unsigned long long *tot;    // [SLOTS][TOT_WORDS][2] tagged totals, tag = k + 1
    unsigned long long timeout_ns, startup_timeout_ns;
    double *blocks_out;         // (P, 24) or NULL: the block each search ran on (decision audit)
    double *trace;              // timed build only, or NULL: [TRACE_PAIRS][TRACE_EVENTS] globaltimer stamps (ns) of the pairs
    int trace_k0;               //   trace_k0 .. trace_k0 + TRACE_PAIRS - 1 (see the TR_* constants)
    int rank, nranks;
    unsigned epoch;             // launch epoch of the mailbox tags (ll.cuh: mailbox_tag)
    double *const *mailboxes;   // [nranks] bases, layout [MB_SLOTS][nranks][SB_MAX][2]
    const int32_t *dtab;        // [P][DT_STRIDE] per-pair table (see build_dtab_kernel)
    int nworkers;               // workers that own at least one column
    int granule;                // doubles per ownership quantum (4, 2 or 1)
    int flat_unr;               // butterflies in flight per thread in the flat phase (3, 4 or 5), register version
    int flat_batch;             // staged flat phase: butterflies per cp.async batch (2, 3 or 4; two batches in flight), 0: register version
    int tile_threads;           // working threads 0 .. tile_threads-1 do the tile phase, the others the flat phase
    int nred;                   // CTAs 1 .. nred reduce, CTA nred + 1 keeps gamma / U / W, workers from CTA nred + 2
    int wc_cols;                // row pitch of a worker's private W columns
    double *wg;                 // [nworkers][n][wc_cols] every worker's own columns of W (global, L2-resident, private)
    double *rg;                 // [nworkers][n][n] every worker's copy of the rotation accumulated in this launch, in
                                // activation-rank coordinates (see the dormant-orbital note at catch_up)
};

__device__ __forceinline__ unsigned long long wtimer() {     // watchdog clock: always real
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Event trace of the timed build: who saw what when, for a window of pairs (the per-pair dependency chain made visible).
constexpr int TRACE_PAIRS = 32, TRACE_EVENTS = 16;
enum { TR_SEARCH_START = 0,   // search CTA: super block k is there, contraction starts
       TR_DECIDED = 1,        // search CTA: angle k published
       TR_W_ANGLE = 2,        // worker 0 planner: angle k seen
       TR_W_START = 3,        // worker 0: step k starts (rotation k, produces super block k + 3)
       TR_W_POSTED = 4,       // worker 0: partial of super block k posted  (stored under pair k, written during step k - 3)
       TR_R_GOT = 5,          // reducer 0: all partials of super block k seen
       TR_R_PUB = 6,          // reducer 0: total of super block k published
       TR_F_GOT = 7,          // prefetch warps: total(s) of super block k seen
       TR_F_DONE = 8,         // prefetch warps: reduced super block k ready (full barrier arrives)
       TR_W_END = 9,          // worker 0: step k done
       TR_F_PASS = 10,        // prefetch warps, thread 0 (the output with the most terms): its part of the contraction done
       TR_F_META = 11,        // ... meta data written
       TR_F_LAST = 12 };      // prefetch warps, last thread (no output of its own): arrives at the closing barrier
template <typename P_>
__device__ __forceinline__ void trace_event(const P_ &p, int k, int ev) {
#ifdef QIO_S3_TIMED
    if (p.trace && k >= p.trace_k0 && k < p.trace_k0 + TRACE_PAIRS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(size_t)(k - p.trace_k0) * TRACE_EVENTS + ev] = (double)t;
    }
#endif
}
__device__ __forceinline__ unsigned long long gtimer() {     // phase timers: only in the QIO_S3_TIMED build
#ifndef QIO_S3_TIMED
    return 0ull;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// asynchronous global -> shared copies (no registers, no stall at the issue): 16 bytes through L2, or 8 bytes
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int ID, int COUNT>
__device__ __forceinline__ void named_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
__device__ __forceinline__ void named_sync_rt(int id, int count) {       // count: a multiple of 32, same in all participants
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
template <int COUNT>
__device__ __forceinline__ void named_sync_id(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory");
}
template <int COUNT>
__device__ __forceinline__ void named_arrive_id(int id) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(COUNT) : "memory");
}

// Spin until pred() holds; false on time-out or when another CTA has aborted.
template <typename Pred>
__device__ __forceinline__ bool spin_until(const Params &p, Pred pred) {
    if (pred()) return true;
    const unsigned long long t0 = wtimer();
    unsigned it = 0;
    for (;;) {
        if (pred()) return true;
        if ((++it & 0xff) == 0) {
            if (ld_volatile_u32(p.abort_flag) != 0) return false;
            if (wtimer() - t0 > (ld_volatile_u32(p.abort_flag + 1) != 0 ? p.timeout_ns : p.startup_timeout_ns)) {
                atomicExch(p.abort_flag, 1u);
                return false;
            }
        }
    }
}

// ---- the orbitals of pairs lo..hi in order of first appearance -----------------------------------------------------
struct DSet {
    int m;
    int D[MAXM];
};
__device__ __forceinline__ void make_dset(DSet &d, const int32_t *pairs, int lo, int hi) {
    d.m = 0;
    for (int k = lo; k <= hi; ++k)
        for (int h = 0; h < 2; ++h) {
            const int x = pairs[2 * k + h];
            bool found = false;
            for (int q = 0; q < d.m; ++q) found |= (d.D[q] == x);
            if (!found) d.D[d.m++] = x;
        }
}
__device__ __forceinline__ int dset_pos(const DSet &d, int x) {
    int r = -1;
    for (int q = 0; q < d.m; ++q)
        if (d.D[q] == x) r = q;
    return r;
}
// super block k covers pairs max(0, k - LOOK + 1) .. k
__device__ __forceinline__ int sb_lo(int k) { return k - (LOOK - 1) > 0 ? k - (LOOK - 1) : 0; }

// Per-pair table, built once per launch from the pair list (everything in it is independent of the angles):
//   [0] m = |D_k|, [1..6] D_k (orbitals of pairs k-2..k in order of first appearance), [7], [8] position of i_k, j_k in D_k,
//   [9], [10] position of (i, j) of pair k-2, [11], [12] of pair k-1 (-1 if that pair does not exist), [13], [14] inactive
//   flags of i_k, j_k.
__global__ void build_dtab_kernel(const int32_t *__restrict__ pairs, const int32_t *__restrict__ inactive, int P,
                                  int32_t *__restrict__ dtab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    DSet d;
    make_dset(d, pairs, sb_lo(k), k);
    int32_t *t = dtab + (size_t)k * DT_STRIDE;
    t[0] = d.m;
    for (int q = 0; q < MAXM; ++q) t[1 + q] = q < d.m ? d.D[q] : -1;
    const int i = pairs[2 * k], j = pairs[2 * k + 1];
    t[7] = dset_pos(d, i);
    t[8] = dset_pos(d, j);
    for (int h = 0; h < 2; ++h) {
        const int v = k - 2 + h;
        t[9 + 2 * h] = v >= 0 ? dset_pos(d, pairs[2 * v]) : -1;
        t[10 + 2 * h] = v >= 0 ? dset_pos(d, pairs[2 * v + 1]) : -1;
    }
    t[13] = inactive[i] != 0;
    t[14] = inactive[j] != 0;
    t[15] = 0;
}

// Angle record k.  Returns false on abort.
__device__ __forceinline__ bool read_record(const Params &p, int k, int &accepted, double &c, double &s) {
    const unsigned long long *w = p.ring + (size_t)(k & (RING - 1)) * 4;
    unsigned long long cb = 0, sbits = 0;
    unsigned t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    const unsigned want = (unsigned)(k + 1);
    const bool ok = spin_until(p, [&]() {
        ll_load_tags(w, cb, t0, t1);
        ll_load_tags(w + 2, sbits, t2, t3);
        return (t0 >> 1) == want && t0 == t1 && t0 == t2 && t0 == t3;
    });
    accepted = (int)(t0 & 1u);
    c = __longlong_as_double((long long)cb);
    s = __longlong_as_double((long long)sbits);
    return ok;
}

// =====================================================================================================================
// Workers
// =====================================================================================================================
struct Piece {
    int slab;       // local slab index
    int d0, len;    // columns [d0, d0 + len) of every d-row of that slab
    int col0;       // first of its columns in the worker's concatenated column space (= in the shared-memory W columns)
    int slabcol;    // column of W[:, a0 + slab] in the shared-memory W columns
    long long base; // slab * n^3 + d0 - col0: element offset of concatenated column 0 of row (0, 0)
};

struct WorkerStep {             // built by thread 0 of a worker for each step, read by all its threads
    int rotated, i, j;
    double c, s;
    int m;                      // orbitals of the super block produced in this step (0: none)
    int D[MAXM];
    int e;                      // tile orbitals: [i, j] (if rotated) followed by the orbitals of D other than i, j
    int E[MAXE];
    int tpos[MAXM];             // position of D_q in E
    int slot;                   // super-block index (k) this step produces, -1 if none
    int alive;
    int nnew, na_before;        // orbitals that wake up in this step (first appearance in a super block), active count before
    int newc[MAXM];
};

struct WorkerSmem {
    Piece piece[MAX_PIECES];
    int npieces, ncols, ncd;    // ncd = owned d-columns (concatenated over the pieces); ncols = ncd + npieces
    WorkerStep st[2];           // double-buffered: the planner warp writes step w + 1 while step w runs
    unsigned long long t_tile, t_wait, t_pass, t_flat, t_d[4];
    unsigned skipmask[2][MAXN3 / 32];   // bit x: the flat phase of that step leaves rows with index x alone (tile phase or dormant)
    unsigned dorm[MAXN3 / 32];  // planner: bit x = orbital x has not appeared in any super block of this launch yet
    short rank[MAXN3];          // activation rank of an orbital (-1: dormant); append-only
    short alist[MAXN3];         // orbital of an activation rank
    int na_act, any_rot;
    short flatlist[2][MAXN3];   // per step: the row indices the flat phase visits (active, not in the tile), ascending
    int nflat[2];
    double dots[MAX_PIECES * MAXM * MAXM * MAXM];       // [piece][q][r][s]
    alignas(16) double tile[MAXE * MAXE * TILE_PITCH];
};

__device__ __forceinline__ int piece_of(const WorkerSmem &sm, int cc) {
    int pc = 0;
#pragma unroll
    for (int k = 1; k < MAX_PIECES; ++k)
        if (k < sm.npieces && cc >= sm.piece[k].col0) pc = k;
    return pc;
}

// Flat phase of a worker: the (axis, x, vector column) items of its owned columns for the `nflat` row indices x of
// `flatlist` (active orbitals outside the tile), strided over `nt` threads (this one is number t): rows (i, x), (j, x)
// (axis 0) or (x, i), (x, j) (axis 1) rotated by (c, s).  A separate function on purpose: its loop wants a clean register file
// (four butterflies = eight independent 16-byte loads in flight per thread).
template <int V, int UNR>
static __device__ __noinline__ void flat_update(double *__restrict__ S, const WorkerSmem *smp, const short *flatlist, int nflat, int n,
                                                int i, int j, double c, double s, int t, int nt) {
    using VT = typename std::conditional<V == 2, double2, double>::type;
    const WorkerSmem &sm = *smp;
    const long long n1 = n, n2 = n1 * n1;
    const int lv = sm.ncd / V;
    const long long b0 = sm.piece[0].base, b1 = sm.piece[1].base, b2 = sm.piece[2].base, b3 = sm.piece[3].base;
    const int c1 = sm.npieces > 1 ? sm.piece[1].col0 : 0x7fffffff, c2 = sm.npieces > 2 ? sm.piece[2].col0 : 0x7fffffff,
              c3 = sm.npieces > 3 ? sm.piece[3].col0 : 0x7fffffff;
    // mixed-radix counter (rho2, v) advanced by nt items; rho2 < 2 nflat: axis = rho2 >= nflat, x = flatlist[rho2 - axis nflat]
    const int dq = nt / lv, dr = nt - dq * lv;
    const long long last = 2LL * nflat * lv;
    long long item = t;
    int rho2 = (int)(item / lv), v = (int)(item - (long long)rho2 * lv);
    const long long safe = b0;          // a valid element offset for the loads of the items beyond the end
    while (item < last) {
        long long off[UNR];
        VT x[UNR], y[UNR];
        unsigned okmask = 0u;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const bool in = item < last;
            const int axis = rho2 >= nflat ? 1 : 0, xx = in ? flatlist[rho2 - axis * nflat] : 0;
            const int cc = v * V;
            const long long pb = cc >= c3 ? b3 : (cc >= c2 ? b2 : (cc >= c1 ? b1 : b0));
            // bit 62 of the offset carries the axis; items beyond the end load (and do not store) a harmless valid element,
            // so that every load is unconditional and the loaded values stay in registers
            off[u] = (in ? (pb + cc + (axis ? xx * n2 : xx * n1)) : safe) | ((long long)axis << 62);
            okmask |= in ? (1u << u) : 0u;
            item += nt;
            v += dr;
            rho2 += dq;
            if (v >= lv) {
                v -= lv;
                ++rho2;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const bool ax = (off[u] >> 62) & 1;
            const long long o = off[u] & ~(1LL << 62);
            const long long si = ax ? (long long)i * n1 : (long long)i * n2, sj = ax ? (long long)j * n1 : (long long)j * n2;
            x[u] = __ldcg(reinterpret_cast<const VT *>(S + o + si));
            y[u] = __ldcg(reinterpret_cast<const VT *>(S + o + sj));
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            if (okmask & (1u << u)) {
                const bool ax = (off[u] >> 62) & 1;
                const long long o = off[u] & ~(1LL << 62);
                const long long si = ax ? (long long)i * n1 : (long long)i * n2, sj = ax ? (long long)j * n1 : (long long)j * n2;
                if constexpr (V == 2) {
                    *reinterpret_cast<VT *>(S + o + si) = make_double2(c * x[u].x + s * y[u].x, c * x[u].y + s * y[u].y);
                    *reinterpret_cast<VT *>(S + o + sj) = make_double2(c * y[u].x - s * x[u].x, c * y[u].y - s * x[u].y);
                } else {
                    *reinterpret_cast<VT *>(S + o + si) = c * x[u] + s * y[u];
                    *reinterpret_cast<VT *>(S + o + sj) = c * y[u] - s * x[u];
                }
            }
    }
}

// The same phase with the loads staged through shared memory: every thread owns 2 x B private 2-row slots and keeps two
// cp.async batches of B butterflies in flight -- batch b + 1 is issued before batch b is waited for, so between B and 2 B
// butterflies per thread are outstanding at any time (the register version drains to zero at every round trip), and the
// data in flight costs no registers, only the element offsets for the write-back.  No barrier: a slot is written (by the
// copy engine) and read by its own thread only.  Item order, arithmetic and stores are those of flat_update.
template <int V, int B>
static __device__ __noinline__ void flat_update_staged(double *__restrict__ S, const WorkerSmem *smp, const short *flatlist, int nflat,
                                                       int n, int i, int j, double c, double s, int t, int nt, void *stage_raw) {
    using VT = typename std::conditional<V == 2, double2, double>::type;
    const WorkerSmem &sm = *smp;
    VT *stage = reinterpret_cast<VT *>(stage_raw) + t;      // slot (batch buffer h, u, row r) at stage[((h B + u) 2 + r) nt]
    const long long n1 = n, n2 = n1 * n1;
    const int lv = sm.ncd / V;
    const long long b0 = sm.piece[0].base, b1 = sm.piece[1].base, b2 = sm.piece[2].base, b3 = sm.piece[3].base;
    const int c1 = sm.npieces > 1 ? sm.piece[1].col0 : 0x7fffffff, c2 = sm.npieces > 2 ? sm.piece[2].col0 : 0x7fffffff,
              c3 = sm.npieces > 3 ? sm.piece[3].col0 : 0x7fffffff;
    const int dq = nt / lv, dr = nt - dq * lv;
    const long long last = 2LL * nflat * lv;
    long long item = t;
    int rho2 = (int)(item / lv), v = (int)(item - (long long)rho2 * lv);
    long long off[2][B];            // element offset of the butterfly's row-pair origin, bit 62 = axis, negative = no item
    auto rows = [&](long long o, long long &si, long long &sj) -> long long {
        const bool ax = (o >> 62) & 1;
        si = ax ? (long long)i * n1 : (long long)i * n2;
        sj = ax ? (long long)j * n1 : (long long)j * n2;
        return o & ~(1LL << 62);
    };
    auto issue = [&](auto hc) -> bool {
        constexpr int h = decltype(hc)::value;
        const bool any = item < last;
#pragma unroll
        for (int u = 0; u < B; ++u) {
            long long o = -1;
            if (item < last) {
                const int axis = rho2 >= nflat ? 1 : 0, xx = flatlist[rho2 - axis * nflat];
                const int cc = v * V;
                const long long pb = cc >= c3 ? b3 : (cc >= c2 ? b2 : (cc >= c1 ? b1 : b0));
                o = (pb + cc + (axis ? xx * n2 : xx * n1)) | ((long long)axis << 62);
                long long si, sj;
                const long long oo = rows(o, si, sj);
                VT *slot = stage + (size_t)((h * B + u) * 2) * nt;
                if constexpr (V == 2) {
                    cp_async_16(slot, S + oo + si);
                    cp_async_16(slot + nt, S + oo + sj);
                } else {
                    cp_async_8(slot, S + oo + si);
                    cp_async_8(slot + nt, S + oo + sj);
                }
            }
            off[h][u] = o;
            item += nt;
            v += dr;
            rho2 += dq;
            if (v >= lv) {
                v -= lv;
                ++rho2;
            }
        }
        cp_async_commit();
        return any;
    };
    auto process = [&](auto hc) {
        constexpr int h = decltype(hc)::value;
#pragma unroll
        for (int u = 0; u < B; ++u)
            if (off[h][u] >= 0) {
                long long si, sj;
                const long long o = rows(off[h][u], si, sj);
                const VT *slot = stage + (size_t)((h * B + u) * 2) * nt;
                const VT x = slot[0], y = slot[nt];
                if constexpr (V == 2) {
                    *reinterpret_cast<VT *>(S + o + si) = make_double2(c * x.x + s * y.x, c * x.y + s * y.y);
                    *reinterpret_cast<VT *>(S + o + sj) = make_double2(c * y.x - s * x.x, c * y.y - s * x.y);
                } else {
                    *reinterpret_cast<VT *>(S + o + si) = c * x + s * y;
                    *reinterpret_cast<VT *>(S + o + sj) = c * y - s * x;
                }
            }
    };
    using H0 = std::integral_constant<int, 0>;
    using H1 = std::integral_constant<int, 1>;
    bool more = issue(H0{});
    while (more) {
        const bool more1 = issue(H1{});
        cp_async_wait_group<1>();
        process(H0{});
        if (!more1) break;
        more = issue(H0{});
        cp_async_wait_group<1>();
        process(H1{});
    }
    cp_async_wait_group<0>();
}

// Dormant orbitals.  An orbital x that has not yet appeared in any super block of this launch cannot take part in a
// rotation, so the flat phase skips every row pair whose free index is x: rows (i|j, x) of axis-1 rotations and rows
// (x, i|j) of axis-2 rotations -- on average a third of the update traffic of a cycle.  What those rows missed is the
// product R of all rotations so far acting on their other index; when x wakes up (three pairs before it is first
// searched) the worker applies it in one go:
//     S[a', y, x, d'] <- sum_{y'} R[y, y'] S[a', y', x, d'],    S[a', x, y, d'] <- sum_{y'} R[y, y'] S[a', x, y', d']
// over the active orbitals y, y' (R is the identity outside them) for its owned columns.  Rg is the worker's private copy
// of R in activation-rank coordinates (rows rank(i), rank(j) are rotated with every accepted pair).  Orbitals still
// dormant at the end of the launch are caught up then.  The input vectors are staged in shared memory first (the update is in place); the small GEMM runs on DMMA tiles.
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

static __device__ __noinline__ void catch_up(const Params &p, const WorkerSmem &sm, const double *Rg, double *cu_in, int x, int A) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = p.n;
    const int g = lane >> 2, t = lane & 3;
    const long long n1 = n, n2 = n1 * n1;
    constexpr int NWW = WORK_THREADS / 32;
    for (int o = 0; o < 2; ++o)
        for (int c0 = 0; c0 < sm.ncd; c0 += CU_COLS) {
            const int nc = min(CU_COLS, sm.ncd - c0);
            // stage the input vectors in[y][c] (pitch CU_PITCH): eight loads in flight per thread
            for (int base = 0; base < A * CU_COLS; base += 8 * WORK_THREADS) {
                double vv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = base + u * WORK_THREADS + tid, y = idx >> 5, c = idx & 31;
                    vv[u] = 0.0;
                    if (idx < A * CU_COLS && c < nc) {
                        const int cc = c0 + c, oy = sm.alist[y];
                        vv[u] = __ldcg(p.S + sm.piece[piece_of(sm, cc)].base + cc + (o == 0 ? oy * n2 + x * n1 : x * n2 + oy * n1));
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = base + u * WORK_THREADS + tid;
                    if (idx < A * CU_COLS) cu_in[(idx >> 5) * CU_PITCH + (idx & 31)] = vv[u];
                }
            }
            named_sync<2, WORK_THREADS>();
            // out[8 rows of a warp][32 columns] = R[8 x A] in[A x 32] on the FP64 tensor pipe (m8n8k4): the A fragment
            // (one element of R per lane) comes straight from L2 and is prefetched one k-step ahead, the B fragments from
            // shared memory (pitch 36: conflict-free)
            for (int y0 = 8 * warp; y0 < A; y0 += 8 * NWW) {
                double acc[4][2];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[ni][0] = acc[ni][1] = 0.0;
                const bool rowok = y0 + g < A;
                const double *rrow = Rg + (size_t)(y0 + g) * n + t;
                double a_next = (rowok && t < A) ? __ldcg(rrow) : 0.0;
                for (int k0 = 0; k0 < A; k0 += 4) {
                    const double a = a_next;
                    a_next = (rowok && k0 + 4 + t < A) ? __ldcg(rrow + k0 + 4) : 0.0;
                    const bool kok = k0 + t < A;
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                        const double bb = kok ? cu_in[(k0 + t) * CU_PITCH + 8 * ni + g] : 0.0;
                        dmma884(acc[ni][0], acc[ni][1], a, bb);
                    }
                }
                if (rowok) {
                    const int oy = sm.alist[y0 + g];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = 8 * ni + 2 * t + h;
                            if (c < nc) {
                                const int cc = c0 + c;
                                p.S[sm.piece[piece_of(sm, cc)].base + cc + (o == 0 ? oy * n2 + x * n1 : x * n2 + oy * n1)] = acc[ni][h];
                            }
                        }
                }
            }
            named_sync<2, WORK_THREADS>();
        }
}

// One worker step: rotation (i, j, c, s) (if accepted) of the owned columns, producing the partial of super block `slot`
// (if any) on the way.  Wg = the owned columns of W (private global copy, rotated here as well); Wr = shared-memory cache of
// its rows E (the only ones a step reads), row t <-> orbital E[t].
template <int V>
__device__ __forceinline__ void worker_step(const Params &p, WorkerSmem &sm, double *Wg, double *Wr, double *Rg, double *cu_in,
                                            int widx, int buf) {
    using VT = typename std::conditional<V == 2, double2, double>::type;
    const int tid = threadIdx.x;
    const int n = p.n, TG = p.tile_threads;
    const long long n1 = n, n2 = n1 * n1;
    const WorkerStep &st = sm.st[buf];
    const bool rot = st.rotated != 0;
    const double c = st.c, s = st.s;
    const int i = st.i, j = st.j, m = st.m, e = st.e, pitch = p.wc_cols;
    const unsigned long long t_begin = (widx == 0 && tid == 0) ? gtimer() : 0;
    if (widx == 0 && tid == 0) trace_event(p, st.slot - LOOK, TR_W_START);
    auto bf = [&](VT &x, VT &y) {
        if constexpr (V == 2) {
            const VT nx_ = make_double2(c * x.x + s * y.x, c * x.y + s * y.y);
            const VT ny_ = make_double2(c * y.x - s * x.x, c * y.y - s * x.y);
            x = nx_;
            y = ny_;
        } else {
            const VT nx_ = c * x + s * y, ny_ = c * y - s * x;
            x = nx_;
            y = ny_;
        }
    };

    // ---- orbitals that wake up in this step: apply what their rows missed (nothing before the first accepted rotation)
    if (st.nnew > 0 && sm.any_rot)
    {
        const unsigned long long tc0 = (widx == 0 && tid == 0) ? gtimer() : 0;
        for (int q = 0; q < st.nnew; ++q) catch_up(p, sm, Rg, cu_in, st.newc[q], st.na_before);
        if (widx == 0 && tid == 0) sm.t_d[2] += gtimer() - tc0;
    }
    // ---- flat group: every other row pair of rotation (i, j) inside the owned columns, while the tile group works ------
    if (tid >= TG) {
        const bool tmf = (widx == 0 && tid == TG);
        // (letting the tile group issue its loads first -- a named barrier here, released after its cp.async loop -- was tried
        // twice: 41.1 -> 44.0 ms per n = 100 cycle with the register flat phase, 39.0 -> 41.0 ms with the staged one)
        const unsigned long long tf0 = tmf ? gtimer() : 0;
        if (rot) {
            const int ft = tid - TG;
            if (p.flat_batch == 4) flat_update_staged<V, 4>(p.S, &sm, sm.flatlist[buf], sm.nflat[buf], n, i, j, c, s, ft, (WORK_THREADS - TG), cu_in);
            else if (p.flat_batch == 3) flat_update_staged<V, 3>(p.S, &sm, sm.flatlist[buf], sm.nflat[buf], n, i, j, c, s, ft, (WORK_THREADS - TG), cu_in);
            else if (p.flat_batch == 2) flat_update_staged<V, 2>(p.S, &sm, sm.flatlist[buf], sm.nflat[buf], n, i, j, c, s, ft, (WORK_THREADS - TG), cu_in);
            else if (p.flat_unr == 4) flat_update<V, 4>(p.S, &sm, sm.flatlist[buf], sm.nflat[buf], n, i, j, c, s, ft, (WORK_THREADS - TG));
            else if (p.flat_unr == 5) flat_update<V, 5>(p.S, &sm, sm.flatlist[buf], sm.nflat[buf], n, i, j, c, s, ft, (WORK_THREADS - TG));
            else flat_update<V, 3>(p.S, &sm, sm.flatlist[buf], sm.nflat[buf], n, i, j, c, s, ft, (WORK_THREADS - TG));
        }
        if (tmf) sm.t_flat += gtimer() - tf0;
        return;
    }
    // ---- tile group from here on --------------------------------------------------------------------------------------
    // ---- accumulated rotation: rows rank(i), rank(j) over the active columns (loads here, stores in w_consume)
    const int nact = st.na_before + st.nnew, ri = rot ? sm.rank[i] : 0, rj = rot ? sm.rank[j] : 0;
    double rx = 0.0, ry = 0.0;
    const bool rrot = rot && tid < nact;
    if (rrot) {
        rx = __ldcg(Rg + (size_t)ri * n + tid);
        ry = __ldcg(Rg + (size_t)rj * n + tid);
    }

    // ---- W columns: rows E into shared memory, rows i, j rotated on the way (and written back).  The loads are issued
    // here and consumed after the tile pass has issued its own: one latency window for both.
    const int f0 = rot ? 2 : 0, nother = m > 0 ? (e - f0) * sm.ncols : 0;
    constexpr int WPRE = 4;
    double wx = 0.0, wy = 0.0, wv[WPRE];
    const bool wrot = rot && tid < sm.ncols;
    if (wrot) {
        wx = __ldcg(Wg + (size_t)i * pitch + tid);
        wy = __ldcg(Wg + (size_t)j * pitch + tid);
    }
#pragma unroll
    for (int k = 0; k < WPRE; ++k) {
        const int idx = tid + k * TG;
        wv[k] = 0.0;
        if (idx < nother) {
            const int t = f0 + idx / sm.ncols, col = idx - (t - f0) * sm.ncols;
            wv[k] = __ldcg(Wg + (size_t)st.E[t] * pitch + col);
        }
    }
    auto w_consume = [&]() {
        auto rot_col = [&](int col, double x, double y) {
            const double xn = c * x + s * y, yn = c * y - s * x;
            Wg[(size_t)i * pitch + col] = xn;
            Wg[(size_t)j * pitch + col] = yn;
            Wr[col] = xn;
            Wr[pitch + col] = yn;
        };
        if (rrot) {
            Rg[(size_t)ri * n + tid] = c * rx + s * ry;
            Rg[(size_t)rj * n + tid] = c * ry - s * rx;
        }
        if (rot)
            for (int col = tid + TG; col < nact; col += TG) {
                const double x = __ldcg(Rg + (size_t)ri * n + col), y = __ldcg(Rg + (size_t)rj * n + col);
                Rg[(size_t)ri * n + col] = c * x + s * y;
                Rg[(size_t)rj * n + col] = c * y - s * x;
            }
        if (wrot) rot_col(tid, wx, wy);
        if (rot)
            for (int col = tid + TG; col < sm.ncols; col += TG)
                rot_col(col, __ldcg(Wg + (size_t)i * pitch + col), __ldcg(Wg + (size_t)j * pitch + col));
#pragma unroll
        for (int k = 0; k < WPRE; ++k) {
            const int idx = tid + k * TG;
            if (idx < nother) {
                const int t = f0 + idx / sm.ncols, col = idx - (t - f0) * sm.ncols;
                Wr[(size_t)t * pitch + col] = wv[k];
            }
        }
        for (int idx = tid + WPRE * TG; idx < nother; idx += TG) {
            const int t = f0 + idx / sm.ncols, col = idx - (t - f0) * sm.ncols;
            Wr[(size_t)t * pitch + col] = __ldcg(Wg + (size_t)st.E[t] * pitch + col);
        }
    };

    // ---- tile phase: the rows (x, y), x, y in E, of all owned columns ------------------------------------------
    // orbits of the rotation inside E x E: X = {IJ} + F (rotated) or X = E (not rotated); an orbit has 1, 2 or 4 rows
    const int m3 = m * m * m, ndot = sm.npieces * m3;
    const unsigned ptag = (unsigned)(st.slot + 1);
    unsigned long long *pslot = p.pbuf + ((size_t)(st.slot & (SLOTS - 1)) * p.nworkers + widx) * SB_MAX * 2;
    // All e x e rows of the chunk arrive asynchronously in ONE round trip whatever the number of threads (cp.async, no
    // registers); the rows the rotation touches are then rotated in shared memory -- the orbit {i, j} x {i, j} (four rows), the
    // orbits {i, j} x f and f x {i, j} (two rows each), same operation order as ever -- and written back.
    auto tile_pass = [&](int t0, int tl) {
        const int tv = tl / V;
        for (int item = tid; item < e * e * tv; item += TG) {
            const int rowidx = item / tv, v = item - rowidx * tv, ea = rowidx / e, eb = rowidx - ea * e;
            const int cc = t0 + v * V;
            const double *g = p.S + sm.piece[piece_of(sm, cc)].base + cc + st.E[ea] * n2 + st.E[eb] * n1;
            VT *dst = reinterpret_cast<VT *>(sm.tile + rowidx * TILE_PITCH) + v;
            if constexpr (V == 2) cp_async_16(dst, g); else cp_async_8(dst, g);
        }
        cp_async_wait_all();
        named_sync_rt(3, TG);
        if (!rot) return;
        const int norb = 1 + 2 * (e - 2);
        for (int item = tid; item < norb * tv; item += TG) {
            const int orb = item / tv, v = item - orb * tv;
            const int cc = t0 + v * V;
            double *gbase = p.S + sm.piece[piece_of(sm, cc)].base + cc;
            auto T = [&](int ea, int eb) { return reinterpret_cast<VT *>(sm.tile + (ea * e + eb) * TILE_PITCH) + v; };
            auto G = [&](int ea, int eb) { return reinterpret_cast<VT *>(gbase + st.E[ea] * n2 + st.E[eb] * n1); };
            if (orb == 0) {
                VT q0 = *T(0, 0), q1 = *T(0, 1), q2 = *T(1, 0), q3 = *T(1, 1);      // rows (i, i), (i, j), (j, i), (j, j)
                bf(q0, q1);       // axis 2 on row i
                bf(q2, q3);       // axis 2 on row j
                bf(q0, q2);       // axis 1 on column i
                bf(q1, q3);       // axis 1 on column j
                *T(0, 0) = q0; *T(0, 1) = q1; *T(1, 0) = q2; *T(1, 1) = q3;
                *G(0, 0) = q0; *G(0, 1) = q1; *G(1, 0) = q2; *G(1, 1) = q3;
            } else if (orb <= e - 2) {
                const int f = orb + 1;                                             // rows (i, f), (j, f): axis 1
                VT q0 = *T(0, f), q2 = *T(1, f);
                bf(q0, q2);
                *T(0, f) = q0; *T(1, f) = q2;
                *G(0, f) = q0; *G(1, f) = q2;
            } else {
                const int f = orb - (e - 2) + 1;                                   // rows (f, i), (f, j): axis 2
                VT q0 = *T(f, 0), q1 = *T(f, 1);
                bf(q0, q1);
                *T(f, 0) = q0; *T(f, 1) = q1;
                *G(f, 0) = q0; *G(f, 1) = q1;
            }
        }
        named_sync_rt(3, TG);
    };
    // dot (piece, q, r, s) over the columns [t0, t0 + tl) of the tile; s varies fastest over the lanes (same tile element:
    // broadcast), the rows a warp reads sit on different banks (TILE_PITCH)
    auto tile_dot = [&](int idx, int t0, int tl) -> double {
        const int pc = idx / m3, qrs = idx - pc * m3, qr = qrs / m, ss = qrs - qr * m, q = qr / m, r = qr - q * m;
        const Piece P = sm.piece[pc];
        const int lo = max(P.col0, t0), hi = min(P.col0 + P.len, t0 + tl);
        const double *row = sm.tile + (st.tpos[q] * e + st.tpos[r]) * TILE_PITCH - t0;
        const double *wrow = Wr + (size_t)st.tpos[ss] * pitch;
        double a0 = 0.0, a1 = 0.0;
        int d = lo;
        for (; d + 1 < hi; d += 2) {
            a0 += row[d] * wrow[d];
            a1 += row[d + 1] * wrow[d + 1];
        }
        if (d < hi) a0 += row[d] * wrow[d];
        return a0 + a1;
    };
    // entry (p, q, r, s) = sum over the pieces of W[D_p, slab] * dot(piece, q, r, s), posted as a tagged 16-byte word
    auto post_entry = [&](int en) {
        const int pp = en / m3, qrs = en - pp * m3;
        const double *wrow = Wr + (size_t)st.tpos[pp] * pitch;
        double acc = 0.0;
        for (int pc = 0; pc < sm.npieces; ++pc) acc += wrow[sm.piece[pc].slabcol] * sm.dots[pc * m3 + qrs];
        ll_store(pslot + 2 * (size_t)en, (unsigned long long)__double_as_longlong(acc), ptag);
    };
    const bool single = sm.ncd <= TILE_LEN;
    const bool tm2 = (widx == 0 && tid == 32);
    unsigned long long t_b2 = tm2 ? gtimer() : 0;
    if (single) {
        tile_pass(0, sm.ncd);
        w_consume();
        named_sync_rt(3, TG);
        if (tm2) {
            const unsigned long long t = gtimer();
            sm.t_pass += t - t_b2;
            t_b2 = t;
        }
        if (m > 0) {
            // four lanes per dot product (columns d = part mod 4), two shuffles, then the entries: all threads take part
            auto dots_quads = [&](auto mc) {
                constexpr int M = decltype(mc)::value, M3 = M * M * M;
                const bool tmd = (widx == 0 && tid == 0);
                unsigned long long td = tmd ? gtimer() : 0;
                if (tmd) sm.t_d[0] += td - t_begin;
                for (int base = 0; base < ndot * 4; base += TG) {
                    const int idx4 = base + tid, idx = idx4 >> 2, part = idx4 & 3;
                    double acc = 0.0;
                    if (idx < ndot) {
                        const int pc = idx / M3, qrs = idx - pc * M3, qr = qrs / M, ss = qrs - qr * M, q = qr / M, r = qr - q * M;
                        const int lo = sm.piece[pc].col0, hi = lo + sm.piece[pc].len;
                        const double *row = sm.tile + (st.tpos[q] * e + st.tpos[r]) * TILE_PITCH;
                        const double *wrow = Wr + (size_t)st.tpos[ss] * pitch;
                        double a0 = 0.0, a1 = 0.0;
                        int d = lo + part;
                        for (; d + 4 < hi; d += 8) {
                            a0 += row[d] * wrow[d];
                            a1 += row[d + 4] * wrow[d + 4];
                        }
                        if (d < hi) a0 += row[d] * wrow[d];
                        acc = a0 + a1;
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    if (part == 0 && idx < ndot) sm.dots[idx] = acc;
                }
                if (tmd) { const unsigned long long t = gtimer(); sm.t_d[1] += t - td; td = t; }
                named_sync_rt(3, TG);
                if (tmd) td = gtimer();
                for (int en = tid; en < M * M3; en += TG) {
                    const int pp = en / M3, qrs = en - pp * M3;
                    const double *wrow = Wr + (size_t)st.tpos[pp] * pitch;
                    double acc = 0.0;
                    for (int pc = 0; pc < sm.npieces; ++pc) acc += wrow[sm.piece[pc].slabcol] * sm.dots[pc * M3 + qrs];
                    ll_store(pslot + 2 * (size_t)en, (unsigned long long)__double_as_longlong(acc), ptag);
                }
                if (tmd) sm.t_d[3] += gtimer() - td;
            };
            switch (m) {
                case 2: dots_quads(std::integral_constant<int, 2>{}); break;
                case 3: dots_quads(std::integral_constant<int, 3>{}); break;
                case 4: dots_quads(std::integral_constant<int, 4>{}); break;
                case 5: dots_quads(std::integral_constant<int, 5>{}); break;
                default: dots_quads(std::integral_constant<int, 6>{}); break;
            }
        }
    } else {
        constexpr int DACC = (MAX_PIECES * MAXM * MAXM * MAXM + MIN_TILE_THREADS - 1) / MIN_TILE_THREADS;
        double dacc[DACC];                           // thread <-> dots (piece, q, r, s) = tid, tid + TG, ...
#pragma unroll
        for (int h = 0; h < DACC; ++h) dacc[h] = 0.0;
        for (int t0 = 0; t0 < sm.ncd; t0 += TILE_LEN) {
            const int tl = min(TILE_LEN, sm.ncd - t0);
            tile_pass(t0, tl);
            if (t0 == 0) w_consume();
            if (m > 0) {
                named_sync_rt(3, TG);
#pragma unroll
                for (int h = 0; h < DACC; ++h)
                    if (tid + h * TG < ndot) dacc[h] += tile_dot(tid + h * TG, t0, tl);
                named_sync_rt(3, TG);        // the tile is overwritten by the next chunk
            }
        }
        if (m > 0) {
#pragma unroll
            for (int h = 0; h < DACC; ++h)
                if (tid + h * TG < ndot) sm.dots[tid + h * TG] = dacc[h];
            named_sync_rt(3, TG);
            for (int en = tid; en < m * m3; en += TG) post_entry(en);
        }
    }
    if (widx == 0 && tid == 0) {
        sm.t_tile += gtimer() - t_begin;
        trace_event(p, st.slot, TR_W_POSTED);
    }

    if (widx == 0 && tid == 0) trace_event(p, st.slot - LOOK, TR_W_END);
}

// The planner warp of a worker: the plan of step w (rotation w, super block w + LOOK); w < 0: the prologue super blocks
// 0 .. LOOK-1 from the initial state (no rotation).
__device__ __forceinline__ void plan_worker_step(const Params &p, WorkerSmem &sm, int w, int buf) {
    WorkerStep &st = sm.st[buf];
    unsigned *skipmask = sm.skipmask[buf];
    const int lane = threadIdx.x & 31;
    int slot = w + LOOK;        // w = -3, -2, -1: super blocks 0, 1, 2 of the initial state
    // Everything that depends on the pair list only comes first and overlaps the wait for the angle -- including the row
    // list of the flat phase, which is only used when the rotation is accepted and is therefore built for that case.
    int m = 0, Dq = -1;
    if (slot < p.P) {
        m = p.dtab[(size_t)slot * DT_STRIDE];
        if (lane < MAXM) Dq = p.dtab[(size_t)slot * DT_STRIDE + 1 + lane];
    } else {
        slot = -1;
    }
    const bool has = lane < m;
    // orbitals that appear in a super block for the first time wake up in this step
    const bool isnew = has && sm.rank[Dq] < 0;
    const unsigned newm = __ballot_sync(0xffffffffu, isnew);
    const int na_before = sm.na_act, nnew = __popc(newm);
    if (isnew) {
        const int r = na_before + __popc(newm & ((1u << lane) - 1u));
        sm.rank[Dq] = (short)r;
        sm.alist[r] = (short)Dq;
        st.newc[r - na_before] = Dq;
        atomicAnd(&sm.dorm[Dq >> 5], ~(1u << (Dq & 31)));
    }
    __syncwarp();
    skipmask[lane] = sm.dorm[lane];             // MAXN3 / 32 = 32 words: one per lane
    int i = 0, j = 0;
    if (w >= 0) {
        i = p.pairs[2 * w];
        j = p.pairs[2 * w + 1];
    }
    // E if the rotation is accepted: [i, j] + the orbitals of D other than i, j, in the order of D; if it is not: D itself
    const bool isI = has && w >= 0 && Dq == i, isJ = has && w >= 0 && Dq == j, other = has && !isI && !isJ;
    const unsigned om = __ballot_sync(0xffffffffu, other);
    const int tpos_acc = isI ? 0 : (isJ ? 1 : 2 + __popc(om & ((1u << lane) - 1u)));
    const int e_acc = 2 + __popc(om);
    __syncwarp();
    // skip mask / row list of the flat phase (accepted case: the tile holds i, j and D)
    if (w >= 0) {
        if (has) atomicOr(&skipmask[Dq >> 5], 1u << (Dq & 31));
        if (lane == 0) {
            atomicOr(&skipmask[i >> 5], 1u << (i & 31));
            atomicOr(&skipmask[j >> 5], 1u << (j & 31));
        }
        __syncwarp();
        const int lo = 32 * lane, nn = p.n;
        const unsigned valid = lo + 32 <= nn ? ~0u : (lo < nn ? ((1u << (nn - lo)) - 1u) : 0u);
        unsigned bits = ~skipmask[lane] & valid;
        const int cnt = __popc(bits);
        int pos = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, pos, off);
            if (lane >= off) pos += up;
        }
        const int total = __shfl_sync(0xffffffffu, pos, 31);
        pos -= cnt;                                   // exclusive prefix
        short *fl = sm.flatlist[buf];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            fl[pos++] = (short)(lo + b);
        }
        if (lane == 0) sm.nflat[buf] = total;
    } else if (lane == 0) {
        sm.nflat[buf] = 0;
    }
    // (an L2 prefetch of the tile rows one, two or four steps ahead from here -- cp.async.bulk.prefetch.L2 per row segment --
    // changed nothing: 39.0 -> 39.1 / 39.1 / 40.0 ms per n = 100 cycle; the tile loads do not wait for DRAM but behind the flat
    // group's requests of the same SM)
    if (has) st.D[lane] = Dq;
    if (lane == 0) {
        st.i = i;
        st.j = j;
        st.m = m;
        st.slot = slot;
        st.nnew = nnew;
        st.na_before = na_before;
        sm.na_act = na_before + nnew;
    }
    // ---- the angle -------------------------------------------------------------------------------------------------
    int acc = 0, alive = 1;
    double c = 1.0, s = 0.0;
    if (w >= 0) {
        if (lane == 0) {
            const unsigned long long tw0 = gtimer();
            alive = read_record(p, w, acc, c, s) ? 1 : 0;
            sm.t_wait += gtimer() - tw0;
            if ((int)blockIdx.x == p.nred + 2) trace_event(p, w, TR_W_ANGLE);
        }
        alive = __shfl_sync(0xffffffffu, alive, 0);
        acc = __shfl_sync(0xffffffffu, acc, 0);
    }
    // what depends on it: the layout of the tile
    const int tpos = acc ? tpos_acc : lane;
    if (has) {
        st.tpos[lane] = tpos;
        if (!acc || other) st.E[tpos] = Dq;
    }
    if (lane == 0) {
        st.alive = alive;
        st.rotated = acc;
        st.c = c;
        st.s = s;
        st.e = acc ? e_acc : m;
        if (acc) {
            st.E[0] = i;
            st.E[1] = j;
        }
    }
}

__device__ __forceinline__ void worker_main(const Params &p, WorkerSmem &sm, double *Wr, int widx) {
    double *Wg = p.wg + (size_t)widx * p.n * p.wc_cols;
    double *Rg = p.rg + (size_t)widx * p.n * p.n;
    double *cu_in = Wr + (size_t)MAXE * p.wc_cols;            // [n][CU_PITCH]
    const int tid = threadIdx.x, n = p.n;
    // ---- owned columns: quanta [lo, hi) of the flat (slab, d / granule) space -------------------------------------
    if (tid == 0) {
        const long long qrow = n / p.granule, Q = (long long)p.na * qrow;
        long long lo = Q * widx / p.nworkers, hi = Q * (widx + 1) / p.nworkers;
        int np = 0, col = 0;
        while (lo < hi && np < MAX_PIECES) {
            const int slab = (int)(lo / qrow);
            const long long end = min(hi, (long long)(slab + 1) * qrow);
            Piece &pc = sm.piece[np++];
            pc.slab = slab;
            pc.d0 = (int)(lo - (long long)slab * qrow) * p.granule;
            pc.len = (int)(end - lo) * p.granule;
            pc.col0 = col;
            pc.base = (long long)slab * n * n * n + pc.d0 - col;
            col += pc.len;
            lo = end;
        }
        sm.ncd = col;
        for (int k = 0; k < np; ++k) sm.piece[k].slabcol = col++;
        sm.npieces = np;
        sm.ncols = col;
    }
    for (int k = tid; k < MAXN3 / 32; k += THREADS) {
        const int lo = 32 * k;
        sm.dorm[k] = lo + 32 <= n ? ~0u : (lo < n ? ((1u << (n - lo)) - 1u) : 0u);     // all orbitals dormant
        sm.skipmask[0][k] = sm.skipmask[1][k] = 0u;
    }
    for (int k = tid; k < MAXN3; k += THREADS) sm.rank[k] = -1;
    if (tid == 0) {
        sm.t_tile = sm.t_wait = sm.t_pass = sm.t_flat = 0;
        sm.t_d[0] = sm.t_d[1] = sm.t_d[2] = sm.t_d[3] = 0;
        sm.na_act = 0;
        sm.any_rot = 0;
    }
    // R = 1
    for (int e = tid; e < n * n; e += THREADS) Rg[e] = (e / n == e % n) ? 1.0 : 0.0;
    __syncthreads();
    // this worker's columns of W into its private copy
    for (int pc = 0; pc < sm.npieces; ++pc) {
        const Piece P = sm.piece[pc];
        for (int e = tid; e < n * P.len; e += THREADS) {
            const int x = e / P.len, d = e - x * P.len;
            Wg[(size_t)x * p.wc_cols + P.col0 + d] = __ldcg(p.W + (size_t)x * n + P.d0 + d);
        }
        for (int x = tid; x < n; x += THREADS) Wg[(size_t)x * p.wc_cols + P.slabcol] = __ldcg(p.W + (size_t)x * n + p.a0 + P.slab);
    }
    __syncthreads();
    const bool timing = (widx == 0 && tid == 0);
    const bool planner = tid >= WORK_THREADS;
    unsigned long long t_plan = 0, t_step = 0, tq = timing ? gtimer() : 0;
    if (planner) plan_worker_step(p, sm, -LOOK, 0);
    __syncthreads();
    for (int w = -LOOK; w < p.P; ++w) {
        const int buf = (w + LOOK) & 1;
        if (!sm.st[buf].alive) return;
        if (timing) {
            const unsigned long long t = gtimer();
            t_plan += t - tq;
            tq = t;
        }
        if (planner) {
            if (w + 1 < p.P) plan_worker_step(p, sm, w + 1, buf ^ 1);     // overlaps step w
        } else if (sm.st[buf].rotated || sm.st[buf].m > 0) {
            if ((n & 1) == 0) worker_step<2>(p, sm, Wg, Wr, Rg, cu_in, widx, buf);
            else worker_step<1>(p, sm, Wg, Wr, Rg, cu_in, widx, buf);
            if (tid == 0 && sm.st[buf].rotated) sm.any_rot = 1;
        }
        if (timing) {
            const unsigned long long t = gtimer();
            t_step += t - tq;
            tq = t;
        }
        __syncthreads();
    }
    // orbitals that never appeared in a pair: their rows still miss every rotation of this launch
    if (sm.any_rot && sm.na_act < n && !planner) {
        const int A = sm.na_act;
        for (int x = 0; x < n; ++x)
            if (sm.rank[x] < 0) catch_up(p, sm, Rg, cu_in, x, A);
    }
    if (timing) {
        p.summary[16] = (double)t_plan;       // waiting at the step barrier (planner not ready / stragglers)
        p.summary[17] = (double)t_step;       // tile phase + flat phase
        p.summary[18] = (double)sm.t_tile;    // tile phase (incl. partial post) of worker 0
        p.summary[19] = (double)sm.t_wait;    // of [16]: inside the wait for the angle record
        p.summary[22] = (double)sm.t_pass;    // tile pass (global loads, rotation, stores) up to the CTA barrier, thread 128
        p.summary[23] = (double)sm.t_flat;
        for (int q = 0; q < 4; ++q) p.summary[25 + q] = (double)sm.t_d[q];    // flat phase as seen by thread 128 (not a dot-product thread)
    }
}

// =====================================================================================================================
// CTAs 1..nred: reduction of the partial super blocks (fixed order), publication of the total.  Reducer r owns the
// entries [ne r / nred, ne (r + 1) / nred).
// =====================================================================================================================
#ifdef S3_ROLE_NOINLINE
#define S3_ROLE static __device__ __noinline__
#else
#define S3_ROLE __device__ __forceinline__
#endif
S3_ROLE void reducer_main(const Params &p, double *scratch /* [THREADS] smem: sub-sums [sub][nl] */, int *flag_sm, int ridx) {
    const int tid = threadIdx.x;
    unsigned long long t_wait = 0, t_red = 0, tq = gtimer();
    if (tid == 0) flag_sm[0] = 0;
    __syncthreads();
    for (int k = 0; k < p.P; ++k) {
        const int slot = k & (SLOTS - 1);
        const int m = p.dtab[(size_t)k * DT_STRIDE];
        const int ne = m * m * m * m, e_lo = ne * ridx / p.nred, e_hi = ne * (ridx + 1) / p.nred, nl = e_hi - e_lo;
        // `ns` sub-sums per entry over interleaved subsets of the workers: every thread has RLD tagged loads in flight and
        // re-polls a batch until all its tags are those of super block k (with 142 workers and ns = 8 the whole column of an
        // entry is two round trips through L2)
        // (12 in flight -- two round trips instead of three with 142 workers -- measured SLOWER end to end: the extra registers
        // cost the worker paths of the same kernel more than the reducers gain; profiles/r2_final_summary.md)
#ifndef S3_RLD
#define S3_RLD 8
#endif
        constexpr int RLD = S3_RLD;
        const int ns = nl > 0 ? max(1, min(16, THREADS / nl)) : 1;
        const unsigned want = (unsigned)(k + 1);
        const unsigned long long *pb = p.pbuf + (size_t)slot * p.nworkers * SB_MAX * 2;
        bool ok = true;
        for (int it = tid; it < nl * ns; it += THREADS) {
            const int el = it % nl, sub = it / nl;
            double a8[RLD];
#pragma unroll
            for (int u = 0; u < RLD; ++u) a8[u] = 0.0;
            for (int w0 = sub; w0 < p.nworkers; w0 += RLD * ns) {
                unsigned long long v[RLD];
                ok &= spin_until(p, [&]() {
                    bool all = true;
#pragma unroll
                    for (int u = 0; u < RLD; ++u) {
                        const int wk = w0 + u * ns;
                        if (wk < p.nworkers) all &= ll_load(pb + ((size_t)wk * SB_MAX + e_lo + el) * 2, v[u], want);
                        else v[u] = 0ull;
                    }
                    return all;
                });
#pragma unroll
                for (int u = 0; u < RLD; ++u) a8[u] += __longlong_as_double((long long)v[u]);
            }
            // fixed summation tree (independent of arrival order)
            static_assert(RLD == 12 || RLD == 8, "reducer: 8 or 12 loads in flight");
            if (RLD == 12) {
#pragma unroll
                for (int u = 0; u < 6; ++u) a8[u] += a8[u + 6];
#pragma unroll
                for (int u = 0; u < 3; ++u) a8[u] += a8[u + 3];
                scratch[sub * nl + el] = (a8[0] + a8[1]) + a8[2];
            } else {
#pragma unroll
                for (int w = 4; w > 0; w >>= 1)
#pragma unroll
                    for (int u = 0; u < w; ++u) a8[u] += a8[u + w];
                scratch[sub * nl + el] = a8[0];
            }
        }
        if (!ok) flag_sm[0] = -1;
        __syncthreads();
        if (tid == 0) {
            const unsigned long long t = gtimer();
            t_wait += t - tq;
            tq = t;
            if (ridx == 0) trace_event(p, k, TR_R_GOT);
        }
        if (flag_sm[0] < 0) return;
        for (int el = tid; el < nl; el += THREADS) {
            double total = scratch[el];
            for (int sub = 1; sub < ns; ++sub) total += scratch[sub * nl + el];
            const int e = e_lo + el;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(total);
            if (p.nranks > 1) {
                const unsigned tag = mailbox_tag(p.epoch, k);
                for (int dst = 0; dst < p.nranks; ++dst) {
                    unsigned long long *w = reinterpret_cast<unsigned long long *>(p.mailboxes[dst]) +
                                            (((size_t)(k & (MB_SLOTS - 1)) * p.nranks + p.rank) * SB_MAX + e) * 2;
                    ll_store(w, bits, tag);
                }
            } else {
                ll_store(p.tot + ((size_t)slot * TOT_WORDS + e) * 2, bits, (unsigned)(k + 1));
            }
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned long long t = gtimer();
            t_red += t - tq;
            tq = t;
            if (ridx == 0) trace_event(p, k, TR_R_PUB);
        }
    }
    if (ridx == 0 && tid == 0) {
        p.summary[20] = (double)t_wait;       // reducer 0: waiting for the partials
        p.summary[21] = (double)t_red;        // reducer 0: summing + publishing
    }
}

// =====================================================================================================================
// CTA NRED + 1: gamma, U, W (global) and the gamma part of the super blocks
// =====================================================================================================================
S3_ROLE void aux_main(const Params &p, int *flag_sm, double *cs_sm) {
    const int tid = threadIdx.x, n = p.n;
    const size_t ld = 2 * (size_t)n;
    auto publish_gamma = [&](int k) {       // gamma entries of super block k from the current gamma
        if (k >= p.P) return;
        if (tid < 1 + MAXM) flag_sm[1 + tid] = p.dtab[(size_t)k * DT_STRIDE + tid];
        __syncthreads();
        const int m = flag_sm[1];
        if (tid < 2 * m * m) {
            const int spin = tid / (m * m), r = tid - spin * m * m, pp = r / m, qq = r - pp * m;
            const double v = __ldcg(p.gamma + (2 * (size_t)flag_sm[2 + pp] + spin) * ld + 2 * flag_sm[2 + qq] + spin);
            ll_store(p.tot + ((size_t)(k & (SLOTS - 1)) * TOT_WORDS + SB_MAX + tid) * 2,
                     (unsigned long long)__double_as_longlong(v), (unsigned)(k + 1));
        }
        __syncthreads();
    };
    for (int b = 0; b < LOOK; ++b) publish_gamma(b);
    for (int w = 0; w < p.P; ++w) {
        if (tid == 0) {
            int acc;
            double c, s;
            const bool ok = read_record(p, w, acc, c, s);
            flag_sm[0] = ok ? acc : -1;
            cs_sm[0] = c;
            cs_sm[1] = s;
        }
        __syncthreads();
        const int acc = flag_sm[0];
        if (acc < 0) return;
        const double c = cs_sm[0], s = cs_sm[1];
        const int i = p.pairs[2 * w], j = p.pairs[2 * w + 1];
        if (acc) {
            // gamma <- V gamma V^T on the same-spin blocks: columns, then rows;  U <- V U;  W <- V W
            for (int k = tid; k < 2 * n; k += THREADS) {
                const int a = k >> 1, spin = k & 1;
                double *row = p.gamma + (2 * (size_t)a + spin) * ld;
                const double x = __ldcg(row + 2 * i + spin), y = __ldcg(row + 2 * j + spin);
                row[2 * i + spin] = c * x + s * y;
                row[2 * j + spin] = c * y - s * x;
            }
            for (int k = tid; k < n; k += THREADS) {
                double *ui = p.U + (size_t)i * n + k, *uj = p.U + (size_t)j * n + k;
                const double x = __ldcg(ui), y = __ldcg(uj);
                *ui = c * x + s * y;
                *uj = c * y - s * x;
                double *wi = p.W + (size_t)i * n + k, *wj = p.W + (size_t)j * n + k;
                const double x2 = __ldcg(wi), y2 = __ldcg(wj);
                *wi = c * x2 + s * y2;
                *wj = c * y2 - s * x2;
            }
            __syncthreads();
            for (int k = tid; k < 2 * n; k += THREADS) {
                const int b = k >> 1, spin = k & 1;
                double *gi = p.gamma + (2 * (size_t)i + spin) * ld + 2 * b + spin, *gj = p.gamma + (2 * (size_t)j + spin) * ld + 2 * b + spin;
                const double x = __ldcg(gi), y = __ldcg(gj);
                *gi = c * x + s * y;
                *gj = c * y - s * x;
            }
            __syncthreads();
        }
        publish_gamma(w + LOOK);
    }
}

// =====================================================================================================================
// CTA 0: prefetch (warps 10..15) and search (warps 0..9)
// =====================================================================================================================
constexpr int MAXY = 4;                              // orbitals of pairs k-1 and k: what is left after the fetch warps' contraction
constexpr int SBP_WORDS = MAXY * MAXY * MAXY * MAXY + 2 * MAXY * MAXY;   // reduced super block + its gamma part

struct SearchSmem {
    LineSearchSmemT<SEARCH_THREADS> ls;
    double sb[TOT_WORDS];           // super block (+ gamma part) being fetched; consumed by the fetch warps themselves
    double sbp[2][SBP_WORDS];       // double-buffered REDUCED super block of the pair being / to be searched: rotation k-2 applied,
                                    // restricted to the <= 4 orbitals Y of pairs k-1 and k: [y1][y2][y3][y4] (stride 4), then gamma [spin][y][y']
    int sb_ok[2];
    double blk[QIO_B200_BLOCK_DOUBLES];
    PairCoef pc;
    double2 coarse_cs[320];
    int32_t fine_len[320];
    int meta[2][8];                 // per buffer: Y index of i_k, of j_k, inactive(i), inactive(j), Y index of i_{k-1}, j_{k-1} (-1: no such pair)
    // angle history for the fetch warps: written by search thread 0 the moment pair k is decided
    int hist_acc[4];
    double hist_c[4], hist_s[4];
    double ctab[2][25];             // c^x s^y (x + 5 y) of the rotation the prefetch warps apply, per buffer
    int decided;                    // last decided pair (-1 at start); volatile accesses
    int fetch_abort;
};

// named barriers: 1 = line search (320), 2 + b = full[b], 4 + b = empty[b] (512: one side arrives, the other syncs),
// 6 = prefetch warps among themselves (192)
// The prefetch warps: gather super block k (the workers' / ranks' totals), then take the FIRST of the two pending rotations
// (pair k-2, known since the end of search k-2) off the search threads' critical path:
//     P'[y1,y2,y3,y4] = sum_{pqrs} G0[Y_1,p] G0[Y_2,q] G0[Y_3,r] G0[Y_4,s] P_k[p,q,r,s],   Y = orbitals of pairs k-1 and k (<= 4),
// G0 = rotation k-2 in the orbitals D_k (rows of i_{k-2}, j_{k-2} have two non-zeros, every other row is a unit vector), as four
// single-index passes; likewise for the gamma part.  The search threads are left with rotation k-1 on <= 4^4 numbers.
S3_ROLE void fetch_main(const Params &p, SearchSmem &sm) {
    const int ft = threadIdx.x - SEARCH_THREADS;
    for (int k = 0; k < p.P; ++k) {
        const int b = k & 1;
        if (k >= 2) named_sync_id<THREADS>(4 + b);          // buffer b has been consumed (pair k - 2)
        const int32_t *dt = p.dtab + (size_t)k * DT_STRIDE;
        const int m_k = dt[0];
        const int ne = m_k * m_k * m_k * m_k, ng = 2 * m_k * m_k;
        const int pos_i = dt[7], pos_j = dt[8], pa0 = dt[9], pb0 = dt[10], pa1 = dt[11], pb1 = dt[12];
        const int in_i = dt[13], in_j = dt[14];       // (loaded here, not after the contraction: an L2 round trip on the chain otherwise)
        bool ok = true;
#ifdef S3_FETCH_BATCH       // (all words of a thread in one round trip: measured slower, same reason as S3_RLD = 12)
        if (p.nranks == 1) {
            // one rank: every word this thread is responsible for (<= 8) is polled in the same round trip
            constexpr int FMAX = (TOT_WORDS + FETCH_THREADS - 1) / FETCH_THREADS;
            const unsigned long long *tb = p.tot + (size_t)(k & (SLOTS - 1)) * TOT_WORDS * 2;
            unsigned long long v[FMAX];
            ok &= spin_until(p, [&]() {
                bool all = true;
#pragma unroll
                for (int u = 0; u < FMAX; ++u) {
                    const int e = ft + u * FETCH_THREADS;
                    if (e < ne + ng) all &= ll_load(tb + (size_t)(e >= ne ? SB_MAX + (e - ne) : e) * 2, v[u], (unsigned)(k + 1));
                }
                return all;
            });
#pragma unroll
            for (int u = 0; u < FMAX; ++u) {
                const int e = ft + u * FETCH_THREADS;
                if (e < ne + ng) sm.sb[e >= ne ? SB_MAX + (e - ne) : e] = __longlong_as_double((long long)v[u]);
            }
        } else
#endif
        for (int e = ft; e < ne + ng; e += FETCH_THREADS) {
            const bool is_g = e >= ne;
            const int word = is_g ? SB_MAX + (e - ne) : e;
            double acc = 0.0;
            if (p.nranks > 1 && !is_g) {
                // the ranks' totals, eight loads in flight, added in rank order (identical on every rank)
                const unsigned tag = mailbox_tag(p.epoch, k);
                const unsigned long long *w0 = reinterpret_cast<const unsigned long long *>(p.mailboxes[p.rank]) +
                                               ((size_t)(k & (MB_SLOTS - 1)) * p.nranks * SB_MAX + e) * 2;
                for (int r0 = 0; r0 < p.nranks; r0 += 8) {
                    unsigned long long v[8];
                    ok &= spin_until(p, [&]() {
                        bool all = true;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (r0 + u < p.nranks) all &= ll_load(w0 + (size_t)(r0 + u) * SB_MAX * 2, v[u], tag);
                        return all;
                    });
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (r0 + u < p.nranks) acc += __longlong_as_double((long long)v[u]);
                }
            } else {
                const unsigned long long *w = p.tot + ((size_t)(k & (SLOTS - 1)) * TOT_WORDS + word) * 2;
                unsigned long long v = 0;
                ok &= spin_until(p, [&]() { return ll_load(w, v, (unsigned)(k + 1)); });
                acc = __longlong_as_double((long long)v);
            }
            sm.sb[word] = acc;
        }
        // rotation k-2 (identity if that pair does not exist or was not accepted)
        int acc0 = 0;
        double c0 = 1.0, s0 = 0.0;
        if (k >= 2 && ok) {
            ok &= spin_until(p, [&]() { return *reinterpret_cast<volatile int *>(&sm.decided) >= k - 2; });
            acc0 = sm.hist_acc[(k - 2) & 3];
            c0 = sm.hist_c[(k - 2) & 3];
            s0 = sm.hist_s[(k - 2) & 3];
        }
        // every coefficient of the contraction below is +- c0^x s0^y with x + y <= 4: 25 numbers, one thread each
        if (ft < 25) {
            const int x = ft % 5, y = ft / 5;
            double v = 1.0;
            for (int q = 0; q < x; ++q) v *= c0;
            for (int q = 0; q < y; ++q) v *= s0;
            sm.ctab[b][ft] = v;
        }
        if (!ok) sm.fetch_abort = 1;
        if (k == 0 && ft == 0 && ok) atomicExch(p.abort_flag + 1, 1u);      // start-up is over: the normal time-out applies
        named_sync<6, FETCH_THREADS>();
        if (ft == 0) trace_event(p, k, TR_F_GOT);       // every word of the super block (and angle k - 2) is here
        if (!sm.fetch_abort) {
            // Y: the orbitals of pairs k-1 and k, as positions in D_k, in order of first appearance
            // (kept in scalars and selected with compares: a runtime-indexed array would live in local memory)
            int Y0 = -1, Y1 = -1, Y2 = -1, Y3 = -1, nY = 0;
            auto add = [&](int cand) {
                if (cand < 0 || cand == Y0 || cand == Y1 || cand == Y2 || cand == Y3) return;
                if (nY == 0) Y0 = cand;
                else if (nY == 1) Y1 = cand;
                else if (nY == 2) Y2 = cand;
                else Y3 = cand;
                ++nY;
            };
            add(pa1);
            add(pb1);
            add(pos_i);
            add(pos_j);
            auto yidx = [&](int pos) { return pos == Y0 ? 0 : (pos == Y1 ? 1 : (pos == Y2 ? 2 : (pos == Y3 ? 3 : -1))); };
            // row Y[y] of G0: cA e_pA + cB e_pB
            auto g0row = [&](int y, int &pA, double &cA, int &pB, double &cB) {
                const int pos = y == 0 ? Y0 : (y == 1 ? Y1 : (y == 2 ? Y2 : Y3));
                if (acc0 && pa0 >= 0 && pos == pa0) { pA = pa0; cA = c0; pB = pb0; cB = s0; }
                else if (acc0 && pa0 >= 0 && pos == pb0) { pA = pa0; cA = -s0; pB = pb0; cB = c0; }
                else { pA = pos; cA = 1.0; pB = pos; cB = 0.0; }
            };
            // ONE pass: every output P'[y1,y2,y3,y4] is the sum of <= 16 products c1 c2 c3 c4 P[p1,p2,p3,p4] over the (at most two)
            // non-zeros of the four rows of G0 -- in a run of pairs only the rows of the shared orbital have two, so most
            // outputs are one or two terms.  (Four single-index passes with a barrier each took 2.7 us next to the busy
            // search warps and sat on the dependency chain of the sharded runs.)
            const int m = m_k, n1 = nY, n2 = nY * nY, n3 = n2 * nY, m2 = m * m, m3 = m2 * m;
            for (int o = ft; o < n1 * n3 + 2 * n2; o += FETCH_THREADS) {
                if (o < n1 * n3) {
                    const int y1 = o / n3, r1 = o - y1 * n3, y2 = r1 / n2, r2 = r1 - y2 * n2, y3 = r2 / n1, y4 = r2 - y3 * n1;
                    // rows in coded form: term A / B = position, exponent code (x + 5 y of c0^x s0^y) and sign; nt = 1 or 2 terms.
                    // A product of four coefficients is then a table entry (sum of the codes) with a sign: ONE FP64 instruction per
                    // term instead of four -- the FP64 pipe of this SM belongs to the search warps, and the output with all four
                    // indices on the shared orbital of a run (16 terms) used to take 2 us on the dependency chain.
                    auto coded = [&](int y, int &pA, int &eA, int &gA, int &pB, int &eB, int &nt) {
                        const int pos = y == 0 ? Y0 : (y == 1 ? Y1 : (y == 2 ? Y2 : Y3));
                        if (acc0 && pa0 >= 0 && pos == pa0) { pA = pa0; eA = 1; gA = 0; pB = pb0; eB = 5; nt = 2; }          // c0 e_a + s0 e_b
                        else if (acc0 && pa0 >= 0 && pos == pb0) { pA = pa0; eA = 5; gA = 1; pB = pb0; eB = 1; nt = 2; }    // -s0 e_a + c0 e_b
                        else { pA = pos; eA = 0; gA = 0; pB = pos; eB = 0; nt = 1; }
                    };
                    int pA0, eA0, gA0, pB0, eB0, t0, pA1, eA1, gA1, pB1, eB1, t1, pA2, eA2, gA2, pB2, eB2, t2, pA3, eA3, gA3, pB3, eB3, t3;
                    coded(y1, pA0, eA0, gA0, pB0, eB0, t0);
                    coded(y2, pA1, eA1, gA1, pB1, eB1, t1);
                    coded(y3, pA2, eA2, gA2, pB2, eB2, t2);
                    coded(y4, pA3, eA3, gA3, pB3, eB3, t3);
                    const double *ct = sm.ctab[b];
                    double acc = 0.0;
                    for (int a = 0; a < t0; ++a) {
                        const int e1 = a ? eB0 : eA0, g1 = a ? 0 : gA0, q1 = (a ? pB0 : pA0) * m3;
                        for (int bq = 0; bq < t1; ++bq) {
                            const int e2 = e1 + (bq ? eB1 : eA1), g2 = g1 ^ (bq ? 0 : gA1), q2 = q1 + (bq ? pB1 : pA1) * m2;
                            double part = 0.0;
                            for (int cq = 0; cq < t2; ++cq) {
                                const int e3 = e2 + (cq ? eB2 : eA2), g3 = g2 ^ (cq ? 0 : gA2), q3 = q2 + (cq ? pB2 : pA2) * m;
                                for (int d = 0; d < t3; ++d) {
                                    const double cf = ct[e3 + (d ? eB3 : eA3)];
                                    const int neg = g3 ^ (d ? 0 : gA3);
                                    const double sgn = __hiloint2double(__double2hiint(cf) ^ (neg << 31), __double2loint(cf));
                                    part = fma(sgn, sm.sb[q3 + (d ? pB3 : pA3)], part);
                                }
                            }
                            acc += part;
                        }
                    }
                    sm.sbp[b][((y1 * MAXY + y2) * MAXY + y3) * MAXY + y4] = acc;
                } else {
                    const int og = o - n1 * n3, spin = og / n2, r = og - spin * n2, y = r / n1, yy = r - y * n1;
                    int pA, pB, qA, qB;
                    double cA, cB, dA, dB;
                    g0row(y, pA, cA, pB, cB);
                    g0row(yy, qA, dA, qB, dB);
                    const double *gm = sm.sb + SB_MAX + spin * m2;
                    sm.sbp[b][MAXY * MAXY * MAXY * MAXY + spin * MAXY * MAXY + y * MAXY + yy] =
                        cA * (dA * gm[pA * m + qA] + dB * gm[pA * m + qB]) + cB * (dA * gm[pB * m + qA] + dB * gm[pB * m + qB]);
                }
            }
            if (ft == 0) trace_event(p, k, TR_F_PASS);
            if (ft == 0) {
                int *mt = sm.meta[b];
                mt[0] = yidx(pos_i);
                mt[1] = yidx(pos_j);
                mt[2] = in_i;
                mt[3] = in_j;
                mt[4] = pa1 >= 0 ? yidx(pa1) : -1;
                mt[5] = pb1 >= 0 ? yidx(pb1) : -1;
            }
            if (ft == 0) trace_event(p, k, TR_F_META);
            if (ft == 80) trace_event(p, k, TR_F_LAST);       // a thread whose output has a single term (all rows unit vectors)
        }
        named_sync<6, FETCH_THREADS>();
        if (ft == 0) {
            sm.sb_ok[b] = sm.fetch_abort ? 0 : 1;
            trace_event(p, k, TR_F_DONE);
        }
        __threadfence_block();
        named_arrive_id<THREADS>(2 + b);                    // full[b]
        if (sm.fetch_abort) {
            // keep the barrier protocol matched: the searcher leaves after seeing sb_ok = 0 and arrives nowhere else
            return;
        }
    }
}

S3_ROLE void search_main(const Params &p, SearchSmem &sm) {
    const int tid = threadIdx.x;
    ThetaTablesDev tab = p.tab;
    if (tab.n_coarse <= 320) {
        for (int t = tid; t < tab.n_coarse; t += SEARCH_THREADS) {
            sm.coarse_cs[t] = reinterpret_cast<const double2 *>(tab.coarse_cs)[t];
            sm.fine_len[t] = tab.fine_len[t];
        }
        tab.coarse_cs = reinterpret_cast<const double *>(sm.coarse_cs);
        tab.fine_len = sm.fine_len;
    }
    if (tid < 8) sm.ls.prof[tid] = 0;
    int accepted_total = 0, flags_or = 0;
    // the last LOOK - 1 = 2 angles (accepted, c, s), newest last
    int h_acc[2] = {0, 0};
    double h_c[2] = {1.0, 1.0}, h_s[2] = {0.0, 0.0};
    unsigned long long t_wait = 0, t_setup = 0, t_search = 0, t_pub = 0, tq = gtimer();
    bool alive = true;
    for (int k = 0; k < p.P && alive; ++k) {
        const int b = k & 1;
        named_sync_id<THREADS>(2 + b);                      // full[b]: super block k is in sm.sb[b]
        if (!sm.sb_ok[b]) {
            alive = false;
            break;
        }
        if (tid == 0) {
            const unsigned long long t = gtimer();
            t_wait += t - tq;
            tq = t;
            trace_event(p, k, TR_SEARCH_START);
        }
        const int *mt = sm.meta[b];
        const bool i_in = mt[2] != 0, j_in = mt[3] != 0;
        {   // B = G1^(x4) P': the pending rotation k-1 on the reduced super block.  Target orbital t (0: i_k, 1: j_k) is the
            // combination cf[t][0] Y[yy[t][0]] + cf[t][1] Y[yy[t][1]] of the orbitals before rotation k-1.
            if (tid < QIO_B200_BLOCK_DOUBLES) {
                const int ya1 = mt[4], yb1 = mt[5];
                const bool r1 = h_acc[1] && ya1 >= 0;
                // coefficients / Y indices of target 0 (i_k) and 1 (j_k); selected with compares, never indexed at run time
                double c00, c01, c10, c11;
                int y00, y01, y10, y11;
                auto target = [&](int yx, double &cA, int &yA, double &cB, int &yB) {
                    if (r1 && yx == ya1) { cA = h_c[1]; yA = ya1; cB = h_s[1]; yB = yb1; }
                    else if (r1 && yx == yb1) { cA = -h_s[1]; yA = ya1; cB = h_c[1]; yB = yb1; }
                    else { cA = 1.0; yA = yx; cB = 0.0; yB = yx; }
                };
                target(mt[0], c00, y00, c01, y01);
                target(mt[1], c10, y10, c11, y11);
                auto cf = [&](int t, int a) { return t ? (a ? c11 : c10) : (a ? c01 : c00); };
                auto yy = [&](int t, int a) { return t ? (a ? y11 : y10) : (a ? y01 : y00); };
                const double *sp = sm.sbp[b];
                double acc = 0.0;
                if (tid < 16) {
                    const int al = tid >> 3, be = (tid >> 2) & 1, ga = (tid >> 1) & 1, de = tid & 1;
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int r = 0; r < 2; ++r)
#pragma unroll
                                for (int d = 0; d < 2; ++d)
                                    acc += ((cf(al, a) * cf(be, q)) * (cf(ga, r) * cf(de, d))) *
                                           sp[((yy(al, a) * MAXY + yy(be, q)) * MAXY + yy(ga, r)) * MAXY + yy(de, d)];
                } else {
                    const int u = tid - 16, spin = u >> 2, al = (u >> 1) & 1, be = u & 1;
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            acc += (cf(al, a) * cf(be, q)) * sp[MAXY * MAXY * MAXY * MAXY + spin * MAXY * MAXY + yy(al, a) * MAXY + yy(be, q)];
                }
                sm.blk[tid] = acc;
            }
            ls_sync<SEARCH_THREADS>();
            if (tid == 0) {
                make_coef(sm.blk, sm.pc);
                sm.ls.flags = 0;
            }
            if (p.blocks_out && tid >= 32 && tid < 32 + QIO_B200_BLOCK_DOUBLES)
                p.blocks_out[(size_t)k * QIO_B200_BLOCK_DOUBLES + (tid - 32)] = sm.blk[tid - 32];
            ls_sync<SEARCH_THREADS>();
        }
        named_arrive_id<THREADS>(4 + b);                    // empty[b]: the prefetch warps may overwrite sm.sb[b]
        if (tid == 0) {
            const unsigned long long t = gtimer();
            t_setup += t - tq;
            tq = t;
        }
        qio_b200_ls_result res;
        // the angle record is published by thread 0 the moment the scan has decided (before the margins are taken)
        auto publish = [&](bool accepted, double cth, double sth) {
            sm.hist_acc[k & 3] = accepted ? 1 : 0;          // for the prefetch warps (rotation k is their G0 of pair k + 2)
            sm.hist_c[k & 3] = cth;
            sm.hist_s[k & 3] = sth;
            __threadfence_block();
            *reinterpret_cast<volatile int *>(&sm.decided) = k;
            trace_event(p, k, TR_DECIDED);
            unsigned long long *w = p.ring + (size_t)(k & (RING - 1)) * 4;
            const unsigned tag = 2u * (unsigned)(k + 1) + (accepted ? 1u : 0u);
            ll_store(w, (unsigned long long)__double_as_longlong(cth), tag);
            ll_store(w + 2, (unsigned long long)__double_as_longlong(sth), tag);
        };
        cta_linesearch_t<SEARCH_THREADS, PairCoef, decltype(publish), false, false>(sm.pc, i_in, j_in, tab, sm.ls, res, publish);
        if (tid == 0) {
            const unsigned long long t = gtimer();
            t_search += t - tq;
            tq = t;
            p.results[k] = res;
            const unsigned long long t2 = gtimer();
            t_pub += t2 - tq;
            tq = t2;
        }
        h_acc[0] = h_acc[1];
        h_c[0] = h_c[1];
        h_s[0] = h_s[1];
        h_acc[1] = res.accepted;
        h_c[1] = res.cos_theta;
        h_s[1] = res.sin_theta;
        accepted_total += res.accepted;
        flags_or |= res.flags;
    }
    if (tid == 0) {
        p.summary[0] = (double)accepted_total;
        p.summary[2] = (double)flags_or;
        p.summary[3] = alive ? 0.0 : 1.0;
        p.summary[5] = (double)t_wait;        // waiting for the super block (prefetch not ready)
        p.summary[6] = (double)t_search;      // line search
        p.summary[7] = (double)t_pub;         // record + result stores
        p.summary[8] = 0.0;
        p.summary[9] = (double)t_setup;       // T, contraction, coefficients
        for (int q = 0; q < 6; ++q) p.summary[10 + q] = (double)sm.ls.prof[q];
        p.summary[24] = 3.0;                  // engine id
    }
}

__global__ void __launch_bounds__(THREADS, 1) sweep3_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int cta = blockIdx.x;
    if (cta == 0) {
        SearchSmem &sm = *reinterpret_cast<SearchSmem *>(dyn);
        if (threadIdx.x == 0) {
            sm.fetch_abort = 0;
            sm.decided = -1;
        }
        __syncthreads();
        if (threadIdx.x < SEARCH_THREADS) search_main(p, sm);
        else fetch_main(p, sm);
    } else if (cta <= p.nred) {
        double *scratch = reinterpret_cast<double *>(dyn);
        int *flags = reinterpret_cast<int *>(scratch + THREADS);
        reducer_main(p, scratch, flags, cta - 1);
    } else if (cta == p.nred + 1) {
        int *flags = reinterpret_cast<int *>(dyn);
        double *cs = reinterpret_cast<double *>(dyn + 64);
        aux_main(p, flags, cs);
    } else {
        const int widx = cta - (p.nred + 2);
        if (widx >= p.nworkers) return;
        WorkerSmem &sm = *reinterpret_cast<WorkerSmem *>(dyn);
        double *Wr = reinterpret_cast<double *>(dyn + ((sizeof(WorkerSmem) + 15) / 16) * 16);
        worker_main(p, sm, Wr, widx);
    }
}

struct Layout {
    size_t off_ring, off_count, off_abort, off_tot, off_pbuf, off_dtab, off_wg, off_rg, total;
};
static Layout ws_layout(int max_workers, int n) {
    Layout L;
    L.off_ring = 0;
    L.off_count = L.off_ring + sizeof(unsigned long long) * RING * 4;
    L.off_abort = L.off_count + 64;
    L.off_tot = L.off_abort + 64;
    L.off_pbuf = L.off_tot + sizeof(unsigned long long) * 2 * SLOTS * TOT_WORDS;
    L.off_dtab = L.off_pbuf + sizeof(unsigned long long) * 2 * (size_t)SLOTS * max_workers * SB_MAX;
    L.off_wg = L.off_dtab + sizeof(int32_t) * DT_STRIDE * ((size_t)n * (n - 1) / 2 + 1);
    L.off_wg = (L.off_wg + 255) / 256 * 256;
    // private W columns: every worker holds at most ceil(n^2 / workers') + MAX_PIECES columns (na <= n); bound with the
    // smallest worker count that can occur (Q >= workers) -- n (n + MAX_PIECES + 4) columns in total cover every case
    L.off_rg = L.off_wg + sizeof(double) * (size_t)n * ((size_t)n * n + (size_t)max_workers * (MAX_PIECES + 8));
    L.total = L.off_rg + sizeof(double) * (size_t)max_workers * n * n;       // private accumulated rotations
    return L;
}

}  // namespace S3_NS

#ifndef QIO_S3_TIMED
size_t sweep3_workspace_bytes(int n) { return s3::ws_layout(sm_count(), n).total; }
size_t sweep3_mailbox_bytes(int nranks) {
    return sizeof(unsigned long long) * 2 * (size_t)s3::MB_SLOTS * (size_t)(nranks > 0 ? nranks : 1) * s3::SB_MAX;
}
#endif

// Shared memory of a worker for this problem, or 0 if the pipelined kernel cannot take it.
static size_t worker_smem_bytes(int n, int na, int grid, int nred, int &nworkers, int &granule, int &wc_cols) {
    using namespace S3_NS;
    const int first_worker = nred + 2;
    if (grid < first_worker + 1 || n > MAXN3 || na < 1) return 0;
#ifndef S3_GRANULE_MAX
#define S3_GRANULE_MAX 4
#endif
    granule = (n % 4 == 0 && S3_GRANULE_MAX >= 4) ? 4 : ((n % 2 == 0) ? 2 : 1);
    const long long qrow = n / granule, Q = (long long)na * qrow;
    nworkers = (int)std::min<long long>(grid - first_worker, Q);
    // the longest owned range, and the most pieces it can break into
    const long long per = (Q + nworkers - 1) / nworkers;
    const long long pieces = (per + qrow - 1) / qrow + 1;
    if (pieces > MAX_PIECES) return 0;
    wc_cols = (int)(per * granule + pieces);
    wc_cols += (wc_cols & 1);      // even pitch
    return ((sizeof(WorkerSmem) + 15) / 16) * 16 + sizeof(double) * ((size_t)MAXE * wc_cols + (size_t)n * CU_PITCH);
}

#ifndef QIO_S3_TIMED
bool sweep3_supported(int n, int na) {
    int nw, g, cols;
    // with the larger reducer count (fewer workers): what fits then fits with four reducers as well
    const size_t b = worker_smem_bytes(n, na, sm_count(), S3_NS::NRED_MAX, nw, g, cols);
    return b > 0 && b <= 200 * 1024;
}

#endif

int S3_LAUNCH(double *gamma, double *S, double *U, double *W, int n, int a0, int na, const int32_t *pairs, int P,
                  const int32_t *inactive, const qio_b200_theta_tables *h_tables, qio_b200_ls_result *results,
                  double *summary, void *workspace, const qio_b200_peers *h_peers, const qio_b200_sweep_options &opts,
                  cudaStream_t st) {
    using namespace S3_NS;
    const int grid = sm_count();
    Params p;
    p.nred = (h_peers && h_peers->nranks > 1) ? NRED_MAX : 4;
    const size_t wsm = worker_smem_bytes(n, na, grid, p.nred, p.nworkers, p.granule, p.wc_cols);
    if (!(wsm > 0 && wsm <= 200 * 1024)) {
        set_error("sweep_cycle: problem does not fit the pipelined kernel");
        return QIO_B200_ERR_UNSUPPORTED;
    }
    const Layout L = ws_layout(grid, n);
    QIO_REQUIRE((long long)P <= (long long)n * (n - 1) / 2 + 1, "sweep_cycle: more pairs than n (n - 1) / 2 (pipelined kernel)");
    QIO_REQUIRE((long long)P + 1 < (1LL << LL_STEP_BITS), "sweep_cycle: too many pairs for the tag width");
    char *ws = reinterpret_cast<char *>(workspace);
    // every signalling word of the workspace (angle ring, abort / started flags, totals, partials) starts at zero: tags
    // are step indices + 1 and need no launch counter (the library keeps no state between calls)
    QIO_CUDA(cudaMemsetAsync(ws, 0, L.off_dtab, st));
    QIO_CUDA(cudaMemsetAsync(summary, 0, 32 * sizeof(double), st));
    p.S = S;
    p.gamma = gamma;
    p.U = U;
    p.W = W;
    p.n = n;
    p.a0 = a0;
    p.na = na;
    p.pairs = pairs;
    p.P = P;
    p.inactive = inactive;
    p.tab = to_dev(*h_tables);
    p.results = results;
    p.summary = summary;
    p.ring = reinterpret_cast<unsigned long long *>(ws + L.off_ring);
    p.abort_flag = reinterpret_cast<unsigned *>(ws + L.off_abort);
    p.tot = reinterpret_cast<unsigned long long *>(ws + L.off_tot);
    p.pbuf = reinterpret_cast<unsigned long long *>(ws + L.off_pbuf);
    p.timeout_ns = 1000000ull * (opts.timeout_ms ? opts.timeout_ms : 4000u);
    p.startup_timeout_ns = 1000000ull * (opts.startup_timeout_ms ? opts.startup_timeout_ms : 60000u);
#ifndef S3_FLAT_UNR
#define S3_FLAT_UNR 5
#endif
    p.flat_unr = S3_FLAT_UNR;      // butterflies in flight per thread of the flat group (3, 4 or 5)
#ifndef S3_TILE_THREADS_SMALL
#define S3_TILE_THREADS_SMALL 256
#endif
    // bytes one rotation reads on this rank: a large share -> the flat phase is the longer one -> more flat threads
#ifndef S3_TILE_THREADS_LARGE
#define S3_TILE_THREADS_LARGE 224
#endif
    p.tile_threads = (32.0 * (double)n * n * na < 20e6) ? S3_TILE_THREADS_SMALL : S3_TILE_THREADS_LARGE;
    // staged flat phase: 2 batches x B butterflies x 2 rows x 16 bytes (8 for odd n) per flat thread, on top of the catch-up
    // staging area (never live at the same time); the largest batch that fits into the 227 KB of an SM
#ifndef S3_FLAT_BATCH_MAX
#define S3_FLAT_BATCH_MAX 4
#endif
    size_t wsm_staged = wsm;
    p.flat_batch = 0;
    {
        const size_t cu_bytes = sizeof(double) * (size_t)n * CU_PITCH, slot = (n & 1) ? 8 : 16;
        const int batches[3] = {4, 3, 2};
        for (int b : batches) {
            const size_t stage = 2 * (size_t)b * 2 * (size_t)(WORK_THREADS - p.tile_threads) * slot;
            const size_t total = wsm - cu_bytes + std::max(cu_bytes, stage);
            if (b <= S3_FLAT_BATCH_MAX && total <= 227 * 1024) {
                p.flat_batch = b;
                wsm_staged = total;
                break;
            }
        }
    }
    p.blocks_out = opts.blocks_out;
    p.trace = opts.trace_out;
    p.trace_k0 = opts.trace_first_pair;
    if (p.trace) QIO_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(double) * TRACE_PAIRS * TRACE_EVENTS, st));
    p.dtab = reinterpret_cast<const int32_t *>(ws + L.off_dtab);
    p.wg = reinterpret_cast<double *>(ws + L.off_wg);
    p.rg = reinterpret_cast<double *>(ws + L.off_rg);
    QIO_REQUIRE(L.off_wg + sizeof(double) * (size_t)p.nworkers * n * p.wc_cols <= L.off_rg, "sweep_cycle: workspace too small for the private W columns");
    p.rank = h_peers ? h_peers->rank : 0;
    p.nranks = h_peers ? h_peers->nranks : 1;
    p.epoch = h_peers ? h_peers->epoch : 0u;
    QIO_REQUIRE(p.nranks == 1 || (p.epoch >= 1 && p.epoch < (1u << (32 - LL_STEP_BITS))), "sweep_cycle: mailbox epoch must be in 1..4095");
    p.mailboxes = p.nranks > 1 ? reinterpret_cast<double *const *>(h_peers->mailboxes) : nullptr;
    size_t dyn = std::max(wsm_staged, sizeof(SearchSmem));
    dyn = std::max(dyn, sizeof(double) * THREADS + 256);
    QIO_CUDA(cudaFuncSetAttribute(sweep3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int blocks_per_sm = 0;
    QIO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, sweep3_kernel, THREADS, dyn));
    QIO_REQUIRE(blocks_per_sm >= 1, "sweep_cycle: pipelined kernel does not fit on an SM");
    if (P > 0) {
        build_dtab_kernel<<<(P + 127) / 128, 128, 0, st>>>(pairs, inactive, P, reinterpret_cast<int32_t *>(ws + L.off_dtab));
        QIO_LAUNCH_CHECK();
        void *args[] = {&p};
        QIO_CUDA(cudaLaunchCooperativeKernel((const void *)sweep3_kernel, dim3(grid), dim3(THREADS), args, dyn, st));
    }
    return QIO_B200_OK;
}

}  // namespace qio
