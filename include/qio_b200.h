/* qio_b200.h -- C ABI of the B200-native QIO orbital-entropy hot path.
 *
 * The reference (schilling-group/QIO) is pure Python and has NO foreign-function interface;
 * each entry point below replaces the numpy/scipy body of one reference function (cited as
 * file:line of the reference tree) and is what a ctypes binding inside the reference's own
 * modules would call (INTEGRATION.md shows those stubs).
 *
 * Conventions
 *   - every array argument is a DEVICE pointer owned by the caller unless its name starts
 *     with h_ (host); float64 and int32 only; C (row-major) order.
 *   - n            number of spatial orbitals.
 *   - gamma        (2n, 2n) spin-orbital 1-RDM, spin interleaved (even = up, odd = down).
 *   - Gamma        (n, n, n, n) 2-RDM block <a_up^+ b_dn^+ c_dn d_up>.  A rank may hold only the
 *                  slabs Gamma[a0 : a0+na, :, :, :]  ("sharded by leading orbital index"); the
 *                  pointer then addresses the first local slab.  na == n, a0 == 0 is the full tensor.
 *   - inactive     int32 mask of length n (1 = orbital counts in the cost); the reference's
 *                  Python-list membership test `k in inactive_indices`.
 *   - pairs        int32 (P, 2) array of (i, j).
 *   - stream       a cudaStream_t (passed as void*); all work is enqueued asynchronously on it.
 *   - return value 0 on success, negative on error; qio_b200_last_error() gives the message
 *     (thread local).  Nothing throws across the ABI.  No CPU fallback exists: without a CUDA
 *     device every compute entry point returns QIO_B200_ERR_CUDA.
 */
#ifndef QIO_B200_H
#define QIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QIO_B200_OK 0
#define QIO_B200_ERR_ARG (-1)
#define QIO_B200_ERR_CUDA (-2)
#define QIO_B200_ERR_UNSUPPORTED (-3)

#define QIO_B200_ABI_VERSION 2

/* pair-block record: 16 Gamma entries Gamma[I_m, I_n, I_p, I_q] (I = (i, j), index m*8+n*4+p*2+q),
 * then gamma_up 2x2 (row-major over (I_m, I_n)), then gamma_dn 2x2. */
#define QIO_B200_BLOCK_DOUBLES 24

/* flag bits reported by the entropy evaluations (reference qio/entropy.py:15-20) */
#define QIO_B200_FLAG_WARN_NEGATIVE 1 /* some spectrum entry < -1e-6  -> warnings.warn   */
#define QIO_B200_FLAG_RAISE_GT1 2     /* no negative entry and one > 1 -> ValueError     */

/* per-pair line-search record written by qio_b200_pair_linesearch / qio_b200_jacobi_cycle */
typedef struct qio_b200_ls_result {
    int32_t coarse_index; /* index into the 314-point coarse grid, -1 = nothing improved   */
    int32_t fine_index;   /* index into the fine grid of that coarse point, -1 = None       */
    int32_t flags;        /* QIO_B200_FLAG_* accumulated over all evaluations of the pair   */
    int32_t accepted;     /* 1 when a rotation angle was found (fine_index >= 0)             */
    double theta;         /* the angle (NaN when not accepted)                               */
    double cost;          /* pair cost at the end of the search                              */
    double cost0;         /* pair cost at theta = 0                                          */
    double coarse_margin; /* min(|cost0 - best|, second best - best) on the coarse grid      */
    double fine_margin;   /* min over the fine scan of |cost - (new_cost + 1e-8)|            */
                          /* (both margins are +inf in the records of the pipelined sweep    */
                          /* kernel: they are off its per-pair critical path and come from   */
                          /* the strict-order audit pass over qio_b200_sweep_options.blocks_out) */
    double cos_theta;     /* cos / sin of the accepted angle (1, 0 when not accepted)        */
    double sin_theta;
    double reserved;
} qio_b200_ls_result;

/* angle tables built ON THE HOST with numpy so that theta, cos(theta), sin(theta) are bit-identical
 * to the reference's np.arange / np.cos / np.sin (qio/grad/jacobi.py:36,139,149). */
typedef struct qio_b200_theta_tables {
    int32_t n_coarse;           /* 314                                                     */
    int32_t fine_stride;        /* row pitch of the fine tables (>= max fine length, 201)   */
    const double *coarse_theta; /* [n_coarse]                                               */
    const double *coarse_cs;    /* [n_coarse][2]  cos, sin                                  */
    const int32_t *fine_len;    /* [n_coarse]     200 or 201                                */
    const double *fine_theta;   /* [n_coarse][fine_stride]                                  */
    const double *fine_cs;      /* [n_coarse][fine_stride][2]                               */
} qio_b200_theta_tables;

const char *qio_b200_last_error(void);
int qio_b200_abi_version(void);
/* sm_count, compute capability of the current device; QIO_B200_ERR_CUDA without a device. */
int qio_b200_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- a-1  qio/qio.py:152-174  prep_rdm12 -------------------------------------------------
 * gamma[2a,2b] = gamma[2a+1,2b+1] = dm1[a,b]/2 (other entries 0);
 * Gamma[a,b,c,d] = (2*dm2[a,d,b,c] + dm2[a,c,b,d]) / 6, bit-exact with the numpy expression.
 * dm2 / Gamma address `na` leading slabs (both permutations keep the leading index, so the
 * operation is shard-local).  dm1/gamma may be NULL to skip the 1-RDM. */
int qio_b200_prep_rdm12(const double *dm1, const double *dm2, double *gamma, double *Gamma,
                        int n, int na, void *stream);
/* Self-test of the three-operation correctly rounded x / 6 the prep kernel uses (csrc/prep.cu div6_rn) against the IEEE
 * division: 2 * count pseudo-random doubles; *bad (device) receives the number of results that differ in any bit. */
int qio_b200_selftest_div6(unsigned long long count, unsigned long long seed, unsigned long long *bad, void *stream);

/* ---- a-2/a-3  qio/entropy.py:4-47  shannon / get_cost_fqi ----------------------------------
 * diag (3, m): nu = gamma[2k,2k], nd = gamma[2k+1,2k+1], nn = Gamma[k,k,k,k] for k = inds[t];
 * nn is 0 where slab k is not local (a0 <= k < a0+na), so shards combine by a sum. */
int qio_b200_orbital_diagonals(const double *gamma, const double *Gamma, int n, int a0, int na,
                               const int32_t *inds, int m, double *diag, void *stream);
/* spec (4, m) = [1-nu-nd+nn, nu-nn, nd-nn, nn]; s_masked (m) = -sum_{p>0} p ln p per orbital;
 * s_unmasked (m) = -sum ln(p) p without the mask (NaN on a zero entry, qio/grad/jacobi.py:318);
 * summary (4) = [total masked entropy, min spec, max spec, flags as double].
 * Any of spec / s_masked / s_unmasked may be NULL. */
int qio_b200_orbital_entropies(const double *diag, int m, double *spec, double *s_masked,
                               double *s_unmasked, double *summary, void *stream);

/* shannon(spec) of a flat array of m probabilities (qio/entropy.py:4-22):
 * summary (4) = [-sum_{p>0} p ln p, min, max, flags as double]. */
int qio_b200_shannon(const double *p, long long m, double *summary, void *stream);

/* ---- pair blocks (gather stage of a-4 / a-5 / a-8 / a-9) -----------------------------------
 * blocks (P, 24) as QIO_B200_BLOCK_DOUBLES describes.  Entries whose leading Gamma index is
 * not local are written as 0; the gamma entries are written only by the owner of slab i
 * (pairs[p][0]) -- shards therefore combine by an element-wise sum with exact results. */
int qio_b200_pair_blocks(const double *gamma, const double *Gamma, int n, int a0, int na,
                         const int32_t *pairs, int P, double *blocks, void *stream);

/* ---- a-8  qio/grad/gradient.py:13-146  FQI_grad / FQI_hess ---------------------------------
 * From the blocks of ALL pairs i > j in the order k = i(i-1)/2 + j: dX (n, n) antisymmetric,
 * ddX (n, n) symmetric, zero diagonals.  The reference's literal formulas are kept
 * (gradient.py:18-21 down derivative copies up; :40 Gamma[i,j,j,j] counted twice). */
int qio_b200_fqi_grad_hess(const double *blocks, int n, const int32_t *inactive, double *dX,
                           double *ddX, void *stream);
/* Explicit pair list in any order (orb_entr_grad / orb_entr_hess, gradient.py:62-120):
 * g1[p], g2[p] = d/dtheta, d2/dtheta2 of the pair cost at 0 for blocks (P, 24) of `pairs`. */
int qio_b200_pair_grad_hess(const double *blocks, const int32_t *pairs, int P,
                            const int32_t *inactive, double *g1, double *g2, void *stream);
/* Fused single-GPU form: gathers from (gamma, Gamma) itself (full tensor only). */
int qio_b200_fqi_grad_hess_fused(const double *gamma, const double *Gamma, int n,
                                 const int32_t *inactive, double *dX, double *ddX, void *stream);

/* ---- a-4  qio/grad/jacobi.py:15-60  jacobi_cost --------------------------------------------
 * C candidates: cand_block[t] indexes `blocks`, cand_pair (C, 2) gives (i, j) for the inactive
 * test, cand_cs (C, 2) = cos(theta), sin(theta) computed by the caller.  cost (C), flags (C). */
int qio_b200_pair_cost(const double *blocks, const int32_t *cand_block, const int32_t *cand_pair,
                       const double *cand_cs, int C, const int32_t *inactive, double *cost,
                       int32_t *flags, void *stream);

/* ---- a-5  qio/grad/jacobi.py:117-158  jacobi_direct, for P independent pairs ---------------
 * Sequential semantics reproduced exactly: strict '>' running minimum over the coarse grid,
 * fallback t_opt = 0.01, fine scan with the 1e-8 acceptance margin.  One CTA per pair, one
 * thread per angle.  blocks (P, 24) in the order of `pairs`.  strict_order != 0 evaluates every
 * candidate in the reference's own floating-point operation order (16-term sums, library log);
 * 0 uses the polynomial form of the rotated occupations and a branch-free log (values agree to a few
 * 1e-16; the returned angle can differ only when two candidates are that close, see the margins). */
int qio_b200_pair_linesearch(const double *blocks, const int32_t *pairs, int P,
                             const int32_t *inactive, const qio_b200_theta_tables *h_tables,
                             qio_b200_ls_result *results, int strict_order, void *stream);

/* ---- a-6  qio/grad/jacobi.py:63-115  jacobi_transform --------------------------------------
 * In-place Givens rotation of orbitals (i, j) by (c, s) = (cos t, sin t) on all four axes of
 * Gamma (only the n^4-(n-2)^4 entries that change are touched) and on the up-up / down-down
 * blocks of gamma.  zero_offspin != 0 additionally zeroes the up-down / down-up entries of
 * gamma, which is what the reference's freshly allocated gamma_ ends up with (jacobi.py:87).
 * Full tensor only (na == n). */
int qio_b200_givens_apply(double *gamma, double *Gamma, int n, int i, int j, double c, double s,
                          int zero_offspin, void *stream);

/* ---- 2-index / 4-index orbital rotation (qio/grad/jacobi.py:200-202, gradient.py:212-214) ---
 * gamma_out = (U x 1_2) gamma_in (U x 1_2)^T, all four spin blocks.  U is (n, n) row-major. */
int qio_b200_rotate_rdm1(const double *U, const double *gamma_in, double *gamma_out, int n,
                         void *stream);
/* Gamma <- sum U[i,a] U[j,b] U[k,c] U[l,d] Gamma[a,b,c,d], in place, as four single-axis FP64
 * tensor-core (DMMA) GEMM passes that cyclically permute the axes.  `work` is n^4 doubles. */
int qio_b200_rotate_rdm2(const double *U, double *Gamma, double *work, int n, void *stream);
/* One pass: out[(b,c,d), i] = sum_a U[i,a] in[a, (b,c,d)]  for `rows` = number of (b,c,d)
 * columns of `in` (row pitch of `in` is `ld_in` doubles); out is (rows, n).  Building block of
 * the sharded rotation (the caller re-shards between passes). */
int qio_b200_rotate_axis(const double *U, const double *in, double *out, int n, long long rows,
                         long long ld_in, void *stream);
/* `batch` such passes in ONE launch: problem b reads in + b * batch_stride_in, writes out + b * batch_stride_out
 * (strides in doubles).  The local slabs of a leading-index shard are rotated on their trailing axes this way
 * (rows = n^2, ld_in = n^2, strides n^3): three launches per shard instead of three per slab. */
int qio_b200_rotate_axis_batched(const double *U, const double *in, double *out, int n, long long rows,
                                 long long ld_in, int batch, long long batch_stride_in, long long batch_stride_out,
                                 void *stream);

/* Rectangular cyclic pass (qio/solver/tccsd.py:157-158,164-172: the amplitude projections
 * 'ijab,Ii,Jj,Aa,Bb->IJAB' between the CAS and the CCSD orbital spaces are four of these):
 *   out[m, i] = sum_{k < kdim} U[i * ldu + k] in[k, m],   in (kdim, rows) -> out (rows, n_out):
 * the leading axis (length kdim) is contracted with the n_out x kdim matrix U and re-appears as the last axis. */
int qio_b200_contract_axis_cyclic(const double *U, long long ldu, const double *in, double *out, int n_out,
                                  int kdim, long long rows, void *stream);

/* Single-axis contraction that keeps the index order (used to flush the deferred sweep state):
 *   last_axis == 0:  out[i, m] = sum_a U[i, a] in[a, m]   in, out are (n, rows)
 *   last_axis != 0:  out[m, i] = sum_d in[m, d] U[i, d]   in, out are (rows, n)              */
int qio_b200_contract_axis(const double *U, const double *in, double *out, int n, long long rows,
                           int last_axis, void *stream);

/* Partial leading-axis contraction over a shard of the contracted index (multi-GPU rotation / flush):
 *   out[i, m] = sum_{k < kdim} U[i * ldu + k] in[k, m],   in (kdim, rows), out (n, rows);
 * the ranks' outputs are summed by a reduce-scatter. */
int qio_b200_contract_axis_partial(const double *U, long long ldu, const double *in, double *out, int n,
                                   int kdim, long long rows, void *stream);

/* ---- a-7 (deferred-axis form)  the Jacobi cycle as one persistent cooperative kernel ----------
 * State: S with axes 1, 2 in the current basis and axes 0, 3 in the basis of the last flush, W (n, n)
 * the rotation accumulated since:  Gamma[a,b,c,d] = sum W[a,a'] W[d,d'] S[a',b,c,d'].  A rank holds the
 * slabs S[a0 : a0+na] of the (deferred) leading axis; gamma, U, W are replicated on every rank.
 * sweep_reset sets W = 1 (S = Gamma then).  sweep_cycle runs the pair sequence exactly like
 * qio_b200_jacobi_cycle (same result records) but touches only 64 n^3 B of S per accepted rotation, all
 * row-contiguous and shard-local, and searches ahead of the pair being applied.  With peers, every
 * rank calls it with the same arguments on its own shard; one small super block (256 doubles in a run of pairs that
 * share an orbital, at most 1296) per pair is exchanged through peer-mapped mailboxes inside the kernel (no NCCL call
 * on the path) and all ranks take identical decisions.
 * summary (32 doubles) = [0] number accepted, [2] OR of flags, [3] != 0: a wait inside the pipelined kernel timed out
 * (cycle incomplete), [24] engine (3 = pipelined, 0 = barrier kernel); the other entries are phase timers in ns
 * (line-search stages in SM cycles), filled by the barrier kernel always and by the pipelined kernel only when
 * options.timers is set (a second build of the kernel that carries the timers).
 * deferred_diagonals is orbital_diagonals through the pending W (per-shard partial sums).  To obtain Gamma
 * again: contract_axis(W, S -> work, leading) then contract_axis(W, work -> S, last), then sweep_reset.
 * The library keeps NO state between calls: everything that selects a code path travels in qio_b200_sweep_options,
 * everything that distinguishes launches (the mailbox epoch) in qio_b200_peers, and the signalling words of the
 * workspace are cleared by every launch. */
typedef struct qio_b200_peers {
    int32_t rank, nranks;
    uint32_t epoch;         /* 1..4095; must differ between consecutive sweep_cycle calls and agree on all ranks; before
                               a value is reused every rank must have zeroed its own mailbox behind a barrier */
    uint32_t reserved;
    void *const *mailboxes; /* DEVICE array [nranks] of mailbox base pointers as mapped into THIS process
                               (own + cudaIpc-opened peers); each mailbox has qio_b200_mailbox_bytes(nranks) */
} qio_b200_peers;
/* Options of one sweep_cycle call (host struct; NULL = all defaults). */
typedef struct qio_b200_sweep_options {
    int32_t engine;              /* 0 = automatic (qio_b200_sweep_engine), 2 = barrier kernel, 3 = pipelined kernel
                                    (QIO_B200_ERR_UNSUPPORTED when the shape does not fit it) */
    int32_t timers;              /* != 0: the build of the pipelined kernel that carries its phase timers (~5 % slower) */
    uint32_t timeout_ms;         /* time-out of every wait inside the pipelined kernel; 0 = 4000 */
    uint32_t startup_timeout_ms; /* the same until the first super block has arrived (launch skew between ranks,
                                    lazy module loading, an attached profiler); 0 = 60000 */
    double *blocks_out;          /* DEVICE (P, QIO_B200_BLOCK_DOUBLES) or NULL: the pair block every line search of the
                                    cycle ran on -- feeding it to qio_b200_pair_linesearch(strict_order = 1) re-derives
                                    every decision of the cycle in the reference's own operation order (decision audit) */
    double *trace_out;           /* DEVICE (32, 16) or NULL, timers build only: globaltimer stamps (ns) of ten events along the
                                    per-pair dependency chain (search start, angle published, worker saw it / started / posted
                                    its partial, reducer got all partials / published, prefetch got the total / finished) for
                                    the 32 pairs starting at trace_first_pair (profiling aid; csrc/sweep3.cu TR_*) */
    int32_t trace_first_pair;
    int32_t reserved;
} qio_b200_sweep_options;
size_t qio_b200_sweep_workspace_bytes(int n);
/* Which kernel sweep_cycle runs by default for (n, na): 3 = pipelined (no grid barriers: column-owning worker CTAs, reducer
 * CTAs and a search CTA that works three rotations ahead of the data it needs; used wherever a worker's columns of W fit in
 * shared memory), 2 = one grid barrier per pair.  Same result records either way. */
int qio_b200_sweep_engine(int n, int na);
size_t qio_b200_mailbox_bytes(int nranks);
int qio_b200_sweep_reset(double *W, int n, void *stream);
int qio_b200_sweep_cycle(double *gamma, double *S, double *U, double *W, int n, int a0, int na,
                         const int32_t *pairs, int P, const int32_t *inactive,
                         const qio_b200_theta_tables *h_tables, qio_b200_ls_result *results, double *summary,
                         void *workspace, const qio_b200_peers *h_peers, const qio_b200_sweep_options *h_opts,
                         void *stream);
int qio_b200_deferred_diagonals(const double *gamma, const double *S, const double *W, int n, int a0, int na,
                                const int32_t *inds, int m, double *diag, void *stream);
/* zeroes the up-down / down-up entries of gamma when summary[0] > 0 (the reference's jacobi_transform
 * returns zeros there, qio/grad/jacobi.py:87). */
int qio_b200_zero_offspin_if_rotated(double *gamma, int n, const double *summary, void *stream);
/* cudaMalloc'ed, zeroed buffer + its 64-byte CUDA IPC handle; open / close a peer's; free one's own; clear zeroes
 * one's own mailbox again (stream-ordered) -- required, between two barriers over the ranks, before an epoch is reused. */
int qio_b200_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle64);
int qio_b200_ipc_clear(void *ptr, size_t bytes, void *stream);
int qio_b200_ipc_open(const unsigned char *handle64, void **ptr);
int qio_b200_ipc_close(void *ptr);
int qio_b200_ipc_free(void *ptr);
/* Self-test of the tagged-word signalling over the peer mailboxes (which must hold at least
 * max(qio_b200_mailbox_bytes, qio_b200_ll_stress_bytes) bytes): `iterations` rounds in which every rank stores `lanes`
 * (<= 256) tagged 16-byte words into every rank's mailbox and verifies every payload bit of the words it receives.
 * All ranks call it together with the same arguments and a fresh epoch.  result (device, 3 x uint64) = [payload
 * mismatches (torn or stale words that passed the tag check), time-outs, words checked]. */
size_t qio_b200_ll_stress_bytes(int nranks);
int qio_b200_ll_stress(const qio_b200_peers *h_peers, int iterations, int lanes, unsigned long long *result,
                       void *stream);

/* ---- a-7  qio/grad/jacobi.py:211-233  one Jacobi cycle over a given pair sequence ------------
 * For each pair in order: line search on the CURRENT RDMs, and when an angle is accepted the
 * Givens update of (gamma, Gamma) and U <- V(t,i,j) U.  Everything stays on the device; the
 * host sees only `results` (P records) afterwards.  summary (4) = [number accepted,
 * cost after the cycle over the inactive orbitals (get_cost_fqi), OR of flags, 0].
 * workspace: qio_b200_jacobi_workspace_bytes(n, P) bytes. */
size_t qio_b200_jacobi_workspace_bytes(int n, int P);
int qio_b200_jacobi_cycle(double *gamma, double *Gamma, double *U, int n, const int32_t *pairs,
                          int P, const int32_t *inactive, const qio_b200_theta_tables *h_tables,
                          qio_b200_ls_result *results, double *summary, void *workspace,
                          void *stream);

/* ---- a-9 (extension, no reference counterpart)  two-orbital RDM / mutual information ----------
 * From the pair blocks: the 16 x 16 two-orbital RDM in its (N, S_z) blocks 1,2,2,1,4,1,2,2,1,
 * closed with the 2-body data only (the nine 3- and 4-body expectation values listed in
 * qio_b200/csrc/mi_formulas.inc are taken from `supp` (P, QIO_B200_MI_SUPP_DOUBLES) when given, else 0 --
 * exact for two-electron states; same-spin 2-RDM entries use the singlet relation Gamma[abcd] - Gamma[abdc]).
 * Block eigenvalues by cyclic Jacobi in registers; out: s_pair (P) and eig (P, 16). */
#define QIO_B200_MI_SUPP_DOUBLES 9
int qio_b200_pair_entropy(const double *blocks, const double *supp, int P, double *s_pair,
                          double *eig, int32_t *flags, void *stream);
/* Mutual-information matrix I (n, n): I[i,j] = I[j,i] = s1[i] + s1[j] - s_pair[p] for pairs[p] = (i, j); zero diagonal.
 * s1 (n) = one-orbital entropies (qio_b200_orbital_entropies, s_masked), s_pair (P) from qio_b200_pair_entropy. */
int qio_b200_mutual_information(const double *s1, const double *s_pair, const int32_t *pairs, int P, int n,
                                double *I, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QIO_B200_H */
