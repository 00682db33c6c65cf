// Self-test of the tagged-word signalling (ll.cuh) over peer-mapped memory: every rank runs ONE CTA of `lanes` threads; in
// iteration `it` lane l stores the word (payload(it, rank, l), tag(epoch, it)) into slot [it & 7][rank][l] of EVERY rank's
// mailbox (NVLink stores for the peers), then polls its own mailbox until the words of iteration `it` of all ranks have
// arrived and checks every payload bit.  A torn 16-byte observation that slipped through the two-tag check would show up
// as a payload mismatch; a lost word as a time-out.  Both 32-bit halves of the payload change with every iteration, so a
// stale half can never equal the new one.  Flow control is implicit: a rank posts iteration it + 1 only after it has read
// iteration `it` of every rank, so a slot (reused every 8 iterations) is never overwritten while someone still reads it.
#include "common.cuh"
#include "ll.cuh"

namespace qio {

constexpr int LLS_SLOTS = 8;
constexpr int LLS_MAX_LANES = 256;

__device__ __forceinline__ unsigned long long lls_payload(unsigned it, int rank, int lane) {
    const unsigned lo = it * 2654435761u + (unsigned)lane * 40503u + (unsigned)rank;
    const unsigned hi = ~(it * 2246822519u) ^ ((unsigned)rank << 24) ^ (unsigned)lane;
    return ((unsigned long long)hi << 32) | lo;
}

__global__ void __launch_bounds__(LLS_MAX_LANES)
    ll_stress_kernel(double *const *mailboxes, int rank, int nranks, unsigned epoch, int iterations, int lanes,
                     unsigned long long *result /* [0] mismatches, [1] time-outs, [2] words checked */) {
    const int lane = threadIdx.x;
    if (lane >= lanes) return;
    unsigned long long bad = 0, lost = 0, seen = 0;
    for (int it = 0; it < iterations; ++it) {
        const unsigned tag = mailbox_tag(epoch, it);
        const size_t slot = (size_t)(it & (LLS_SLOTS - 1)) * nranks;
        const unsigned long long mine = lls_payload((unsigned)it, rank, lane);
        for (int dst = 0; dst < nranks; ++dst)
            ll_store(reinterpret_cast<unsigned long long *>(mailboxes[dst]) + ((slot + rank) * LLS_MAX_LANES + lane) * 2, mine, tag);
        for (int r = 0; r < nranks; ++r) {
            const unsigned long long *w = reinterpret_cast<const unsigned long long *>(mailboxes[rank]) + ((slot + r) * LLS_MAX_LANES + lane) * 2;
            unsigned long long v = 0;
            bool ok = false;
            for (long long spin = 0; spin < (1LL << 24) && !ok; ++spin) ok = ll_load(w, v, tag);
            if (!ok) {
                ++lost;
                continue;
            }
            ++seen;
            if (v != lls_payload((unsigned)it, r, lane)) ++bad;
        }
        if (lost) break;        // the other ranks will time out as well
    }
    atomicAdd(result, bad);
    atomicAdd(result + 1, lost);
    atomicAdd(result + 2, seen);
}

}  // namespace qio

extern "C" size_t qio_b200_ll_stress_bytes(int nranks) {
    return sizeof(unsigned long long) * 2 * (size_t)qio::LLS_SLOTS * (size_t)(nranks > 0 ? nranks : 1) * qio::LLS_MAX_LANES;
}

extern "C" int qio_b200_ll_stress(const qio_b200_peers *h_peers, int iterations, int lanes, unsigned long long *result,
                                  void *stream) {
    using namespace qio;
    QIO_REQUIRE(h_peers && h_peers->mailboxes && h_peers->nranks >= 1 && h_peers->rank >= 0 && h_peers->rank < h_peers->nranks,
                "ll_stress: bad peers");
    QIO_REQUIRE(iterations >= 0 && iterations + 1 < (1 << LL_STEP_BITS) && lanes >= 1 && lanes <= LLS_MAX_LANES && result,
                "ll_stress: bad arguments");
    QIO_REQUIRE(h_peers->epoch >= 1 && h_peers->epoch < (1u << (32 - LL_STEP_BITS)), "ll_stress: mailbox epoch must be in 1..4095");
    QIO_CUDA(cudaMemsetAsync(result, 0, 3 * sizeof(unsigned long long), as_stream(stream)));
    ll_stress_kernel<<<1, LLS_MAX_LANES, 0, as_stream(stream)>>>(reinterpret_cast<double *const *>(h_peers->mailboxes), h_peers->rank,
                                                                  h_peers->nranks, h_peers->epoch, iterations, lanes, result);
    QIO_LAUNCH_CHECK();
    return QIO_B200_OK;
}
