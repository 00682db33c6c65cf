// Tagged 16-byte words ("LL" signalling) shared by the sweep kernels: data and flag travel in ONE store, the
// consumer polls the word itself, no fence and no separate flag.
//
// A word carries 64 data bits and a 32-bit tag as two 8-byte halves, each (32 data bits | tag << 32).  The PTX memory
// model guarantees single-copy atomicity only per aligned 8-byte access: a 16-byte vector store may become visible as
// two independent 8-byte writes in either order (locally through L2 as well as through NVLink into a peer's mailbox).
// With the tag repeated in both halves a torn observation (one new half, one stale half) shows two different tags and
// is rejected -- the reader accepts a word only when both halves carry the tag it waits for.  (This is the layout of
// NCCL's LL protocol; the earlier (value64, tag64) layout relied on 16-byte atomicity.)
//
// Tags are never 0 (words are zero-initialised) and are unique per slot within a launch; buffers that survive a launch
// (the peer mailboxes) carry a launch epoch in the upper tag bits (see mailbox_tag).
#pragma once
#include <stdint.h>

namespace qio {

__device__ __forceinline__ void ll_store(unsigned long long *w, unsigned long long bits, unsigned tag) {
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"((bits & 0xffffffffull) | t), "l"((bits >> 32) | t)
                 : "memory");
}

// Loads the word; true when both halves carry `want` (then `bits` is the value that was stored with it).
__device__ __forceinline__ bool ll_load(const unsigned long long *w, unsigned long long &bits, unsigned want) {
    unsigned long long a, b;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(w) : "memory");
    bits = (a & 0xffffffffull) | (b << 32);
    return (unsigned)(a >> 32) == want && (unsigned)(b >> 32) == want;
}

// Same, returning the two tags (for records whose tag carries payload bits).
__device__ __forceinline__ void ll_load_tags(const unsigned long long *w, unsigned long long &bits, unsigned &t0, unsigned &t1) {
    unsigned long long a, b;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(w) : "memory");
    bits = (a & 0xffffffffull) | (b << 32);
    t0 = (unsigned)(a >> 32);
    t1 = (unsigned)(b >> 32);
}

// Mailbox tag: 12 bits of launch epoch (1..4095, the host clears the mailboxes behind a barrier when it wraps) and
// 20 bits of step index + 1 (the pipelined kernel takes n <= 1024: at most 523,777 pairs).
constexpr int LL_STEP_BITS = 20;
__device__ __forceinline__ unsigned mailbox_tag(unsigned epoch, int step) {
    return (epoch << LL_STEP_BITS) | (unsigned)(step + 1);
}

}  // namespace qio
