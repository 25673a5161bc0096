// vg_peer.cuh -- sum of a small block of doubles across the GPUs of one NVLink / NVSwitch domain, done by the
// kernel that produced the block (no collective launch, no host in the loop).
//
// Every rank owns an inbox in its own HBM, mapped into the other ranks' address spaces (CUDA IPC):
//   inbox = slot[2][PEER_MAX_RANKS][2 * PEER_SLOT_DOUBLES] 64-bit words
// One CTA per rank runs peer_allreduce for exchange number `epoch` (the same number on every rank).  The protocol
// is the low-latency one: no fences, no separate flag.  A double travels as two 64-bit words
//   (low 32 bits | tag << 32), (high 32 bits | tag << 32),   tag = epoch mod 2^32,
// each written by one 8-byte store, which is atomic, so a word either still carries an old tag or is complete:
//   1. the CTA stores its block into slot[epoch & 1][my rank] of EVERY rank's inbox (peer stores over NVLink);
//   2. thread i polls word pair i of every rank's slot in its OWN inbox until the tags match, and adds the values
//      up in rank order -- every rank forms the same sum in the same order, bit-identical across ranks.
// Two slot parities are enough: a rank posts exchange e + 2 only after it has collected e + 1, which needed every
// peer's words of e + 1, and a peer posts those only after it has collected exchange e.  The two steps need not
// sit in the same kernel: a launch may post exchange e in its tail and leave the collection to the head of the
// next launch (or to a one-block kernel), so that the NVLink round trip hides behind the next launch's work --
// the rule above is all that has to hold (collect e - 1 before posting e).
#pragma once
#include <cuda_runtime.h>

namespace vg {

constexpr int PEER_SLOT_DOUBLES = 4096;
constexpr int PEER_MAX_RANKS = 16;

struct PeerCtx {
    unsigned long long *const *inbox;   // device array: every rank's inbox as mapped here (own entry: the local pointer)
    int rank, n;                        // n <= 1: no exchange
    unsigned long long epoch;           // starts at 1 (a zeroed inbox carries tag 0)
};

__host__ __device__ inline size_t peer_inbox_bytes()
{
    return sizeof(unsigned long long) * 2 * PEER_MAX_RANKS * 2 * PEER_SLOT_DOUBLES;
}

__device__ __forceinline__ unsigned long long *peer_slot(unsigned long long *inbox, int parity, int r)
{
    return inbox + ((size_t)parity * PEER_MAX_RANKS + r) * 2 * PEER_SLOT_DOUBLES;
}

// A poll gives up after PEER_SPIN_LIMIT reads (ten seconds or more of waiting: a peer that never arrives -- a rank that died or
// issued a different sequence of exchanges) and poisons the sum with NaN instead of hanging the GPU.
constexpr long long PEER_SPIN_LIMIT = 1ll << 25;

// step 1: this rank's block into every rank's inbox.  Called by every thread of ONE CTA, count <= PEER_SLOT_DOUBLES;
// buf may have been written by this CTA just before (the barrier orders it).
__device__ __forceinline__ void peer_post(const double *buf, int count, const PeerCtx &pc)
{
    const int tid = threadIdx.x, nt = blockDim.x, parity = (int)(pc.epoch & 1ull);
    const unsigned long long tag = (pc.epoch & 0xffffffffull) << 32;
    __syncthreads();
    for (int i = tid; i < count; i += nt) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(buf[i]);
        const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
        for (int r = 0; r < pc.n; r++) {
            unsigned long long *dst = peer_slot(pc.inbox[r], parity, pc.rank) + 2 * i;
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(dst), "l"(w0) : "memory");
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(dst + 1), "l"(w1) : "memory");
        }
    }
}

__device__ __forceinline__ unsigned long long peer_poll(const unsigned long long *src, unsigned long long tag)
{
    unsigned long long a = 0;
    for (long long it = 0; it < PEER_SPIN_LIMIT; it++) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
        if ((a & 0xffffffff00000000ull) == tag) return a & 0xffffffffull;
    }
    return ~0ull;           // gave up
}

// step 2: buf[0..count) <- the sum over the ranks' blocks of exchange pc.epoch, in rank order
__device__ __forceinline__ void peer_collect(double *buf, int count, const PeerCtx &pc)
{
    const int tid = threadIdx.x, nt = blockDim.x, parity = (int)(pc.epoch & 1ull);
    const unsigned long long tag = (pc.epoch & 0xffffffffull) << 32;
    unsigned long long *mine = pc.inbox[pc.rank];
    for (int i = tid; i < count; i += nt) {
        double s = 0.0;
        for (int r = 0; r < pc.n; r++) {
            const unsigned long long *src = peer_slot(mine, parity, r) + 2 * i;
            const unsigned long long a = peer_poll(src, tag), b = peer_poll(src + 1, tag);
            s += (a == ~0ull || b == ~0ull) ? __longlong_as_double(0x7ff8000000000000ll)
                                            : __longlong_as_double((long long)(a | (b << 32)));
        }
        buf[i] = s;
    }
}

__device__ __forceinline__ void peer_allreduce(double *buf, int count, const PeerCtx &pc)
{
    peer_post(buf, count, pc);
    peer_collect(buf, count, pc);
}

}  // namespace vg
