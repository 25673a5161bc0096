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
// Two slot parities are enough: a rank starts exchange e + 2 only after it has finished e + 1, which needed every
// peer's words of e + 1, and a peer writes those only after it has consumed exchange e.
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

// buf[0..count) <- sum over ranks; called by every thread of ONE CTA, count <= PEER_SLOT_DOUBLES.
// buf may have been written by this CTA just before (the barrier below orders it).
__device__ __forceinline__ void peer_allreduce(double *buf, int count, const PeerCtx &pc)
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
    unsigned long long *mine = pc.inbox[pc.rank];
    for (int i = tid; i < count; i += nt) {
        double s = 0.0;
        for (int r = 0; r < pc.n; r++) {
            const unsigned long long *src = peer_slot(mine, parity, r) + 2 * i;
            unsigned long long a, b;
            do {
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(src) : "memory");
            } while ((a & 0xffffffff00000000ull) != tag);
            do {
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(src + 1) : "memory");
            } while ((b & 0xffffffff00000000ull) != tag);
            s += __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
        }
        buf[i] = s;
    }
}

}  // namespace vg
