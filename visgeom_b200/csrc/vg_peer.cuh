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
//   2. the CTA polls the word pairs of every rank's slot in its OWN inbox until the tags match (all pairs in flight
//      together), and adds the values up in rank order -- every rank forms the same sum in the same order,
//      bit-identical across ranks.
// Two slot parities are enough: a rank posts exchange e + 2 only after it has collected e + 1, which needed every
// peer's words of e + 1, and a peer posts those only after it has collected exchange e.  The two steps need not
// sit in the same kernel: a launch may post exchange e in its tail and leave the collection to the head of the
// next launch (or to a one-block kernel), so that the NVLink round trip hides behind the next launch's work --
// the rule above is all that has to hold (collect e - 1 before posting e).  With one process per rank (all ranks
// issue the same sequence of calls) even the posting moves to the head of the next launch: a grid cannot end before
// its remote stores are acknowledged, so a tail that posts pays one NVLink round trip per step (measured: 1.1 us
// of a 33 us step on two GPUs).
#pragma once
#include <cuda_runtime.h>

namespace vg {

constexpr int PEER_SLOT_DOUBLES = 4096;
constexpr int PEER_MAX_RANKS = 16;

struct PeerCtx {
    unsigned long long *const *inbox;   // device array: every rank's inbox as mapped here (own entry: the local pointer)
    int rank, n;                        // n <= 1: no exchange
    unsigned long long epoch;           // starts at 1 (a zeroed inbox carries tag 0)
    unsigned long long *fail;           // nullable: host-mapped word; a poll that gave up leaves its exchange number here
    long long spin_limit;               // polls before a collect gives up (<= 0: PEER_SPIN_LIMIT)
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
// issued a different sequence of exchanges): the sum is poisoned with NaN instead of hanging the GPU, and the
// exchange number goes to PeerCtx::fail, which the host checks after its next synchronisation (VG_ERR_PEER).
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

// step 2: buf[0..count) <- the sum over the ranks' blocks of exchange pc.epoch, in rank order.  Called by every
// thread of ONE CTA.  The (element, rank) pairs are spread over the threads -- each polls its pair's two words with
// one 16-byte load, so the round trips to the slots overlap instead of queueing behind one another -- and leave the
// values in `scratch` (cap doubles of shared memory); element i's thread then adds its row up in rank order: every
// rank forms the same sum in the same order.  A pair whose words never arrive poisons the sum with NaN and reports
// the exchange number through pc.fail (the host turns it into an error code at its next fetch).
__device__ __forceinline__ void peer_collect(double *buf, int count, const PeerCtx &pc, double *scratch, int cap)
{
    const int tid = threadIdx.x, nt = blockDim.x, parity = (int)(pc.epoch & 1ull);
    const unsigned long long tag = (pc.epoch & 0xffffffffull) << 32;
    const long long limit = pc.spin_limit > 0 ? pc.spin_limit : PEER_SPIN_LIMIT;
    unsigned long long *mine = pc.inbox[pc.rank];
    const int chunk = cap / pc.n > 0 ? cap / pc.n : 1;
    for (int c0 = 0; c0 < count; c0 += chunk) {
        const int nc = count - c0 < chunk ? count - c0 : chunk;
        __syncthreads();                                   // scratch is free
        for (int q = tid; q < nc * pc.n; q += nt) {
            const int i = c0 + q / pc.n, r = q % pc.n;
            const unsigned long long *src = peer_slot(mine, parity, r) + 2 * i;
            unsigned long long a = 0, b = 0;
            bool ok = false;
            for (long long it = 0; it < limit && !ok; it++) {
                asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
                ok = (a & 0xffffffff00000000ull) == tag && (b & 0xffffffff00000000ull) == tag;
            }
            if (!ok && pc.fail) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(pc.fail), "l"(pc.epoch) : "memory");
            scratch[q] = ok ? __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)))
                            : __longlong_as_double(0x7ff8000000000000ll);
        }
        __syncthreads();
        for (int i = tid; i < nc; i += nt) {
            double s = 0.0;
            for (int r = 0; r < pc.n; r++) s += scratch[i * pc.n + r];
            buf[c0 + i] = s;
        }
    }
}

__device__ __forceinline__ void peer_allreduce(double *buf, int count, const PeerCtx &pc, double *scratch, int cap)
{
    peer_post(buf, count, pc);
    peer_collect(buf, count, pc, scratch, cap);
}

}  // namespace vg
