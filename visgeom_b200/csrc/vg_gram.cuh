// vg_gram.cuh -- per-image normal-equation block [J r]^T [J r] on the FP64 MMA path (mma.sync.m8n8k4.f64,
// SASS DMMA), taken by row parity so that the structural zeros of the intrinsic Jacobian rows are never
// multiplied.  Used by the fused reprojection kernel (vg_eval_impl.cuh); replaces the J^T J / J^T r
// accumulation Ceres performs on the blocks GenericProjectionJac::Evaluate returns
// (src/calibration/calib_cost_functions.cpp:93-114).
//
// Not a stand-alone header: vg_eval_impl.cuh includes it after Layout<>, pk(), smem_u32() and dmma_8x8x4().
#pragma once

namespace vg {

namespace {

// ---- Gram matrix by row parity --------------------------------------------------------------
// Every camera model's intrinsic rows are  u: [d_0..d_{KD-1} | f 0 | 1 0]   v: [d.. | 0 f | 0 1]
// (K = KD + 4; zero rows on a failed projection), so a u-row and a v-row have only
// N = KD + 3 + 6L non-zero columns each:  d(KD), f, one, chain elements (6 each), residual.
// The rows are therefore taken four u-rows or four v-rows at a time (k-steps of one parity), each
// parity with its own accumulators, and the 8-column blocks are cut from those N columns instead
// of the W = K + 6L + 1 dense ones: EUCM mono 2 MMA tiles per k-step instead of 3, MEI 3 instead
// of 6.  When the last block has at most four columns (T), the three tiles (B,B), (B,T), (T,T) of it
// and the previous block B become two MMAs on mixed fragments:
//     am * am   with am = [B_0..B_3 | T]     ->  B_lo x B_lo,  B_lo x T,  T x T
//     xB * bm   with bm = [B_4..B_7 | T]     ->  B x B_hi,  B_hi x T
template <int MODEL, int L> struct ParityGram {
    using LY = Layout<MODEL, L>;
    static constexpr int K = LY::K, KD = K - 4, D = LY::D, W = LY::W;
    static constexpr int N = KD + 3 + 6 * L;
    static constexpr int NB = (N + 7) / 8;
    static constexpr int T = N - 8 * (NB - 1);
    static constexpr bool MERGE = (NB >= 2 && T <= 4);
    // fragments loaded per k-step: the whole blocks, plus (MERGE) the two mixed ones
    //   am  = [B_0..B_3 | T]   and   bm = [B_4..B_7 | T]      (B = last whole block)
    static constexpr int NBX = MERGE ? NB - 1 : NB;                 // whole blocks
    static constexpr int NLD = MERGE ? NBX + 2 : NBX;
    static constexpr int NREG = MERGE ? (NBX - 1) * NBX / 2 + (NBX - 1) : NBX * (NBX + 1) / 2;   // tiles of whole blocks, without (B,B)
    static constexpr int NCROSS = MERGE ? NBX - 1 : 0;              // (block, T) tiles: x[b] * bm
    static constexpr int NTILE = NREG + NCROSS + (MERGE ? 2 : 0);   // + am*am and x[B]*bm
    // full column of effective column e in a row of parity par
    __host__ __device__ static constexpr int full_col(int e, int par)
    {
        return e < KD ? e : e == KD ? KD + par : e == KD + 1 ? KD + 2 + par : e < N - 1 ? K + (e - KD - 2) : D;
    }
};

template <int MODEL, int L> struct GramFrag {
    double v[2][ParityGram<MODEL, L>::NTILE][2];      // [row parity][tile][fragment element]
};

// per-lane output map of the fragments: packed indices of the u-parity and v-parity value of every
// fragment element (-1: not stored; equal: the entry is the sum of both parities)
template <int MODEL, int L> struct GramMap {
    int iu[ParityGram<MODEL, L>::NTILE][2], iv[ParityGram<MODEL, L>::NTILE][2];
};

template <int MODEL, int L>
__device__ __forceinline__ void gram_map_init(GramMap<MODEL, L> &m, const int lane)
{
    using PG = ParityGram<MODEL, L>;
    constexpr int N = PG::N, NB = PG::NB, T = PG::T, W = PG::W, NBX = PG::NBX;
    const int kr = lane & 3, ci = lane >> 2;
    auto set = [&](int t, int q, bool valid, int ei, int ej) {
        int iu = -1, iv = -1;
        if (valid) {
            if (ei > ej) { const int x = ei; ei = ej; ej = x; }
            iu = pk(PG::full_col(ei, 0), PG::full_col(ej, 0), W);
            iv = pk(PG::full_col(ei, 1), PG::full_col(ej, 1), W);
        }
        m.iu[t][q] = iu; m.iv[t][q] = iv;
    };
    constexpr int NW = PG::MERGE ? NBX - 1 : NBX;       // whole blocks that form ordinary tiles among themselves
    int t = 0;
#pragma unroll
    for (int bi = 0; bi < NW; bi++)
#pragma unroll
        for (int bj = bi; bj < NBX; bj++) {               // MERGE: up to and including (bi, B)
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int ei = 8 * bi + ci, ej = 8 * bj + 2 * kr + q;
                set(t, q, ei <= ej && ej < N, ei, ej);
            }
            t++;
        }
    if (PG::MERGE) {
        const int Bc = 8 * (NB - 2), Tc = 8 * (NB - 1);
#pragma unroll
        for (int bi = 0; bi < PG::NCROSS; bi++) {         // x[bi] * bm: only the T half is new
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int j = 2 * kr + q;
                set(t, q, j >= 4 && j - 4 < T, 8 * bi + ci, Tc + j - 4);
            }
            t++;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {                     // am * am
            const int i = ci, j = 2 * kr + q;
            const int ei = i < 4 ? Bc + i : Tc + i - 4, ej = j < 4 ? Bc + j : Tc + j - 4;
            const bool ok = (i < 4 || i - 4 < T) && (j < 4 || j - 4 < T) && i <= j;
            set(t, q, ok, ei, ej);
        }
        t++;
#pragma unroll
        for (int q = 0; q < 2; q++) {                     // xB * bm
            const int i = ci, j = 2 * kr + q;
            if (j < 4) set(t, q, i <= 4 + j, Bc + i, Bc + 4 + j);
            else set(t, q, i >= 4 && j - 4 < T, Bc + i, Tc + j - 4);
        }
    }
}

// write the fragments of one image into a packed upper-triangular block h; map: GramMap as (iu, iv) per fragment
// element and lane (char2 or short2; -1: not stored)
template <int MODEL, int L, typename MAP_T>
__device__ __forceinline__ void gram_frag_emit(const GramFrag<MODEL, L> &f, const MAP_T *map, double *h,
                                               const int lane)
{
    using PG = ParityGram<MODEL, L>;
#pragma unroll
    for (int t = 0; t < PG::NTILE; t++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            // parity-independent entries (iu == iv) receive the sum from both stores
            const MAP_T m = map[(t * 2 + q) * 32 + lane];
            const int iu = m.x, iv = m.y;
            const bool same = iu == iv;
            const double su = same ? f.v[0][t][q] + f.v[1][t][q] : f.v[0][t][q];
            const double sv = same ? su : f.v[1][t][q];
            if (iu >= 0) { h[iu] = su; h[iv] = sv; }
        }
    // structural zeros: products of a u-only and a v-only intrinsic column
    if (lane < 4) {
        constexpr int KD = PG::KD, W = PG::W;
        const int a = lane < 2 ? KD : (lane == 2 ? KD + 1 : KD + 2);
        const int b = lane == 0 ? KD + 1 : (lane == 1 ? KD + 3 : (lane == 2 ? KD + 2 : KD + 3));
        h[pk(a, b, W)] = 0.0;
    }
}

// Gram matrix of one staged image by row parity; overwrites `f`.
template <int MODEL, int L>
__device__ __forceinline__ void gram_slot_parity(const double *rs, const double *Jas, const double *const (&Jes)[L],
                                                 const double *zero, const int lane, const int P,
                                                 GramFrag<MODEL, L> &f)
{
    using PG = ParityGram<MODEL, L>;
    constexpr int K = PG::K, KD = PG::KD, N = PG::N, NB = PG::NB, T = PG::T, NBX = PG::NBX, NTILE = PG::NTILE;
    constexpr int NLD = PG::NLD;
    const int kr = lane & 3, ci = lane >> 2;
    // source of effective column e in this lane's u-row of k-step 0 (corner kr); dv: doubles to the same column
    // of the v-row; st: doubles per k-step (four corners).  Missing columns read a zero slot with stride 0.
    const double *pu[NLD];
    int dv[NLD], st[NLD];
    auto src = [&](const int slot_i, const int e, const bool exists) {
        const double *a;
        int d, s;
        if (!exists || e >= N) { a = zero; d = 0; s = 0; }
        else if (e < KD) { a = Jas + (size_t)2 * kr * K + e; d = K; s = 8 * K; }
        else if (e == KD) { a = Jas + (size_t)2 * kr * K + KD; d = K + 1; s = 8 * K; }
        else if (e == KD + 1) { a = Jas + (size_t)2 * kr * K + KD + 2; d = K + 1; s = 8 * K; }
        else if (e < N - 1) {
            const int el = (e - KD - 2) / 6, q = (e - KD - 2) - 6 * el;
            const double *base = Jes[0];
#pragma unroll
            for (int ee = 1; ee < L; ee++) if (el == ee) base = Jes[ee];
            a = base + (size_t)2 * kr * 6 + q; d = 6; s = 48;
        }
        else { a = rs + 2 * kr; d = 1; s = 8; }
        pu[slot_i] = a; dv[slot_i] = d; st[slot_i] = s;
    };
#pragma unroll
    for (int b = 0; b < NBX; b++) src(b, 8 * b + ci, true);
    if (PG::MERGE) {
        const int Bc = 8 * (NB - 2), Tc = 8 * (NB - 1);
        src(NBX, ci < 4 ? Bc + ci : Tc + ci - 4, ci < 4 || ci - 4 < T);            // am
        src(NBX + 1, ci < 4 ? Bc + 4 + ci : Tc + ci - 4, ci < 4 || ci - 4 < T);    // bm
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int t = 0; t < NTILE; t++) { f.v[h][t][0] = 0.0; f.v[h][t][1] = 0.0; }
    // k-step s: corners 4s..4s+3, their u-rows then their v-rows (two independent accumulator sets)
    auto kstep = [&](const bool guard) {
        double x[2][NLD];
#pragma unroll
        for (int b = 0; b < NLD; b++) {
            x[0][b] = guard ? 0.0 : pu[b][0];
            x[1][b] = guard ? 0.0 : pu[b][dv[b]];
            pu[b] += st[b];
        }
#pragma unroll
        for (int par = 0; par < 2; par++) {
            constexpr int NW = PG::MERGE ? NBX - 1 : NBX;
            int t = 0;
#pragma unroll
            for (int bi = 0; bi < NW; bi++)
#pragma unroll
                for (int bj = bi; bj < NBX; bj++) { dmma_8x8x4(f.v[par][t][0], f.v[par][t][1], x[par][bi], x[par][bj]); t++; }
            if (PG::MERGE) {
#pragma unroll
                for (int bi = 0; bi < PG::NCROSS; bi++) { dmma_8x8x4(f.v[par][t][0], f.v[par][t][1], x[par][bi], x[par][NBX + 1]); t++; }
                dmma_8x8x4(f.v[par][t][0], f.v[par][t][1], x[par][NBX], x[par][NBX]); t++;
                dmma_8x8x4(f.v[par][t][0], f.v[par][t][1], x[par][NBX - 1], x[par][NBX + 1]);
            }
        }
    };
    const int nfull = P >> 2;
    int s = 0;
    for (; s + 1 < nfull; s += 2) { kstep(false); kstep(false); }
    if (s < nfull) kstep(false);
    if (P & 3) kstep(kr >= (P & 3));     // ragged tail: corners beyond P contribute zeros
}

}  // namespace

}  // namespace vg
