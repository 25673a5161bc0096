// vg_eval_impl.cuh -- fused reprojection residual + analytic Jacobian + per-image
// normal-equation kernel for sm_100a (one instantiation per camera model and
// chain length).
//
// Replaces, for every image of a dataset, GenericProjectionJac::Evaluate
// (src/calibration/calib_cost_functions.cpp:28-117) and the J^T J / J^T r build
// Ceres performs on the block it returns.
//
// Persistent CTAs: CTA b owns the image groups b, b + gridDim.x, ... (static, hence
// deterministic); a group is G consecutive images.  Per group:
//   0. pose   : (once per PCG groups) one thread per image accumulates the transform
//               chain with rotation matrices and stages R, t and, per chain element,
//               R12, M12, t13 (InterJacobian's members, jacobian.h:139-152).
//   A. corner : one thread per (image, corner): X = R Xb + t, ONE evaluation of the
//               camera model (projection, dP/dX, dP/dintr share rho, eta and their
//               reciprocals), residual and Jacobian rows written into shared memory in
//               exactly the Ceres block layout.  Observations: coalesced 16-byte loads.
//   S. store  : the staged blocks of the G images are contiguous in global memory, so
//               one elected thread streams each region out with a TMA bulk copy
//               (cp.async.bulk shared::cta -> global, SASS UBLKCP); no register round trip.
//   B. normal : every warp takes one balanced tile of the per-image packed
//               [J r]^T [J r] (structural zeros of the intrinsic rows skipped), re-reads
//               the staged rows, accumulates in registers with the rows split over
//               S = 32/G lanes, and combines the splits through a small shared-memory
//               transpose.  The blocks leave through coalesced stores; their per-CTA
//               sums (for the shared normal-equation block) stay in registers until the
//               CTA ends.
// The kernel is HBM-write bound by design (224 B written per EUCM corner); tensor cores
// are not used -- there is no dense contraction on this path.
#pragma once
#include "vg_eval.cuh"
#include "vg_math.cuh"

#include <cstdint>
#include <type_traits>
#include <vector>

namespace vg {

struct LaunchPlan { int G, threads, PCG; long long smem; };
bool plan_eval(int model, int L, int P, LaunchPlan *pl);

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// TMA bulk copy shared -> global (SASS: UBLKCP), tracked by the bulk async-group
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void bulk_store_chunked(double *gdst, const double *ssrc, size_t bytes)
{
    const size_t CH = 32768;
    char *g = reinterpret_cast<char *>(gdst);
    const char *s = reinterpret_cast<const char *>(ssrc);
    while (bytes) {
        const size_t n = bytes < CH ? bytes : CH;
        bulk_store(g, s, static_cast<uint32_t>(n));
        g += n; s += n; bytes -= n;
    }
}

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__host__ __device__ constexpr int round_up2(int x) { return (x + 1) & ~1; }
// packed upper-triangular index of (a,b), a <= b, in a W x W symmetric matrix
__host__ __device__ constexpr int pk(int a, int b, int W) { return a * W - a * (a - 1) / 2 + (b - a); }

// ---- phase B plan: atoms, balanced tiles --------------------------------------------------
// Every model's intrinsic Jacobian row has the shape (eucm.h:208-222, ucm.h:176-192,
// mei.h:257-283)
//     u-row: [ d_0 .. d_{KD-1} | f 0 | 1 0 ]      v-row: [ d_0 .. d_{KD-1} | 0 f | 0 1 ]
// so a row is described by the operands  d_i, f, c (the 1, or 0 when the projection failed),
// p_{e,j} (chain element e) and r, plus its parity.  The products that make up the packed
// block are grouped into "atoms"; atoms are bin-packed (largest first) into NT tiles of
// roughly equal accumulator count, one tile per warp.
constexpr int MAX_TILES = 48;
constexpr int MAX_ACC = 40;

enum AtomKind { A_DDA, A_DDB, A_DR, A_DF, A_FF, A_FR, A_DPA, A_DPB, A_FP, A_PPA, A_PPB, A_PR, A_PQA, A_PQB };

template <int KD, int L> struct Plan {
    static constexpr int K = KD + 4;
    static constexpr int KH = (KD + 1) / 2;
    static constexpr int NA = 6 + 6 * L + L * (L - 1);
    // symbolic operand codes
    static constexpr int OP_F = KD, OP_C = KD + 1, OP_P = KD + 2, OP_R = KD + 2 + 6 * L, NOPS = OP_R + 1;

    struct AtomDesc { int kind, e, b, n; };
    __host__ __device__ static constexpr AtomDesc atom(int i)
    {
        const int dda = KH * KD - KH * (KH - 1) / 2, ddall = KD * (KD + 1) / 2;
        if (i == 0) return {A_DDA, 0, 0, dda};
        if (i == 1) return {A_DDB, 0, 0, ddall - dda};
        if (i == 2) return {A_DR, 0, 0, KD};
        if (i == 3) return {A_DF, 0, 0, 2 * KD};
        if (i == 4) return {A_FF, 0, 0, 3};
        if (i == 5) return {A_FR, 0, 0, 3};
        i -= 6;
        if (i < 6 * L) {
            const int e = i / 6, q = i % 6;
            if (q == 0) return {A_DPA, e, 0, 6 * KH};
            if (q == 1) return {A_DPB, e, 0, 6 * (KD - KH)};
            if (q == 2) return {A_FP, e, 0, 12};
            if (q == 3) return {A_PPA, e, 0, 11};
            if (q == 4) return {A_PPB, e, 0, 10};
            return {A_PR, e, 0, 6};
        }
        i -= 6 * L;
        for (int a = 0; a < L; a++)
            for (int b = a + 1; b < L; b++) {
                if (i < 2) return {i == 0 ? A_PQA : A_PQB, a, b, 18};
                i -= 2;
            }
        return {A_FF, 0, 0, 0};
    }
    // j-th product of an atom -> (operand a, operand b)
    struct Pair { int a, b; };
    __host__ __device__ static constexpr Pair atom_pair(AtomDesc at, int j)
    {
        switch (at.kind) {
        case A_DDA: case A_DDB: {
            const int i0 = at.kind == A_DDA ? 0 : KH, i1 = at.kind == A_DDA ? KH : KD;
            for (int i = i0; i < i1; i++) { if (j < KD - i) return {i, i + j}; j -= KD - i; }
            return {0, 0};
        }
        case A_DR: return {j, OP_R};
        case A_DF: return {j / 2, (j % 2) ? OP_C : OP_F};
        case A_FF: return {j == 2 ? OP_C : OP_F, j == 0 ? OP_F : OP_C};
        case A_FR: return {j == 0 ? OP_F : (j == 1 ? OP_C : OP_R), OP_R};
        case A_DPA: return {j / 6, OP_P + 6 * at.e + j % 6};
        case A_DPB: return {KH + j / 6, OP_P + 6 * at.e + j % 6};
        case A_FP: return {j < 6 ? OP_F : OP_C, OP_P + 6 * at.e + j % 6};
        case A_PPA: case A_PPB: {
            const int i0 = at.kind == A_PPA ? 0 : 2, i1 = at.kind == A_PPA ? 2 : 6;
            for (int i = i0; i < i1; i++) { if (j < 6 - i) return {OP_P + 6 * at.e + i, OP_P + 6 * at.e + i + j}; j -= 6 - i; }
            return {0, 0};
        }
        case A_PR: return {OP_P + 6 * at.e + j, OP_R};
        case A_PQA: return {OP_P + 6 * at.e + j / 6, OP_P + 6 * at.b + j % 6};
        default: return {OP_P + 6 * at.e + 3 + j / 6, OP_P + 6 * at.b + j % 6};
        }
    }
    // local column of an operand for a row of parity par
    __host__ __device__ static constexpr int column(int op, int par)
    {
        if (op < KD) return op;
        if (op == OP_F) return KD + par;
        if (op == OP_C) return KD + 2 + par;
        if (op == OP_R) return K + 6 * L;
        return K + (op - OP_P);
    }

    int NT;
    int TOT;                       // accumulators over all tiles
    int off[MAX_TILES];            // first accumulator of a tile in plan order
    int nacc[MAX_TILES];
    short ta[MAX_TILES][MAX_ACC], tb[MAX_TILES][MAX_ACC];
    bool has_ff[MAX_TILES];

    __host__ __device__ constexpr Plan() : NT(0), TOT(0), off{}, nacc{}, ta{}, tb{}, has_ff{}
    {
        int total = 0, order[NA > 0 ? NA : 1] = {};
        for (int i = 0; i < NA; i++) { total += atom(i).n; order[i] = i; }
        int nt = (total + 23) / 24;     // ~22-24 accumulators per tile
        if (nt < 1) nt = 1;
        if (nt > MAX_TILES) nt = MAX_TILES;
        NT = nt;
        // largest first
        for (int i = 0; i < NA; i++)
            for (int j = i + 1; j < NA; j++)
                if (atom(order[j]).n > atom(order[i]).n) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
        for (int i = 0; i < NA; i++) {
            const AtomDesc at = atom(order[i]);
            if (at.n == 0) continue;
            int best = 0;
            for (int t = 1; t < nt; t++) if (nacc[t] < nacc[best]) best = t;
            for (int j = 0; j < at.n; j++) {
                const Pair pr = atom_pair(at, j);
                ta[best][nacc[best]] = (short)pr.a;
                tb[best][nacc[best]] = (short)pr.b;
                nacc[best]++;
            }
            if (at.kind == A_FF) has_ff[best] = true;
        }
        int o = 0;
        for (int t = 0; t < nt; t++) { off[t] = o; o += nacc[t]; }
        TOT = o;
    }
    // packed entry e of the W x W block -> (position in plan order) << 2 | mode
    // mode 0: even-row sum + odd-row sum, 1: even rows only (u), 2: odd rows only (v), 3: structural zero
    __host__ void entry_table(int *table) const
    {
        const int Wd = K + 6 * L + 1, ne = Wd * (Wd + 1) / 2;
        for (int e = 0; e < ne; e++) table[e] = 3;
        for (int t = 0; t < NT; t++)
            for (int j = 0; j < nacc[t]; j++) {
                const int ca0 = column(ta[t][j], 0), cb0 = column(tb[t][j], 0);
                const int ca1 = column(ta[t][j], 1), cb1 = column(tb[t][j], 1);
                const int cu = ca0 * Wd - ca0 * (ca0 - 1) / 2 + (cb0 - ca0);
                const int cv = ca1 * Wd - ca1 * (ca1 - 1) / 2 + (cb1 - ca1);
                const int pos = off[t] + j;
                if (cu == cv) table[cu] = (pos << 2) | 0;
                else { table[cu] = (pos << 2) | 1; table[cv] = (pos << 2) | 2; }
            }
    }
};

template <int KD, int L> struct PlanHolder { static constexpr Plan<KD, L> value{}; };

template <int MODEL, int L> struct Layout {
    static constexpr int K = Camera<MODEL>::K;
    static constexpr int KD = K - 4;              // columns before [fu, fv, u0, v0]
    static constexpr int D = K + 6 * L;
    static constexpr int W = D + 1;
    static constexpr int NE = W * (W + 1) / 2;
    static constexpr int POSE = round_up2(12 + 21 * L);
    static constexpr int NT_ = PlanHolder<K - 4, L>::value.NT;
    // reduction scratch: one 8 x (32 + 2 pad) slab per warp that can own a phase-B work unit
    static constexpr int SCRATCH = (2 * NT_ < 8 ? 2 * NT_ : 8) * 8 * 34;
    static constexpr int NPART = (NE + 95) / 96;  // per-thread slots of the per-CTA block sums
    static constexpr int TOT = PlanHolder<K - 4, L>::value.TOT;   // accumulators of all tiles (plan order)
    // doubles of shared memory: poses of PCG groups + staging of one group of G images
    __host__ __device__ static constexpr long long smem_doubles(int G, int P, int PCG)
    {
        return (long long)PCG * G * POSE + (long long)G * 2 * P * (1 + K + 6 * L) + 4LL * G * TOT + SCRATCH;
    }
};

template <int N>
__device__ __forceinline__ void store_row(double *p, const double (&a)[N], bool aligned16)
{
    if (aligned16) {
#pragma unroll
        for (int i = 0; i + 1 < N; i += 2) *reinterpret_cast<double2 *>(p + i) = make_double2(a[i], a[i + 1]);
        if (N & 1) p[N - 1] = a[N - 1];
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = a[i];
    }
}

// Shared-memory view of the staged group
template <int MODEL, int L> struct Stage {
    using LY = Layout<MODEL, L>;
    double *pose, *rs, *Jas, *Jes[L], *Hs, *scratch;
    __device__ Stage(double *base, int G, int P, int PCG)
    {
        pose = base;            base += (size_t)PCG * G * LY::POSE;
        rs = base;              base += (size_t)G * 2 * P;
        Jas = base;             base += (size_t)G * 2 * P * LY::K;
#pragma unroll
        for (int e = 0; e < L; e++) { Jes[e] = base; base += (size_t)G * 2 * P * 6; }
        scratch = base;         base += LY::SCRATCH;      // keeps 16-byte alignment (all sizes above are even)
        Hs = base;
    }
};

// ---- phase B: one tile ---------------------------------------------------------------------
template <int MODEL, int L, int T>
__device__ __forceinline__ void run_tile(const Stage<MODEL, L> &st, const int warp, const int lane, const int G,
                                         const int S, const int P, const int nv, const int rg, const int RG)
{
    using LY = Layout<MODEL, L>;
    using PL = Plan<LY::KD, L>;
    using PH = PlanHolder<LY::KD, L>;
    constexpr int K = LY::K, KD = LY::KD, W = LY::W, NACC = PH::value.nacc[T];
    if constexpr (NACC == 0) return;
    const int g = lane / S, s = lane - g * S, par = s & 1;
    const bool valid = g < nv;

    // which operands this tile needs
    constexpr auto needs = [](int op) {
        for (int j = 0; j < PH::value.nacc[T]; j++)
            if (PH::value.ta[T][j] == op || PH::value.tb[T][j] == op) return true;
        return false;
    };
    double acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; j++) acc[j] = 0.0;

    if (valid) {
        // this warp's corner range (row group rg of RG); a lane keeps one row parity:
        // rows k = 2 c + par for corners c = c_lo + (s >> 1), step S/2
        const int cpg = (P + RG - 1) / RG;
        const int c_lo = rg * cpg, c_hi = min(P, c_lo + cpg);
        const int hs = S >> 1;
        const int k0 = 2 * (c_lo + (s >> 1)) + par;
        const int nit = (c_hi - c_lo - (s >> 1) + hs - 1) / hs;
        const double *ja = st.Jas + ((size_t)g * 2 * P + k0) * K;
        const double *pr = st.rs + (size_t)g * 2 * P + k0;
        const double *pe[L];
#pragma unroll
        for (int e = 0; e < L; e++) pe[e] = st.Jes[e] + ((size_t)g * 2 * P + k0) * 6;
        const int stepK = S * K, step6 = S * 6;
#pragma unroll 2
        for (int it = 0; it < nit; it++) {
            double op[PL::NOPS];
            static_for<0, KD>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                if constexpr (needs(i)) op[i] = ja[i];
            });
            if constexpr (needs(PL::OP_F)) op[PL::OP_F] = ja[KD + par];
            if constexpr (needs(PL::OP_C)) op[PL::OP_C] = ja[KD + 2 + par];
            if constexpr (needs(PL::OP_R)) op[PL::OP_R] = pr[0];
            ja += stepK;
            pr += S;
            static_for<0, L>([&](auto ec) {
                constexpr int e = decltype(ec)::value;
                constexpr bool any = needs(PL::OP_P + 6 * e) || needs(PL::OP_P + 6 * e + 1) || needs(PL::OP_P + 6 * e + 2) ||
                                     needs(PL::OP_P + 6 * e + 3) || needs(PL::OP_P + 6 * e + 4) || needs(PL::OP_P + 6 * e + 5);
                if constexpr (any) {
                    const double2 *q2 = reinterpret_cast<const double2 *>(pe[e]);
                    pe[e] += step6;
                    static_for<0, 3>([&](auto hc) {
                        constexpr int h = decltype(hc)::value;
                        if constexpr (needs(PL::OP_P + 6 * e + 2 * h) || needs(PL::OP_P + 6 * e + 2 * h + 1)) {
                            const double2 v = q2[h];
                            op[PL::OP_P + 6 * e + 2 * h] = v.x;
                            op[PL::OP_P + 6 * e + 2 * h + 1] = v.y;
                        }
                    });
                }
            });
            static_for<0, NACC>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int a = PH::value.ta[T][j], b = PH::value.tb[T][j];
                acc[j] = fma(op[a], op[b], acc[j]);
            });
        }
    }

    // ---- combine the S row splits through shared memory, 8 accumulators at a time: lane (qg, qa)
    //      sums the S partials of accumulator 8c + qa of image qg, even and odd rows apart, and
    //      leaves the pair in plan order; the store loop maps packed entries onto these slots
    double *sc = st.scratch + (size_t)warp * (8 * 34);
    double2 *hplan = reinterpret_cast<double2 *>(st.Hs) + (size_t)rg * G * LY::TOT;
    const int qg = lane >> 3, qa = lane & 7;
    constexpr int NCH = (NACC + 7) / 8;
    static_for<0, NCH>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        static_for<0, 8>([&](auto ac) {
            constexpr int j = 8 * c + decltype(ac)::value;
            if constexpr (j < NACC) sc[decltype(ac)::value * 34 + lane] = acc[j];
        });
        __syncwarp();
        const int n = 8 * c + qa;
        if (qg < G && n < NACC) {
            double ev = 0.0, od = 0.0;
            const double2 *src = reinterpret_cast<const double2 *>(sc + qa * 34 + qg * S);
            for (int q = 0; q < S / 2; q++) { const double2 v = src[q]; ev += v.x; od += v.y; }
            hplan[(size_t)qg * LY::TOT + PH::value.off[T] + n] = make_double2(ev, od);
        }
        __syncwarp();
    });
}

template <int MODEL, int L, int T>
__device__ __forceinline__ void dispatch_tile(const int t, const Stage<MODEL, L> &st, const int warp, const int lane,
                                              const int G, const int S, const int P, const int nv, const int rg,
                                              const int RG)
{
    if constexpr (T < PlanHolder<Layout<MODEL, L>::KD, L>::value.NT) {
        if (t == T) run_tile<MODEL, L, T>(st, warp, lane, G, S, P, nv, rg, RG);
        else dispatch_tile<MODEL, L, T + 1>(t, st, warp, lane, G, S, P, nv, rg, RG);
    }
}

// ---- phase 0: one image's transform chain -> staged pose record ---------------------
template <int L>
__device__ __forceinline__ void chain_pose(const EvalArgs &args, const int img, double *ps)
{
    double Racc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double tacc[3] = {0, 0, 0};
    const int sidx = args.seq_index ? args.seq_index[img] : img;
#pragma unroll
    for (int e = 0; e < L; e++) {
        const double *x = args.xi[e] + (size_t)sidx * args.xi_stride[e];
        const double t0 = x[0], t1 = x[1], t2 = x[2];
        double Re[9], Jl[9], R12[9], M12[9], t13[3];
        rodrigues_and_left_jacobian(x[3], x[4], x[5], Re, Jl);
        if (!args.inverse[e]) {
            // X1 = T_acc T_e X : xi13 is the chain after composing this element
#pragma unroll
            for (int i = 0; i < 9; i++) R12[i] = Racc[i];
#pragma unroll
            for (int i = 0; i < 3; i++)
                t13[i] = fma(Racc[3 * i + 2], t2, fma(Racc[3 * i + 1], t1, fma(Racc[3 * i], t0, tacc[i])));
            mat3_mul(R12, Jl, M12);
            double Rn[9];
            mat3_mul(Racc, Re, Rn);
#pragma unroll
            for (int i = 0; i < 9; i++) Racc[i] = Rn[i];
#pragma unroll
            for (int i = 0; i < 3; i++) tacc[i] = t13[i];
        } else {
            // X1 = T_acc T_e^-1 X : xi13 is the chain before this element; kinematic screw inverted
            double Rn[9], RnJ[9];
            mat3_mul_bt(Racc, Re, Rn);
            mat3_mul(Rn, Jl, RnJ);
#pragma unroll
            for (int i = 0; i < 9; i++) { R12[i] = -Rn[i]; M12[i] = -RnJ[i]; }
#pragma unroll
            for (int i = 0; i < 3; i++) {
                t13[i] = tacc[i];
                tacc[i] = tacc[i] - fma(Rn[3 * i + 2], t2, fma(Rn[3 * i + 1], t1, Rn[3 * i] * t0));
            }
#pragma unroll
            for (int i = 0; i < 9; i++) Racc[i] = Rn[i];
        }
        double *pe = ps + 12 + 21 * e;
#pragma unroll
        for (int i = 0; i < 9; i++) { pe[i] = R12[i]; pe[9 + i] = M12[i]; }
#pragma unroll
        for (int i = 0; i < 3; i++) pe[18 + i] = t13[i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) ps[i] = Racc[i];
#pragma unroll
    for (int i = 0; i < 3; i++) ps[9 + i] = tacc[i];
}

// ---- the kernel -------------------------------------------------------------------
template <int MODEL, int L>
__global__ void __launch_bounds__(256, (L == 1 && Camera<MODEL>::K <= 6) ? 3 : 2)
reproj_eval_kernel(const EvalArgs args, const int G, const int PCG, const int *__restrict__ entry_table)
{
    using LY = Layout<MODEL, L>;
    using CAM = Camera<MODEL>;
    constexpr int K = LY::K;
    constexpr int NT = PlanHolder<LY::KD, L>::value.NT;
    extern __shared__ __align__(16) double smem[];
    const int P = args.P;
    const Stage<MODEL, L> st(smem, G, P, PCG);
    const int tid = threadIdx.x;
    const int n_groups = (args.n_img + G - 1) / G;
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const int S = 32 / G;

    double intr[K];
#pragma unroll
    for (int i = 0; i < K; i++) intr[i] = __ldg(args.intr + i);
    const typename CAM::Consts cc = CAM::prepare(intr);
    const bool first_direct = (args.inverse[0] == 0);   // R12 of element 0 is the identity
    double part[LY::NPART];                              // this CTA's sum of its images' blocks
    int code[LY::NPART];                                 // packed entry -> plan-order slot (see Plan::entry_table)
#pragma unroll
    for (int q = 0; q < LY::NPART; q++) {
        part[q] = 0.0;
        const int e = tid + q * blockDim.x;
        code[q] = (args.H && e < LY::NE) ? __ldg(entry_table + e) : 3;
    }

    for (int j0 = 0; (long long)blockIdx.x + (long long)j0 * gridDim.x < n_groups; j0 += PCG) {
        // ---- phase 0: poses of this CTA's next PCG groups, one thread per image ---------
        if (tid < PCG * G) {
            const int j = tid / G, gi = tid - j * G;
            const long long grp = (long long)blockIdx.x + (long long)(j0 + j) * gridDim.x;
            const long long img = grp * G + gi;
            if (grp < n_groups && img < args.n_img) chain_pose<L>(args, (int)img, st.pose + (size_t)tid * LY::POSE);
        }
        __syncthreads();

        for (int j = 0; j < PCG; j++) {
            const long long grp = (long long)blockIdx.x + (long long)(j0 + j) * gridDim.x;
            if (grp >= n_groups) break;
            const int img0 = (int)grp * G;
            const int nv = min(G, args.n_img - img0);
            const double *pose_grp = st.pose + (size_t)j * G * LY::POSE;

            // ---- phase A: one thread per (image, corner) -------------------------------
            for (int idx = tid; idx < nv * P; idx += blockDim.x) {
                const int g = idx / P;
                const int c = idx - g * P;
                const int img = img0 + g;
                const double *ps = pose_grp + (size_t)g * LY::POSE;
                // issue the (streaming) observation load first: its DRAM latency hides behind the model math
                double2 ob;
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                             : "=d"(ob.x), "=d"(ob.y) : "l"(args.obs + ((size_t)img * P + c) * 2));
                const double bx = __ldg(args.board + 3 * c), by = __ldg(args.board + 3 * c + 1),
                             bz = __ldg(args.board + 3 * c + 2);
                const double X0 = fma(ps[2], bz, fma(ps[1], by, fma(ps[0], bx, ps[9])));
                const double X1 = fma(ps[5], bz, fma(ps[4], by, fma(ps[3], bx, ps[10])));
                const double X2 = fma(ps[8], bz, fma(ps[7], by, fma(ps[6], bx, ps[11])));
                double u, v, Pu[3], Pv[3], Ju[K], Jv[K];
                const bool ok = CAM::eval(intr, cc, X0, X1, X2, u, v, Pu, Pv, Ju, Jv);
                double2 res;
                if (ok) {
                    res.x = u - ob.x;
                    res.y = v - ob.y;
                } else {
                    res.x = DOUBLE_BIG;
                    res.y = DOUBLE_BIG;
#pragma unroll
                    for (int i = 0; i < 3; i++) { Pu[i] = 0.0; Pv[i] = 0.0; }
#pragma unroll
                    for (int i = 0; i < K; i++) { Ju[i] = 0.0; Jv[i] = 0.0; }
                }
                const size_t row = (size_t)g * 2 * P + 2 * c;
                *reinterpret_cast<double2 *>(st.rs + row) = res;
                store_row<K>(st.Jas + row * K, Ju, (K % 2) == 0);
                store_row<K>(st.Jas + (row + 1) * K, Jv, (K % 2) == 0);
#pragma unroll
                for (int e = 0; e < L; e++) {
                    const double *pe = ps + 12 + 21 * e;
                    const double w0 = X0 - pe[18], w1 = X1 - pe[19], w2 = X2 - pe[20];
                    // (w x p)^T M12  ==  -p^T hat(w) M12   (jacobian.h:165,170)
                    const double cu0 = w1 * Pu[2] - w2 * Pu[1], cu1 = w2 * Pu[0] - w0 * Pu[2],
                                 cu2 = w0 * Pu[1] - w1 * Pu[0];
                    const double cv0 = w1 * Pv[2] - w2 * Pv[1], cv1 = w2 * Pv[0] - w0 * Pv[2],
                                 cv2 = w0 * Pv[1] - w1 * Pv[0];
                    double ju[6], jv[6];
                    if (e == 0 && first_direct) {
#pragma unroll
                        for (int q = 0; q < 3; q++) { ju[q] = Pu[q]; jv[q] = Pv[q]; }
                    } else {
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            ju[q] = fma(Pu[2], pe[6 + q], fma(Pu[1], pe[3 + q], Pu[0] * pe[q]));
                            jv[q] = fma(Pv[2], pe[6 + q], fma(Pv[1], pe[3 + q], Pv[0] * pe[q]));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        ju[3 + q] = fma(cu2, pe[15 + q], fma(cu1, pe[12 + q], cu0 * pe[9 + q]));
                        jv[3 + q] = fma(cv2, pe[15 + q], fma(cv1, pe[12 + q], cv0 * pe[9 + q]));
                    }
                    store_row<6>(st.Jes[e] + row * 6, ju, true);
                    store_row<6>(st.Jes[e] + (row + 1) * 6, jv, true);
                }
            }
            fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA engine
            __syncthreads();

            // ---- phase S: stream the Ceres-layout blocks out with TMA bulk copies -------
            bool issued = false;
            if (tid == 0) {
                const size_t rows = (size_t)nv * 2 * P;
                if (args.r) { bulk_store_chunked(args.r + (size_t)img0 * 2 * P, st.rs, rows * 8); issued = true; }
                if (args.Ja) { bulk_store_chunked(args.Ja + (size_t)img0 * 2 * P * K, st.Jas, rows * K * 8); issued = true; }
#pragma unroll
                for (int e = 0; e < L; e++)
                    if (args.Je[e]) {
                        bulk_store_chunked(args.Je[e] + (size_t)img0 * 2 * P * 6, st.Jes[e], rows * 48);
                        issued = true;
                    }
                if (issued) bulk_commit();
            }

            // ---- phase B: per-image normal-equation blocks -----------------------------
            if (args.H) {
                // NT tiles x RG row groups of work units; with RG = 2 each tile's rows are shared by two
                // warps that fill separate copies of the blocks (summed below: two addends, deterministic)
                const int RG = (nw >= 2 * NT && P >= 2) ? 2 : 1;
                for (int u = warp; u < NT * RG; u += nw)
                    dispatch_tile<MODEL, L, 0>(u % NT, st, warp, lane, G, S, P, nv, u / NT, RG);
                __syncthreads();
                double *Hg = args.H + (size_t)img0 * LY::NE;
#pragma unroll
                for (int q = 0; q < LY::NPART; q++) {
                    const int e = tid + q * blockDim.x;
                    if (e < LY::NE) {
                        double sum = 0.0;
                        const int pos = code[q] >> 2, mode = code[q] & 3;
                        const double2 *hp = reinterpret_cast<const double2 *>(st.Hs) + pos;
                        for (int g = 0; g < nv; g++) {
                            double2 v2 = hp[(size_t)g * LY::TOT];
                            if (RG == 2) {
                                const double2 w2 = hp[(size_t)(G + g) * LY::TOT];
                                v2.x += w2.x; v2.y += w2.y;
                            }
                            const double val = mode == 0 ? v2.x + v2.y : (mode == 1 ? v2.x : (mode == 2 ? v2.y : 0.0));
                            Hg[(size_t)g * LY::NE + e] = val;
                            sum += val;
                        }
                        part[q] += sum;
                    }
                }
            }
            if (issued) bulk_wait_read_all();   // staging must outlive the TMA reads
            __syncthreads();                    // staging + Hs are free for the next group
        }
    }
    if (args.H && args.cta_partial) {
#pragma unroll
        for (int q = 0; q < LY::NPART; q++) {
            const int e = tid + q * blockDim.x;
            if (e < LY::NE) args.cta_partial[(size_t)blockIdx.x * LY::NE + e] = part[q];
        }
    }
}


template <int MODEL, int L>
cudaError_t launch_one(const EvalArgs &args, cudaStream_t stream, unsigned long long *launches, int *grid_out,
                       bool query_only)
{
    LaunchPlan pl;
    if (!plan_eval(MODEL, L, args.P, &pl)) return cudaErrorInvalidValue;
    static int configured_bytes[64];    // per instantiation and device; zero-initialised
    static int blocks_per_sm[64];
    static int sm_count[64];
    static int planned_threads[64];
    static int *entry_table[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!entry_table[dev] && !query_only) {
        using LY = Layout<MODEL, L>;
        static constexpr Plan<LY::KD, L> plan{};
        std::vector<int> tab(LY::NE);
        plan.entry_table(tab.data());
        cudaError_t e = cudaMalloc(&entry_table[dev], sizeof(int) * LY::NE);
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(entry_table[dev], tab.data(), sizeof(int) * LY::NE, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
    }
    if (pl.smem > configured_bytes[dev] || planned_threads[dev] != pl.threads) {
        cudaError_t e = cudaFuncSetAttribute(reproj_eval_kernel<MODEL, L>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return e;
        configured_bytes[dev] = (int)pl.smem;
        planned_threads[dev] = pl.threads;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[dev], reproj_eval_kernel<MODEL, L>,
                                                          pl.threads, (size_t)pl.smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm[dev] < 1) blocks_per_sm[dev] = 1;
    }
    const int n_groups = (args.n_img + pl.G - 1) / pl.G;
    int grid = sm_count[dev] * blocks_per_sm[dev];   // one resident wave of persistent CTAs
    if (grid > n_groups) grid = n_groups;
    if (grid_out) *grid_out = grid;
    if (query_only || args.n_img <= 0) return cudaSuccess;
    reproj_eval_kernel<MODEL, L><<<grid, pl.threads, (size_t)pl.smem, stream>>>(args, pl.G, pl.PCG, entry_table[dev]);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_model(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n, int *grid, bool q)
{
    switch (L) {
    case 1: return launch_one<MODEL, 1>(a, s, n, grid, q);
    case 2: return launch_one<MODEL, 2>(a, s, n, grid, q);
    case 3: return launch_one<MODEL, 3>(a, s, n, grid, q);
    case 4: return launch_one<MODEL, 4>(a, s, n, grid, q);
    case 5: return launch_one<MODEL, 5>(a, s, n, grid, q);
    default: return cudaErrorInvalidValue;
    }
}

template <int MODEL> long long smem_for(int L, int G, int P, int PCG)
{
    switch (L) {
    case 1: return Layout<MODEL, 1>::smem_doubles(G, P, PCG) * 8;
    case 2: return Layout<MODEL, 2>::smem_doubles(G, P, PCG) * 8;
    case 3: return Layout<MODEL, 3>::smem_doubles(G, P, PCG) * 8;
    case 4: return Layout<MODEL, 4>::smem_doubles(G, P, PCG) * 8;
    case 5: return Layout<MODEL, 5>::smem_doubles(G, P, PCG) * 8;
    default: return -1;
    }
}

}  // namespace

}  // namespace vg
