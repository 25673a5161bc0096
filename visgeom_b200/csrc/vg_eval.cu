// vg_eval.cu -- fused reprojection residual + analytic Jacobian + per-image
// normal-equation kernel for sm_100a (one instantiation per camera model and
// chain length).
//
// Replaces, for every image of a dataset, GenericProjectionJac::Evaluate
// (src/calibration/calib_cost_functions.cpp:28-117) and the J^T J / J^T r build
// Ceres performs on the block it returns.
//
// One CTA processes a group of G consecutive images in three phases:
//   0. pose   : G threads accumulate each image's transform chain with rotation
//               matrices and stage R, t and, per chain element, R12, M12, t13
//               (InterJacobian's members, jacobian.h:139-152) in shared memory.
//   A. corner : one thread per (image, corner): X = R Xb + t, one shared
//               evaluation of the camera model (projection, dP/dX, dP/dintr),
//               residual and Jacobian rows written into shared memory in exactly
//               the Ceres block layout.  Observations are read with coalesced
//               16-byte loads.
//   S. store  : the staged blocks of the G images are contiguous in global memory,
//               so one elected thread streams each region out with a TMA bulk
//               copy (cp.async.bulk shared::cta -> global); no register round trip.
//   B. normal : warps re-read the staged rows and accumulate the per-image packed
//               upper triangle of [J r]^T [J r] in register tiles (block pairs of
//               <= 6 x 6), rows split over S = 32/G lanes and combined with warp
//               shuffles; the G blocks leave through coalesced stores.
// The kernel is HBM-write bound by design (224 B written per EUCM corner); tensor
// cores are not used -- there is no dense contraction on this path.
#include "vg_eval.cuh"
#include "vg_math.cuh"

#include <cstdint>

namespace vg {

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

__host__ __device__ constexpr int round_up2(int x) { return (x + 1) & ~1; }

template <int KD, int L> __host__ __device__ constexpr int num_vtiles();

template <int MODEL, int L> struct Layout {
    static constexpr int K = Camera<MODEL>::K;
    static constexpr int KD = K - 4;              // columns before [fu, fv, u0, v0]
    static constexpr int D = K + 6 * L;
    static constexpr int W = D + 1;
    static constexpr int NE = W * (W + 1) / 2;
    static constexpr int NVT = num_vtiles<KD, L>();   // phase-B virtual tiles
    static constexpr int POSE = round_up2(12 + 21 * L);
    // doubles of shared memory: poses of PCG groups + staging of one group of G images
    __host__ __device__ static constexpr long long smem_doubles(int G, int P, int PCG)
    {
        return (long long)PCG * G * POSE + (long long)G * 2 * P * (1 + K + 6 * L) + (long long)G * NE;
    }
};

// packed upper-triangular index of (a,b), a <= b, in a W x W symmetric matrix
__host__ __device__ constexpr int pk(int a, int b, int W) { return a * W - a * (a - 1) / 2 + (b - a); }

// load N consecutive doubles; 16-byte vector loads where the compile-time phase allows
template <int N, bool ROW_EVEN, int PHASE>
__device__ __forceinline__ void load_row(const double *p, double (&a)[N])
{
    if constexpr (!ROW_EVEN) {
#pragma unroll
        for (int i = 0; i < N; i++) a[i] = p[i];
    } else {
        int i = 0;
        if constexpr (PHASE == 1) { a[0] = p[0]; i = 1; }
#pragma unroll
        for (; i + 1 < N; i += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(p + i);
            a[i] = v.x; a[i + 1] = v.y;
        }
        if (i < N) a[i] = p[i];
    }
}

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
    double *pose, *rs, *Jas, *Jes[L], *Hs;
    __device__ Stage(double *base, int G, int P, int PCG)
    {
        pose = base;            base += (size_t)PCG * G * LY::POSE;
        rs = base;              base += (size_t)G * 2 * P;
        Jas = base;             base += (size_t)G * 2 * P * LY::K;
#pragma unroll
        for (int e = 0; e < L; e++) { Jes[e] = base; base += (size_t)G * 2 * P * 6; }
        Hs = base;
    }
};

// ---- phase B: sparsity-aware normal-equation tiles ---------------------------------
// Every model's intrinsic Jacobian row has the shape (eucm.h:208-222, ucm.h:176-192,
// mei.h:257-283)
//     u-row: [ d_0 .. d_{KD-1} | f 0 | 1 0 ]      v-row: [ d_0 .. d_{KD-1} | 0 f | 0 1 ]
// (KD "dense" distortion columns, then fu,fv, then u0,v0), so a row is described by
// d[KD], one focal value f, the centre value c (1, or 0 for a failed projection) and
// its parity; the structural zeros are never multiplied.  A lane always sees rows of
// one parity (row index = lane + multiple of an even S), so the code stays uniform
// and only the output index depends on the parity.
struct TileSpec {
    bool dd, dr, df, ff, fr, rr;   // d x d (sym), d x r, d x {f,c}, {f,c} x {f,c}, {f,c} x r, r x r
    int e;                         // chain element of the p-groups below, -1: none
    bool dp, fp, pp, pr;           // d x p_e, {f,c} x p_e, p_e x p_e (sym), p_e x r
    int b;                         // >= 0: cross block p_e x p_b (e < b)
};

template <int KD, int L> __host__ __device__ constexpr int num_vtiles()
{
    return (KD >= 3 ? 2 + 3 * L : 1 + 2 * L) + L * (L - 1) / 2;
}

template <int KD, int L> __host__ __device__ constexpr TileSpec vtile_spec(int t)
{
    TileSpec s{false, false, false, false, false, false, -1, false, false, false, false, -1};
    if (KD >= 3) {
        if (t == 0) { s.dd = s.dr = true; return s; }
        if (t == 1) { s.df = s.ff = s.fr = s.rr = true; return s; }
        t -= 2;
        if (t < 3 * L) {
            s.e = t / 3;
            if (t % 3 == 0) s.dp = true;
            else if (t % 3 == 1) s.fp = s.pr = true;
            else s.pp = true;
            return s;
        }
        t -= 3 * L;
    } else {
        if (t == 0) {
            s.dd = s.dr = s.df = s.ff = s.fr = s.rr = true;
            if (L == 1) { s.e = 0; s.pr = true; }
            return s;
        }
        t -= 1;
        if (t < 2 * L) {
            s.e = t / 2;
            if (t % 2 == 0) s.dp = s.fp = true;
            else { s.pp = true; s.pr = (L != 1); }
            return s;
        }
        t -= 2 * L;
    }
    // cross blocks p_a x p_b, a < b
    for (int a = 0; a < L; a++)
        for (int b = a + 1; b < L; b++) {
            if (t == 0) { s.e = a; s.b = b; return s; }
            t--;
        }
    return s;
}

template <int N>
__device__ __forceinline__ void xor_reduce(double (&a)[N], const int off)
{
#pragma unroll
    for (int i = 0; i < N; i++) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
}

template <int MODEL, int L, int VT>
__device__ __forceinline__ void run_vtile(const Stage<MODEL, L> &st, const int g, const int s, const int S,
                                          const int P, const bool valid)
{
    using LY = Layout<MODEL, L>;
    constexpr int K = LY::K, KD = LY::KD, D = LY::D, W = LY::W;
    constexpr TileSpec sp = vtile_spec<KD, L>(VT);
    constexpr bool HAS_E = sp.e >= 0, HAS_B = sp.b >= 0;
    constexpr int E = HAS_E ? sp.e : 0, B = HAS_B ? sp.b : 0;
    constexpr bool need_d = sp.dd || sp.dr || sp.df || sp.dp;
    constexpr bool need_f = sp.df || sp.ff || sp.fr || sp.fp;
    constexpr bool need_r = sp.dr || sp.fr || sp.rr || sp.pr;
    constexpr bool need_p = sp.dp || sp.fp || sp.pp || sp.pr || HAS_B;
    constexpr int NDD = KD * (KD + 1) / 2;
    const int par = s & 1;

    double a_dd[NDD], a_dr[KD], a_df[KD], a_dc[KD], a_ffc[3], a_fcr[2], a_rr[1];
    double a_dp[KD * 6], a_fp[6], a_cp[6], a_pp[21], a_pr[6], a_pq[36];
#pragma unroll
    for (int i = 0; i < NDD; i++) a_dd[i] = 0.0;
#pragma unroll
    for (int i = 0; i < KD; i++) { a_dr[i] = 0.0; a_df[i] = 0.0; a_dc[i] = 0.0; }
    a_ffc[0] = a_ffc[1] = a_ffc[2] = 0.0; a_fcr[0] = a_fcr[1] = 0.0; a_rr[0] = 0.0;
#pragma unroll
    for (int i = 0; i < KD * 6; i++) a_dp[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) { a_fp[i] = 0.0; a_cp[i] = 0.0; a_pr[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 21; i++) a_pp[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 36; i++) a_pq[i] = 0.0;

    if (valid) {
        const double *ja = st.Jas + (size_t)g * 2 * P * K;
        const double *je = st.Jes[E] + (size_t)g * 2 * P * 6;
        const double *jb = st.Jes[B] + (size_t)g * 2 * P * 6;
        const double *pr = st.rs + (size_t)g * 2 * P;
        for (int k = s; k < 2 * P; k += S) {
            double d[KD], p[6], q[6], f = 0.0, c = 0.0, r = 0.0;
            if constexpr (need_d) load_row<KD, (K % 2 == 0), 0>(ja + (size_t)k * K, d);
            if constexpr (need_f) { f = ja[(size_t)k * K + KD + par]; c = ja[(size_t)k * K + KD + 2 + par]; }
            if constexpr (need_r) r = pr[k];
            if constexpr (need_p) load_row<6, true, 0>(je + (size_t)k * 6, p);
            if constexpr (HAS_B) load_row<6, true, 0>(jb + (size_t)k * 6, q);
            if constexpr (sp.dd) {
                int n = 0;
#pragma unroll
                for (int i = 0; i < KD; i++)
#pragma unroll
                    for (int j = i; j < KD; j++) { a_dd[n] = fma(d[i], d[j], a_dd[n]); n++; }
            }
            if constexpr (sp.dr) {
#pragma unroll
                for (int i = 0; i < KD; i++) a_dr[i] = fma(d[i], r, a_dr[i]);
            }
            if constexpr (sp.df) {
#pragma unroll
                for (int i = 0; i < KD; i++) { a_df[i] = fma(d[i], f, a_df[i]); a_dc[i] = fma(d[i], c, a_dc[i]); }
            }
            if constexpr (sp.ff) {
                a_ffc[0] = fma(f, f, a_ffc[0]); a_ffc[1] = fma(f, c, a_ffc[1]); a_ffc[2] = fma(c, c, a_ffc[2]);
            }
            if constexpr (sp.fr) { a_fcr[0] = fma(f, r, a_fcr[0]); a_fcr[1] = fma(c, r, a_fcr[1]); }
            if constexpr (sp.rr) a_rr[0] = fma(r, r, a_rr[0]);
            if constexpr (sp.dp) {
#pragma unroll
                for (int i = 0; i < KD; i++)
#pragma unroll
                    for (int j = 0; j < 6; j++) a_dp[i * 6 + j] = fma(d[i], p[j], a_dp[i * 6 + j]);
            }
            if constexpr (sp.fp) {
#pragma unroll
                for (int j = 0; j < 6; j++) { a_fp[j] = fma(f, p[j], a_fp[j]); a_cp[j] = fma(c, p[j], a_cp[j]); }
            }
            if constexpr (sp.pp) {
                int n = 0;
#pragma unroll
                for (int i = 0; i < 6; i++)
#pragma unroll
                    for (int j = i; j < 6; j++) { a_pp[n] = fma(p[i], p[j], a_pp[n]); n++; }
            }
            if constexpr (sp.pr) {
#pragma unroll
                for (int j = 0; j < 6; j++) a_pr[j] = fma(p[j], r, a_pr[j]);
            }
            if constexpr (HAS_B) {
#pragma unroll
                for (int i = 0; i < 6; i++)
#pragma unroll
                    for (int j = 0; j < 6; j++) a_pq[i * 6 + j] = fma(p[i], q[j], a_pq[i * 6 + j]);
            }
        }
    }
    // combine the row splits: same-parity lanes first (offsets S/2 .. 2), then the two
    // parities (offset 1) for the groups whose destination does not depend on parity
    for (int off = S >> 1; off >= 1; off >>= 1) {
        const bool same_parity = off >= 2;
        if constexpr (sp.dd) xor_reduce(a_dd, off);
        if constexpr (sp.dr) xor_reduce(a_dr, off);
        if constexpr (sp.rr) xor_reduce(a_rr, off);
        if constexpr (sp.dp) xor_reduce(a_dp, off);
        if constexpr (sp.pp) xor_reduce(a_pp, off);
        if constexpr (sp.pr) xor_reduce(a_pr, off);
        if constexpr (HAS_B) xor_reduce(a_pq, off);
        if (same_parity) {
            if constexpr (sp.df) { xor_reduce(a_df, off); xor_reduce(a_dc, off); }
            if constexpr (sp.ff) xor_reduce(a_ffc, off);
            if constexpr (sp.fr) xor_reduce(a_fcr, off);
            if constexpr (sp.fp) { xor_reduce(a_fp, off); xor_reduce(a_cp, off); }
        }
    }
    if (!valid || s > 1) return;
    double *h = st.Hs + (size_t)g * LY::NE;
    const int fc = KD + par, cc = KD + 2 + par;     // this parity's focal / centre column
    // parity-dependent entries: lane s = 0 holds the u rows, lane s = 1 the v rows
    if constexpr (sp.df) {
#pragma unroll
        for (int i = 0; i < KD; i++) { h[pk(i, fc, W)] = a_df[i]; h[pk(i, cc, W)] = a_dc[i]; }
    }
    if constexpr (sp.ff) { h[pk(fc, fc, W)] = a_ffc[0]; h[pk(fc, cc, W)] = a_ffc[1]; h[pk(cc, cc, W)] = a_ffc[2]; }
    if constexpr (sp.fr) { h[pk(fc, D, W)] = a_fcr[0]; h[pk(cc, D, W)] = a_fcr[1]; }
    if constexpr (sp.fp) {
#pragma unroll
        for (int j = 0; j < 6; j++) { h[pk(fc, K + 6 * E + j, W)] = a_fp[j]; h[pk(cc, K + 6 * E + j, W)] = a_cp[j]; }
    }
    if (s != 0) return;
    if constexpr (sp.ff) {   // u and v rows never meet in these four entries
        h[pk(KD, KD + 1, W)] = 0.0; h[pk(KD, KD + 3, W)] = 0.0;
        h[pk(KD + 1, KD + 2, W)] = 0.0; h[pk(KD + 2, KD + 3, W)] = 0.0;
    }
    if constexpr (sp.dd) {
        int n = 0;
#pragma unroll
        for (int i = 0; i < KD; i++)
#pragma unroll
            for (int j = i; j < KD; j++) h[pk(i, j, W)] = a_dd[n++];
    }
    if constexpr (sp.dr) {
#pragma unroll
        for (int i = 0; i < KD; i++) h[pk(i, D, W)] = a_dr[i];
    }
    if constexpr (sp.rr) h[pk(D, D, W)] = a_rr[0];
    if constexpr (sp.dp) {
#pragma unroll
        for (int i = 0; i < KD; i++)
#pragma unroll
            for (int j = 0; j < 6; j++) h[pk(i, K + 6 * E + j, W)] = a_dp[i * 6 + j];
    }
    if constexpr (sp.pp) {
        int n = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++) h[pk(K + 6 * E + i, K + 6 * E + j, W)] = a_pp[n++];
    }
    if constexpr (sp.pr) {
#pragma unroll
        for (int j = 0; j < 6; j++) h[pk(K + 6 * E + j, D, W)] = a_pr[j];
    }
    if constexpr (HAS_B) {
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = 0; j < 6; j++) h[pk(K + 6 * E + i, K + 6 * B + j, W)] = a_pq[i * 6 + j];
    }
}

template <int MODEL, int L, int T>
__device__ __forceinline__ void dispatch_tile(const int t, const Stage<MODEL, L> &st, const int g, const int s,
                                              const int S, const int P, const bool valid)
{
    if constexpr (T < Layout<MODEL, L>::NVT) {
        if (t == T) run_vtile<MODEL, L, T>(st, g, s, S, P, valid);
        else dispatch_tile<MODEL, L, T + 1>(t, st, g, s, S, P, valid);
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
// Persistent: CTA b owns image groups b, b + gridDim.x, ... (static, deterministic).
// PCG = groups whose poses one prologue pass stages (PCG * G <= blockDim.x).
template <int MODEL, int L>
__global__ void __launch_bounds__(256, (L == 1 && Camera<MODEL>::K <= 6) ? 3 : 2)
reproj_eval_kernel(const EvalArgs args, const int G, const int PCG)
{
    using LY = Layout<MODEL, L>;
    using CAM = Camera<MODEL>;
    constexpr int K = LY::K;
    extern __shared__ __align__(16) double smem[];
    const int P = args.P;
    const Stage<MODEL, L> st(smem, G, P, PCG);
    const int tid = threadIdx.x;
    const int n_groups = (args.n_img + G - 1) / G;
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const int S = 32 / G;
    const int gB = lane / S, sB = lane - gB * S;

    double intr[K];
#pragma unroll
    for (int i = 0; i < K; i++) intr[i] = __ldg(args.intr + i);
    const bool first_direct = (args.inverse[0] == 0);   // R12 of element 0 is the identity

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
                const double bx = __ldg(args.board + 3 * c), by = __ldg(args.board + 3 * c + 1),
                             bz = __ldg(args.board + 3 * c + 2);
                const double X0 = fma(ps[2], bz, fma(ps[1], by, fma(ps[0], bx, ps[9])));
                const double X1 = fma(ps[5], bz, fma(ps[4], by, fma(ps[3], bx, ps[10])));
                const double X2 = fma(ps[8], bz, fma(ps[7], by, fma(ps[6], bx, ps[11])));
                double u, v, Pu[3], Pv[3], Ju[K], Jv[K];
                const bool ok = CAM::eval(intr, X0, X1, X2, u, v, Pu, Pv, Ju, Jv);
                const double2 ob = *reinterpret_cast<const double2 *>(args.obs + ((size_t)img * P + c) * 2);
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
                for (int t = warp; t < LY::NVT; t += nw) dispatch_tile<MODEL, L, 0>(t, st, gB, sB, S, P, gB < nv);
                __syncthreads();
                double *Hg = args.H + (size_t)img0 * LY::NE;
                for (int i = tid; i < nv * LY::NE; i += blockDim.x) Hg[i] = st.Hs[i];
            }
            if (issued) bulk_wait_read_all();   // staging must outlive the TMA reads
            __syncthreads();                    // staging + Hs are free for the next group
        }
    }
}

struct LaunchPlan { int G, threads, PCG, grid_cap; long long smem; };
bool plan_eval(int model, int L, int P, LaunchPlan *pl);

template <int MODEL, int L>
cudaError_t launch_one(const EvalArgs &args, cudaStream_t stream, unsigned long long *launches)
{
    LaunchPlan pl;
    if (!plan_eval(MODEL, L, args.P, &pl)) return cudaErrorInvalidValue;
    static int configured_bytes[64];    // per instantiation and device; zero-initialised
    static int blocks_per_sm[64];
    static int sm_count[64];
    static int planned_threads[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
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
    if (args.n_img <= 0) return cudaSuccess;
    const int n_groups = (args.n_img + pl.G - 1) / pl.G;
    int grid = sm_count[dev] * blocks_per_sm[dev];   // one resident wave of persistent CTAs
    if (grid > n_groups) grid = n_groups;
    reproj_eval_kernel<MODEL, L><<<grid, pl.threads, (size_t)pl.smem, stream>>>(args, pl.G, pl.PCG);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_model(int L, const EvalArgs &a, cudaStream_t s, unsigned long long *n)
{
    switch (L) {
    case 1: return launch_one<MODEL, 1>(a, s, n);
    case 2: return launch_one<MODEL, 2>(a, s, n);
    case 3: return launch_one<MODEL, 3>(a, s, n);
    case 4: return launch_one<MODEL, 4>(a, s, n);
    case 5: return launch_one<MODEL, 5>(a, s, n);
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

long long smem_any(int model, int L, int G, int P, int PCG)
{
    switch (model) {
    case MODEL_EUCM: return smem_for<MODEL_EUCM>(L, G, P, PCG);
    case MODEL_UCM: return smem_for<MODEL_UCM>(L, G, P, PCG);
    case MODEL_MEI: return smem_for<MODEL_MEI>(L, G, P, PCG);
    default: return -1;
    }
}

// G images per group (power of two, G*P <= 256 so one corner pass covers a group),
// PCG groups per pose prologue (about 16 KB of pose records, PCG*G <= threads).
bool plan_eval(int model, int L, int P, LaunchPlan *pl)
{
    if (P < 1 || L < 1 || L > MAX_CHAIN) return false;
    const long long LIMIT = 227 * 1024;
    const int pose_doubles = ((12 + 21 * L) + 1) & ~1;
    for (int G = 8; G >= 1; G >>= 1) {
        if (G > 1 && G * P > 256) continue;
        int t = ((G * P + 31) / 32) * 32;
        if (t < 96) t = 96;
        if (t > 256) t = 256;
        int pcg = (16 * 1024) / (pose_doubles * 8 * G);
        if (pcg * G > t) pcg = t / G;
        if (pcg < 1) pcg = 1;
        long long b = smem_any(model, L, G, P, pcg);
        if (b < 0) return false;
        if (b > LIMIT) { pcg = 1; b = smem_any(model, L, G, P, pcg); }
        if (b > LIMIT) continue;
        pl->G = G; pl->threads = t; pl->PCG = pcg; pl->smem = b; pl->grid_cap = 0;
        return true;
    }
    return false;
}

}  // namespace

long long eval_smem_bytes(int model, int chain_len, int P, int *images_per_cta, int *threads)
{
    LaunchPlan pl;
    if (!plan_eval(model, chain_len, P, &pl)) return -1;
    if (images_per_cta) *images_per_cta = pl.G;
    if (threads) *threads = pl.threads;
    return pl.smem;
}

cudaError_t launch_eval(int model, int chain_len, const EvalArgs &args, cudaStream_t stream,
                        unsigned long long *launches)
{
    switch (model) {
    case MODEL_EUCM: return launch_model<MODEL_EUCM>(chain_len, args, stream, launches);
    case MODEL_UCM: return launch_model<MODEL_UCM>(chain_len, args, stream, launches);
    case MODEL_MEI: return launch_model<MODEL_MEI>(chain_len, args, stream, launches);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace vg
