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
//   0. pose   : (once per PCG groups) one thread per image accumulates the transform chain
//               with rotation matrices and stages R, t and, per chain element, R12, M12, t13
//               (InterJacobian's members, jacobian.h:139-152).
//   A. corner : one thread per (image, corner): X = R Xb + t, ONE evaluation of the camera
//               model (projection, dP/dX, dP/dintr share rho, eta and their reciprocals),
//               residual and Jacobian rows written into shared memory in exactly the Ceres
//               block layout.  Observations: 16-byte streaming loads issued first.
//   S. store  : the staged blocks of the G images are contiguous in global memory, so one
//               elected thread streams each region out with a TMA bulk copy
//               (cp.async.bulk shared::cta -> global, SASS UBLKCP); no register round trip.
//   B. normal : the per-image block [J r]^T [J r] is a small Gram matrix (2P rows x K+6L+1
//               columns): one warp per image runs it on the FP64 MMA path (mma.sync.m8n8k4.f64,
//               256 FMAs per instruction, fragments loaded straight from the staged rows) by row
//               parity (vg_gram.cuh: the structural zeros of the intrinsic rows are never
//               multiplied: 2 MMA tiles per k-step for EUCM instead of 3, 3 instead of 6 for MEI).
//               The group's packed blocks leave with one more TMA bulk copy; their per-CTA sums
//               (for the shared normal-equation block) stay in registers until the CTA ends.
// The kernel is HBM-write bound by design (224 B written per EUCM corner).  tcgen05/TMEM are
// not applicable (no FP64 there); the only matrix-unit use is the legacy FP64 mma.sync above.
#pragma once
#include <cstdlib>
#include <cstring>
#include "vg_eval.cuh"
#include "vg_math.cuh"

#include <cstdint>
#include <type_traits>
#include <vector>

namespace vg {

#ifdef VG_PHASE_CLOCKS
// developer build only (VG_VARIANT=phase): clock64 cycles per phase, summed over CTAs, as seen by
// lane 0 of warp 0 (row 0: a normal-equation warp) and of the last warp (row 1)
static __device__ unsigned long long g_phase_clocks[2][8];   // per translation unit (no -rdc)
__device__ __forceinline__ long long vg_clock()
{
    long long c;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory");
    return c;
}
#define VG_PC_DECL long long pc_t = vg_clock(); unsigned long long pc_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define VG_PC(i) { const long long pc_n = vg_clock(); pc_acc[i] += (unsigned long long)(pc_n - pc_t); pc_t = pc_n; }
#define VG_PC_FLUSH_ALL if (lane == 0 && (warp == 0 || warp == nw - 1)) for (int i_ = 0; i_ < 8; i_++) atomicAdd(&g_phase_clocks[warp == 0 ? 0 : 1][i_], pc_acc[i_]);
#else
#define VG_PC_DECL
#define VG_PC(i)
#define VG_PC_FLUSH_ALL
#endif

struct LaunchPlan { int G, threads, PCG; long long smem; };
bool plan_eval(int model, int L, int P, LaunchPlan *pl);
// the launch plan's sizes as functions of the board (shared with plan_eval, vg_eval.cu; shared-memory limits may
// still lower them there -- launch_one checks before it picks a kernel with the board compiled in)
__host__ __device__ constexpr int plan_group(int P) { return 4 * P <= 224 ? 4 : (2 * P <= 224 ? 2 : 1); }
__host__ __device__ constexpr int plan_pcg(int pose_doubles, int G)
{
    int pcg = 5632 / (pose_doubles * 8 * G);
    if (pcg * G > 32) pcg = 32 / G;
    return pcg < 1 ? 1 : pcg;
}

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

template <int MODEL, int L> struct Layout {
    static constexpr int K = Camera<MODEL>::K;
    static constexpr int D = K + 6 * L;
    static constexpr int W = D + 1;               // columns of [J r]
    static constexpr int NE = W * (W + 1) / 2;
    static constexpr int NCB = (W + 7) / 8;       // 8-column blocks of the Gram matrix
    static constexpr int POSE = round_up2(12 + 21 * L);
    static constexpr int NPART = (NE + 95) / 96;  // per-thread slots of the per-CTA block sums
    // per-lane output map of the parity Gram fragments: (iu, iv) per fragment element, bytes when they fit
    using map_t = std::conditional_t<(NE <= 127), char2, short2>;
    static constexpr int PNB = (K - 4 + 3 + 6 * L + 7) / 8;                   // 8-column blocks of the parity Gram
    static constexpr int MAPD = (PNB * (PNB + 1) / 2 * 2 * 32 * (int)sizeof(map_t) + 7) / 8;   // doubles (upper bound)
    static constexpr int TAIL = MAPD;
    static constexpr int CAMD = 14;               // intrinsics (<= 10) + Camera::Consts (<= 4) in shared memory
    // doubles of shared memory: poses of PCG groups + staging of one group of G images + packed blocks
    __host__ __device__ static constexpr long long smem_doubles(int G, int P, int PCG)
    {
        // + TAIL: the per-lane output map of the Gram fragments (vg_gram.cuh: GramMap as short2 per element and
        //   lane); the ragged last k-step of the Gram loop also reads (and discards) up to three corners past the
        //   last image, into this area
        // ... + the observations of the CTA's next group (G x P x 2)
        // ... + the camera's parameters and constants (CAMD)
        return 2 + CAMD + (long long)PCG * G * POSE + (long long)G * 2 * P * (1 + K + 6 * L) + 2 + round_up2(G * NE) + TAIL +
               (long long)G * 2 * P;
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
    double *pose, *rs, *Jas, *Jes[L], *Hs, *zero, *obuf, *cam;
    typename LY::map_t *map;
    __device__ Stage(double *base, int G, int P, int PCG)
    {
        zero = base;            base += 2;                 // two zeros for the Gram padding columns
        cam = base;             base += LY::CAMD;          // intrinsics, then the model's per-launch constants
        pose = base;            base += (size_t)PCG * G * LY::POSE;
        rs = base;              base += (size_t)G * 2 * P;
        Jas = base;             base += (size_t)G * 2 * P * LY::K;
        // the chain blocks start 2 (mod 4) doubles after the intrinsic block: Gram fragment loads that mix both
        // (vg_eval_impl.cuh: gram_image) then fall into distinct banks
        base += ((size_t)G * 2 * P * LY::K) % 4 == 0 ? 2 : 0;
#pragma unroll
        for (int e = 0; e < L; e++) { Jes[e] = base; base += (size_t)G * 2 * P * 6; }
        Hs = base;              base += round_up2(G * LY::NE);
        map = reinterpret_cast<typename LY::map_t *>(base);     base += LY::TAIL;
        obuf = base;
    }
};

// D(8x8) += A(8x4) * B(4x8) in fp64 on the matrix unit (SASS: DMMA)
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

}  // namespace
}  // namespace vg
#include "vg_gram.cuh"
namespace vg {
namespace {

// Which Gram formulation an instantiation uses: the row-parity one when it saves at least two MMA tiles per
// k-step over the dense one (MEI: 3 instead of 6), else the dense one (fewer fragment loads per k-step).
#ifndef VG_GRAM_MODE
#define VG_GRAM_MODE -1
#endif
template <int MODEL, int L> __host__ __device__ constexpr bool use_parity_gram()
{
    constexpr int dense = Layout<MODEL, L>::NCB * (Layout<MODEL, L>::NCB + 1) / 2;
    return VG_GRAM_MODE < 0 ? dense - ParityGram<MODEL, L>::NTILE >= 2 : VG_GRAM_MODE > 0;
}

// ---- phase B: Gram matrix of one image's staged rows, one warp ------------------------------
// Column c of [J r]: c < K intrinsic block, then 6 columns per chain element, then the residual.
// Fragment layout of mma.m8n8k4.f64: lane l feeds A[i = l>>2][k = l&3] and B[k = l&3][j = l>>2], so
// for G = R^T R both are R[k0 + (l&3)][8 b + (l>>2)]; it receives C[l>>2][2 (l&3) + {0,1}].
// (Splitting an image's tiles over several warps was measured and lost: the phase is bound by the FP64 pipe of
// the warp's SM sub-partition, 16 cycles per DMMA, and a CTA's images already sit on all four of them.)
template <int MODEL, int L>
__device__ __forceinline__ void gram_image(const Stage<MODEL, L> &st, const int g, const int lane, const int P)
{
    using LY = Layout<MODEL, L>;
    constexpr int K = LY::K, D = LY::D, W = LY::W, NCB = LY::NCB, NTILE = NCB * (NCB + 1) / 2;
    const int kr = lane & 3, ci = lane >> 2;
    // A k-step takes four rows of one parity (the u-rows, then the v-rows, of four consecutive corners): with the
    // rows 16 K or 96 bytes apart the fragment loads of a half-warp then fall into distinct shared-memory banks
    // (rows 4s..4s+3 would put lanes kr = 0 and kr = 3 on the same ones).
    // Per 8-column block: this lane's source pointer (u-row of corner kr), the offset of the v-row below it and
    // the increment per pair of k-steps (8 rows); padding columns read a zero slot with increment 0.
    const double *ptr[NCB];
    int dv[NCB], stride[NCB];
#pragma unroll
    for (int b = 0; b < NCB; b++) {
        const int c = 8 * b + ci;
        if (c < K) { ptr[b] = st.Jas + ((size_t)g * 2 * P + 2 * kr) * K + c; dv[b] = K; stride[b] = 8 * K; }
        else if (c < D) {
            const int e = (c - K) / 6, q = (c - K) - 6 * e;
            const double *base = st.Jes[0];
#pragma unroll
            for (int ee = 1; ee < L; ee++) if (e == ee) base = st.Jes[ee];
            ptr[b] = base + ((size_t)g * 2 * P + 2 * kr) * 6 + q; dv[b] = 6; stride[b] = 48;
        }
        else if (c == D) { ptr[b] = st.rs + (size_t)g * 2 * P + 2 * kr; dv[b] = 1; stride[b] = 8; }
        else { ptr[b] = st.zero; dv[b] = 0; stride[b] = 0; }
    }
    // two accumulator sets (u-rows / v-rows) keep two independent MMA chains per tile
    double acc[2][NTILE][2];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int t = 0; t < NTILE; t++) { acc[h][t][0] = 0.0; acc[h][t][1] = 0.0; }
    auto mma = [&](const int h, const double (&x)[NCB]) {
        int t = 0;
#pragma unroll
        for (int bi = 0; bi < NCB; bi++)
#pragma unroll
            for (int bj = bi; bj < NCB; bj++) { dmma_8x8x4(acc[h][t][0], acc[h][t][1], x[bi], x[bj]); t++; }
    };
    const int rows = 2 * P;
    const int n8 = rows >> 3;                 // pairs of k-steps whose 8 rows all exist
    for (int s = 0; s < n8; s++) {
        double xu[NCB], xv[NCB];
#pragma unroll
        for (int b = 0; b < NCB; b++) {
            xu[b] = ptr[b][0];
            xv[b] = ptr[b][dv[b]];
            ptr[b] += stride[b];
        }
        mma(0, xu);
        mma(1, xv);
    }
    // the last 2, 4 or 6 rows: consecutive rows, those past the image contribute zeros
    const int rem = rows & 7;
    if (rem) {
        double x[NCB];
#pragma unroll
        for (int b = 0; b < NCB; b++) x[b] = kr < rem ? ptr[b][-kr * dv[b]] : 0.0;
        mma(0, x);
        if (rem > 4) {
#pragma unroll
            for (int b = 0; b < NCB; b++) x[b] = 4 + kr < rem ? ptr[b][(4 - kr) * dv[b]] : 0.0;
            mma(1, x);
        }
    }
    double *h = st.Hs + (size_t)g * LY::NE;
    int t = 0;
#pragma unroll
    for (int bi = 0; bi < NCB; bi++)
#pragma unroll
        for (int bj = bi; bj < NCB; bj++) {
            const int i = 8 * bi + ci;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int j = 8 * bj + 2 * kr + q;
                if (i <= j && j < W) h[pk(i, j, W)] = acc[0][t][q] + acc[1][t][q];
            }
            t++;
        }
}

// ---- phase 0: one image's transform chain -> staged pose record ---------------------
template <int L>
__device__ __forceinline__ void chain_pose(const EvalArgs &args, const int img, double *ps)
{
    if (L == 1 && !args.inverse[0]) {
        // one DIRECT element (every monocular dataset): R12 = I, M12 = J_l(r), t13 = t -- no products needed; this is
        // the latency every CTA pays before its first corner phase
        const int sidx = args.seq_index ? args.seq_index[img] : img;
        const double *x = args.xi[0] + (size_t)sidx * args.xi_stride[0];
        const double t0 = x[0], t1 = x[1], t2 = x[2];
        double Re[9], Jl[9];
        rodrigues_and_left_jacobian(x[3], x[4], x[5], Re, Jl);
#pragma unroll
        for (int i = 0; i < 9; i++) { ps[i] = Re[i]; ps[12 + i] = i % 4 == 0 ? 1.0 : 0.0; ps[21 + i] = Jl[i]; }
        ps[9] = t0; ps[10] = t1; ps[11] = t2;
        ps[30] = t0; ps[31] = t1; ps[32] = t2;
        return;
    }
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

// ---- tail: fold the per-CTA block sums into the dataset's sum and (last dataset) the reduced system ----
// Two-level "last CTA done" tree over the rows of cta_partial; every level adds its rows in index order, so the
// result does not depend on which CTA happens to arrive last.  Replaces the separate reduction launch the
// normal-equation build would otherwise need after every evaluation.
// sum of p[r * stride], r in [r0, r1), added in index order; the (L2) loads of 16 rows are in flight together
__device__ __forceinline__ double sum_rows_in_order(const double *p, const int stride, const int r0, const int r1)
{
    double s = 0.0;
    constexpr int ROW_BATCH = 16;
    for (int b = r0; b < r1; b += ROW_BATCH) {
        double v[ROW_BATCH];
#pragma unroll
        for (int i = 0; i < ROW_BATCH; i++) v[i] = b + i < r1 ? __ldcg(p + (size_t)(b + i) * stride) : 0.0;
#pragma unroll
        for (int i = 0; i < ROW_BATCH; i++) s += v[i];
    }
    return s;
}

// ticket with release + acquire ordering at device scope: the CTA barrier before it orders the other threads'
// writes before this thread's release (cumulativity), the one after it their reads after the acquire
__device__ __forceinline__ unsigned int take_ticket(unsigned int *counter)
{
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
}

// dst[e] = sum of src[r * NE + e], r in [r0, r1), for every entry e.  With NE <= blockDim / 2 a thread is an (entry,
// slice of the rows) pair -- half as many dependent batches of L2 loads per thread -- and the slices of an entry are
// added up in slice order through shared memory (sh: blockDim doubles); fixed order either way.
// Returns (threads tid < NE, when 2 NE <= blockDim) the entry the thread stored.
template <int NE>
__device__ __forceinline__ double reduce_rows(const double *src, const int r0, const int r1, double *dst, double *sh)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    double t = 0.0;
    if (2 * NE <= nt) {
        const int S = nt / NE, q = tid / NE, e = tid - q * NE;
        const int per = (r1 - r0 + S - 1) / S;
        if (q < S) {
            const int a = r0 + q * per, b = min(r1, a + per);
            sh[tid] = sum_rows_in_order(src + e, NE, a, b);
        }
        __syncthreads();
        if (tid < NE) {
            for (int k = 0; k < S; k++) t += sh[k * NE + tid];
            dst[tid] = t;
        }
    } else {
        for (int e = tid; e < NE; e += nt) dst[e] = sum_rows_in_order(src + e, NE, r0, r1);
    }
    return t;
}

// ---- the LM loop's decision, by the thread that finished the evaluation (vg_lm_dev.cuh) ---------------------------
// The same tests in the same order as vg_problem_solve's host loop (Ceres' trust-region minimizer: gradient tolerance
// at the current point, iteration / radius limits, step validity, parameter tolerance, rho, radius update, function
// tolerance).  mode 1: first evaluation (records the cost); 2: a candidate's evaluation; limits_pass: only the
// gradient test was due (nothing was evaluated).  ev = [cost, model decrease, |step|^2, |x|^2 over the poses] as the
// evaluation's exchange left them, so = the reduced solve's scalars (SOLVE_OUT).  Out of line: one thread per launch.
__device__ __noinline__ void lm_decide(LmState *st, const double *so, const double *ev, const int mode, const int limits_pass)
{
    // everything the decision reads, in flight together (one round trip to L2 instead of one per branch taken)
    const LmOptions o = st->opt;
    const double cost0 = st->cost, radius0 = st->radius, dec0 = st->decrease_factor;
    const int iter0 = st->iter, inv0 = st->invalid_run, ns0 = st->num_successful, nu0 = st->num_unsuccessful, hcur0 = st->hcur;
    const unsigned long long r = st->records, epoch0 = st->epoch;
    double sv[7], e4[4];
#pragma unroll
    for (int i = 0; i < 7; i++) sv[i] = (mode == 2) ? __ldcg(so + i) : 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) e4[i] = __ldcg(ev + i);

    double cost = cost0, radius = radius0, dec = dec0, new_cost = 0.0, rho = 0.0, step_norm = 0.0, gmax = 0.0;
    int done = 0, accepted = 0, valid = 0, iter = iter0, inv = inv0, ns = ns0, nu = nu0;
    if (mode == 1) {
        cost = e4[0];
    } else {
        gmax = sv[0];
        if (gmax <= o.gradient_tolerance) done = 1 + 1;
        else if (iter >= o.max_num_iterations) done = 1 + 3;
        else if (radius < o.min_radius) done = 1 + 4;
        else {
            iter++;
            bool ok = sv[1] == 0.0 && sv[2] != 0.0;      // every pose block and the reduced system factorised
            double model_change = 0.0, step2 = 0.0, x2 = 0.0;
            new_cost = e4[0];
            if (ok) {
                model_change = -(sv[3] + 0.5 * sv[4] + e4[1]);
                step2 = sv[5] + e4[2];
                x2 = sv[6] + e4[3];
                if (!(model_change > 0.0)) ok = false;
            }
            step_norm = sqrt(step2);
            if (!ok) {
                // invalid step (linear solve failed or the model predicts no decrease)
                nu++;
                if (++inv >= o.max_consecutive_invalid) done = 1 + 5;
                else { radius /= dec; dec *= 2.0; }
            } else {
                valid = 1;
                inv = 0;
                if (step_norm <= o.parameter_tolerance * (sqrt(x2) + o.parameter_tolerance)) {
                    done = 1 + 2;             // candidate discarded, as Ceres stops before taking the step
                } else {
                    rho = (cost - new_cost) / model_change;
                    if (rho > o.min_relative_decrease) {
                        const double cost_change = cost - new_cost, old_cost = cost, t = 2.0 * rho - 1.0;
                        accepted = 1;
                        cost = new_cost;
                        ns++;
                        radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
                        if (radius > o.max_radius) radius = o.max_radius;
                        dec = 2.0;
                        if (fabs(cost_change) <= o.function_tolerance * old_cost) done = 1 + 0;
                    } else {
                        nu++;
                        radius /= dec; dec *= 2.0;
                    }
                }
            }
        }
    }
    const int hcur = hcur0 ^ accepted;        // the next factorisation reads the packed blocks where this evaluation left them
    const unsigned long long epoch = epoch0 + (mode == 2 ? (limits_pass ? 1 : 2) : 0);    // (several ranks: exchanges used)
    st->done = done;
    st->limits = !done && (iter >= o.max_num_iterations || radius < o.min_radius);
    st->migrate = accepted;                   // ... and copies the candidate's poses and slab over
    st->hcur = hcur;
    if (mode == 2) st->init_scale = 0;
    st->iter = iter; st->invalid_run = inv; st->num_successful = ns; st->num_unsuccessful = nu;
    st->radius = radius; st->decrease_factor = dec; st->cost = cost;
    st->epoch = epoch;
    st->records = r + 1;
    LmRecord *rec = &st->rec;
    rec->seq = r + 1;
    rec->iter = iter; rec->done = done; rec->accepted = accepted; rec->valid = valid;
    rec->num_successful = ns; rec->num_unsuccessful = nu; rec->migrate = accepted; rec->hcur = hcur; rec->epoch = epoch;
    rec->cost = cost; rec->radius = radius; rec->prev_cost = cost0; rec->prev_radius = radius0;
    rec->new_cost = new_cost; rec->rho = rho; rec->step_norm = step_norm; rec->gmax = gmax;
}

// The three sums over the poses the decision needs (model decrease, |step|^2, |x|^2): one row per block of the
// back-substitution kernel, added up here -- by one CTA at the head of the candidate's evaluation, under everybody
// else's work -- instead of by that kernel's last block on the way to this launch.  Rows lane, lane + 32, ... per lane,
// lanes in order: the order fast_backsub's own tail uses.
__device__ __noinline__ void lm_model_sums(const double *partial, const int rows, double *dst)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= 3) return;
    constexpr int B = 24;
    double s = 0.0;
    for (int b = lane; b < rows; b += B * 32) {
        double v[B];
#pragma unroll
        for (int i = 0; i < B; i++) v[i] = b + i * 32 < rows ? __ldcg(partial + (size_t)(b + i * 32) * 3 + q) : 0.0;
#pragma unroll
        for (int i = 0; i < B; i++) s += v[i];
    }
    double tot = 0.0;
    for (int l = 0; l < 32; l++) tot += __shfl_sync(0xffffffffu, s, l);
    if (lane == 0) dst[q] = tot;
}

// The deferred exchange this launch owes (one CTA, at its head): post this rank's block if the launch that produced it
// left that to us, form the sum, publish "collected".  Out of line: one CTA in one launch out of many runs it, and
// inlined it costs the kernel's main loop registers.
__device__ __noinline__ void head_exchange(double *buf, int count, PeerCtx pc, int post, unsigned long long *done,
                                           double *scratch, int cap)
{
    if (post) peer_post(buf, count, pc);
    peer_collect(buf, count, pc, scratch, cap);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(done), "l"(pc.epoch) : "memory");
}

template <int NE, bool LMD>
__device__ __forceinline__ void fused_reduce(const EvalArgs &args, double *scratch, const int scratch_cap)
{
    __shared__ int s_last;
    const int tid = threadIdx.x, rows = gridDim.x;
    const int ngrp = (rows + EVAL_REDUCE_GROUP - 1) / EVAL_REDUCE_GROUP, grp = blockIdx.x / EVAL_REDUCE_GROUP;
    const int r0 = grp * EVAL_REDUCE_GROUP, r1 = min(rows, r0 + EVAL_REDUCE_GROUP);
    __syncthreads();                       // the CTA's row of cta_partial is written; the staging area (scratch) is free
    if (tid == 0) s_last = take_ticket(args.tickets + 1 + grp) == (unsigned)(r1 - r0 - 1);
    __syncthreads();
    if (!s_last) return;
    reduce_rows<NE>(args.cta_partial, r0, r1, args.lvl1 + (size_t)grp * NE, scratch);
    __syncthreads();
    if (tid == 0) {
        args.tickets[1 + grp] = 0;
        s_last = take_ticket(args.tickets) == (unsigned)(ngrp - 1);
    }
    __syncthreads();
    if (!s_last) return;
    // one dataset feeding the reduced system alone: entry e of its sums goes straight to its place (emap), without the
    // round trips of the table-driven assembly below
    const bool direct = args.red && args.emap && 2 * NE <= (int)blockDim.x;
    EMapEntry em{-1, -1, 0.0};
    if (direct && tid < NE) em = args.emap[tid];
    const double mine = reduce_rows<NE>(args.lvl1, 0, ngrp, args.ds_sum, scratch);
    if (tid == 0) args.tickets[0] = 0;
    if (!args.red) return;
    if (args.peer.n > 1 && args.peer_deferred && args.collect.n > 1) {
        // the previous exchange must be collected (head of this launch) before its buffer is reused and before the
        // next one is posted
        if (tid == 0) {
            unsigned long long seen;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(args.collect_done) : "memory");
            } while (seen < args.collect.epoch);
        }
        __syncthreads();
    }
    if (direct) {
        if (em.dst0 >= 0) {
            const double v = mine * em.scale;
            args.red[em.dst0] = v;
            if (em.dst1 >= 0) args.red[em.dst1] = v;
        }
    } else {
    __threadfence();
    __syncthreads();                       // this dataset's sums are complete; the earlier datasets' launches are
    for (int o = tid; o < args.n_fin_out; o += blockDim.x) {
        const FinOut f = args.fin_outs[o];
        double v = 0.0;
        for (int si = f.src_begin; si < f.src_end; si++) {
            const FinSrc sr = args.fin_srcs[si];
            v += __ldcg(args.fin_base + (size_t)sr.off + sr.e);
        }
        v *= f.scale;
        args.red[f.dst0] = v;
        if (f.dst1 >= 0) args.red[f.dst1] = v;
    }
    }
    // several GPUs: the same CTA exchanges the reduced system with its peers over NVLink (vg_peer.cuh)
    if (args.peer.n > 1) {
        if (args.peer_deferred == 1) peer_post(args.red, args.peer_count, args.peer);
        else if (!args.peer_deferred) {
            PeerCtx pc = args.peer;
            // (the LM loop on the device numbers its exchanges itself: launches queued past the end of a solve use none)
            if (LMD && args.lm_mode == 2) pc.epoch = *reinterpret_cast<volatile unsigned long long *>(&args.lm->epoch) + 1;
            peer_allreduce(args.red, args.peer_count, pc, scratch, scratch_cap);
        }
        // (peer_deferred == 2: the block stays in args.red; this problem's next launch posts it from its head, so that
        // not even the remote stores' round trip -- which the END of a grid has to wait for -- is paid by a step)
    }
    if (args.host_flag) {
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < args.host_count; i++)
                args.host_value[i] = *reinterpret_cast<volatile double *>(args.red + args.host_index + i);
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(args.host_flag), "l"(args.host_seq) : "memory");
        }
    }
    if (LMD && args.lm_mode) {
        __threadfence();
        __syncthreads();
#ifdef VG_LM_STAMPS
        if (args.lm_mode == 2) { VG_LM_STAMP(args.lm, 2, 1) }
#endif
        if (tid == 0) lm_decide(args.lm, args.lm_so, args.red + args.host_index, args.lm_mode, 0);
    }
}

// ---- phase A stores ---------------------------------------------------------------------
// A corner's two rows (u, v) of a block are 2 N contiguous doubles, the corners of a warp 2 N doubles apart.  Written
// row by row with 16-byte stores, lanes l and l + 4 of a quarter-warp would hit the same banks when N / 2 is odd
// (96-byte and 160-byte corner strides: only the even 16-byte bank groups are used).  The lanes with bit 2 set
// therefore store their v-row first and their u-row second: the quarter-warp then covers all eight bank groups in
// every store instruction.  N odd (UCM): the 2 N doubles go out as N 16-byte stores across the row boundary (corner
// stride N x 16 bytes, N odd: conflict free as it is).
template <int N>
__device__ __forceinline__ void store_pair(double *p, const double (&a)[N], const double (&b)[N], const bool hi)
{
    if constexpr ((N & 1) == 0) {
        double *pf = p + (hi ? N : 0), *ps = p + (hi ? 0 : N);
#pragma unroll
        for (int i = 0; i < N; i += 2)
            *reinterpret_cast<double2 *>(pf + i) = make_double2(hi ? b[i] : a[i], hi ? b[i + 1] : a[i + 1]);
#pragma unroll
        for (int i = 0; i < N; i += 2)
            *reinterpret_cast<double2 *>(ps + i) = make_double2(hi ? a[i] : b[i], hi ? a[i + 1] : b[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i + 1 < N; i += 2) *reinterpret_cast<double2 *>(p + i) = make_double2(a[i], a[i + 1]);
        *reinterpret_cast<double2 *>(p + N - 1) = make_double2(a[N - 1], b[0]);
#pragma unroll
        for (int i = 1; i + 1 < N; i += 2) *reinterpret_cast<double2 *>(p + N + i) = make_double2(b[i], b[i + 1]);
    }
}

// ---- the kernel -------------------------------------------------------------------
// CTA b owns the image groups b, b + gridDim.x, ... (static, hence deterministic); a group is G consecutive images,
// PCG groups share one batch of staged poses (PCG * G <= 32: one warp restages them).  Per group, two CTA barriers:
//   A   corner phase: camera model, residual / Jacobian rows -> staging area
//   B2
//   S   the last warp: TMA bulk stores of the staged blocks; the poses of the next batch when this group is the last of
//       its batch; then it waits until the copies have read the staging area -- all of it while the other warps are
//       in the Gram phase, so nobody waits for these
//   B   one warp per image: Gram matrix on the FP64 MMA path
//   B3
//   H   packed blocks: TMA bulk store (its read is awaited before the next B2), per-CTA running sums
// (Measured and rejected on the way, see DESIGN.md: contiguous image ranges per CTA, an mbarrier in place of a
// barrier for "staging free", an image's Gram tiles spread over several warps.)
#ifndef VG_WIDE_BLOCKS
#define VG_WIDE_BLOCKS 2
#endif
// developer switches (A/B builds through VG_EXTRA_FLAGS; the defaults are the product)
#ifndef VG_SWAP_STORES
#define VG_SWAP_STORES 1      // store_pair's bank-conflict-free order
#endif
#ifndef VG_CAM_SMEM
#define VG_CAM_SMEM 1         // camera parameters read from shared memory in the corner phase (0: kept in registers)
#endif
// PC > 0: the board's point count is compiled in (with the images per group and the pose batch the launch plan picks
// for it: plan_group / plan_pcg), so that the index arithmetic of the corner phase and every shared-memory offset
// are constants; PC == 0: any board, sizes at run time.
// LMD: the instantiation that takes part in the LM loop whose control state lives on the device (EvalArgs::lm_mode, chains
// of one transform only); the ordinary one carries none of that code.
template <int MODEL, int L, int PC, bool LMD = false>
__global__ void __launch_bounds__(224, (L == 1 && Camera<MODEL>::K <= 6) ? 4 : (L == 1 ? VG_WIDE_BLOCKS : 2))
reproj_eval_kernel(const EvalArgs args, const int G_rt, const int PCG_rt)
{
    using LY = Layout<MODEL, L>;
    using CAM = Camera<MODEL>;
    constexpr int K = LY::K;
    extern __shared__ __align__(16) double smem[];
    const int P = PC ? PC : args.P;
    const int G = PC ? plan_group(PC) : G_rt;
    const int PCG = PC ? plan_pcg(LY::POSE, plan_group(PC)) : PCG_rt;
    const Stage<MODEL, L> st(smem, G, P, PCG);
    const int tid = threadIdx.x;
    // the warp index through a shuffle: provably warp-uniform (uniform datapath for the TMA operands)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31, nw = blockDim.x >> 5;
    const int T0 = (nw - 1) * 32;              // the one thread that issues (and waits for) every TMA copy
    const int ngw = nw - 1;                    // warps 0 .. ngw-1 take the Gram phase (nw >= 3: launch plan)
    const bool hi = VG_SWAP_STORES && (tid & 4) != 0;            // store_pair: this lane writes its v-row first
    if (tid < 2) st.zero[tid] = 0.0;
    // the output map of the Gram fragments is a function of the lane only: the first warp leaves it in shared
    // memory (read back per image, so that it does not occupy registers during the corner phase)
    if constexpr (use_parity_gram<MODEL, L>()) {
        if (tid < 32) {
            GramMap<MODEL, L> gm;
            gram_map_init<MODEL, L>(gm, lane);
#pragma unroll
            for (int t = 0; t < ParityGram<MODEL, L>::NTILE; t++)
#pragma unroll
                for (int q = 0; q < 2; q++)
                    st.map[(t * 2 + q) * 32 + lane] = typename LY::map_t{(decltype(LY::map_t::x))gm.iu[t][q], (decltype(LY::map_t::x))gm.iv[t][q]};
        }
    }     // visible to everyone after the first __syncthreads below

    // this CTA's groups
    // (several ranks: the last CTA takes no groups -- it is the one that exchanges the reduced system with the peers, at
    // its head, while the others compute; the same partition in every launch of such a problem, so that its sums do not
    // depend on the exchange mode)
    const int n_groups_all = (args.n_img + G - 1) / G;
    const int work_ctas = args.peer.n > 1 && gridDim.x > 1 ? (int)gridDim.x - 1 : (int)gridDim.x;
    const int my_groups = (int)blockIdx.x < work_ctas && n_groups_all > (int)blockIdx.x
                              ? (n_groups_all - 1 - (int)blockIdx.x) / work_ctas + 1 : 0;
    auto group_first = [&](const int gi) { return (int)(((long long)blockIdx.x + (long long)gi * work_ctas) * G); };
    // poses of a batch of PCG groups starting at group gi0: thread t takes image (t % G) of group gi0 + t / G
    auto stage_poses = [&](const int gi0, const int t) {
        const int j = t / G, gq = t - j * G;
        if (j < PCG && gi0 + j < my_groups) {
            const int img = group_first(gi0 + j) + gq;
            if (img < args.n_img) chain_pose<L>(args, img, st.pose + (size_t)t * LY::POSE);
        }
    };

    // Programmatic dependent launch: the next kernel of the stream may take this CTA's SM slot as soon as it is free.
    // Which grids can be ahead of this one and still running?  Only kernels that release their dependents early: other
    // evaluation launches, and the LM step's kernels inside vg_problem_solve -- where the caller asks every evaluation to
    // wait at its head (EvalArgs::wait_at_head); behind any other kernel this launch starts when that one has completed.
    // An evaluation writes residuals, Jacobians, per-image blocks and, in its tail, the reduction's scratch and result:
    // never what another evaluation READS in its main loop (observations, board, camera, poses).  So the main loop
    // need not wait for the launch ahead: it runs under that launch's stragglers (2 500 groups over 592 CTAs are 4.2
    // rounds: a fifth of the SM slots idle through the last one) and under its reduction tail, and waits only before
    // it touches the reduction's scratch itself.  (Two launches on the same buffers write the same bytes.)  The exception
    // that wait at the head: the LM loop on the device (its state and the poses come from the kernels ahead) and the CTA
    // that collects a deferred peer exchange (it reads the previous launch's result; it has no groups of its own).
    // VG_LATE_WAIT=0: always at the head.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const bool wait_at_head = LMD || !args.late_wait || (args.collect.n > 1 && blockIdx.x == gridDim.x - 1);
    if (wait_at_head) asm volatile("griddepcontrol.wait;" ::: "memory");

    // the LM loop on the device (vg_lm_dev.cuh).  A candidate's evaluation (mode 2): is the solve over (this launch was queued
    // ahead of the decision), or is only the gradient test due?  The two words are in flight while the prologue below
    // reads its inputs; they are looked at before anything is written.  The first evaluation (mode 1): block 0 brings the
    // loop's initial state over from host-mapped memory.
    __shared__ int s_lm_skip;                  // (bit 0: done, bit 1: limits; through shared memory: no register lives that long)
    if (LMD && args.lm_mode == 2 && tid == 33) {
        const int dn = *reinterpret_cast<volatile int *>(&args.lm->done), lim = *reinterpret_cast<volatile int *>(&args.lm->limits);
        s_lm_skip = (dn ? 1 : 0) | (lim ? 2 : 0);
    }
    // (the pose sums of the step, by the CTA with the fewest groups: before the prologue, while few registers are live
    // across the call; harmless when the solve turns out to be over)
    if (LMD && args.lm_mode == 2 && blockIdx.x == gridDim.x - 1)
        lm_model_sums(args.lm_partial, args.lm_partial_rows, args.red + args.host_index + 1);
    if (LMD && args.lm_mode == 1 && blockIdx.x == 0 && tid >= 64 && tid < 64 + (int)(sizeof(LmState) / 8)) {
        const unsigned long long v = *(reinterpret_cast<const volatile unsigned long long *>(args.lm_init) + (tid - 64));
        reinterpret_cast<unsigned long long *>(args.lm)[tid - 64] = v;
    }

    // several GPUs, deferred exchange: one CTA (the last: it has the fewest groups) forms the sum of this problem's
    // previous exchange, which has been crossing NVLink while the launches in between ran (vg_peer.cuh)
    if (args.collect.n > 1 && blockIdx.x == gridDim.x - 1)
        head_exchange(args.collect_buf, args.peer_count, args.collect, args.collect_post, args.collect_done, st.rs,
                      G * 2 * P * (1 + K + 6 * L));

    // the camera's parameters and constants wait in shared memory: read where the corner phase needs them, they do
    // not occupy registers during the Gram phase
#if VG_CAM_SMEM
    if (tid == 32) {        // (not thread 0: that one stages a pose at the same time)
        double intr[K];
#pragma unroll
        for (int i = 0; i < K; i++) { intr[i] = __ldg(args.intr + i); st.cam[i] = intr[i]; }
        CAM::save(CAM::prepare(intr), st.cam + K);
    }
#else
    double intr[K];
#pragma unroll
    for (int i = 0; i < K; i++) intr[i] = __ldg(args.intr + i);
    const typename CAM::Consts cc = CAM::prepare(intr);
#endif
    const bool first_direct = (args.inverse[0] == 0);   // R12 of element 0 is the identity
    double part[LY::NPART];                              // this CTA's sum of its images' blocks
#pragma unroll
    for (int q = 0; q < LY::NPART; q++) part[q] = 0.0;

    // Observations are fetched one group ahead into shared memory (cp.async, every thread the element it will
    // itself consume): under the kernel's own store stream a global read takes longer than a corner pass.
    auto fetch_obs = [&](const int gi) {
        if (gi < my_groups) {
            const int i0 = group_first(gi), n = min(G, args.n_img - i0) * P;
            for (int idx = tid; idx < n; idx += blockDim.x)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(st.obuf + 2 * idx)),
                             "l"(args.obs + ((size_t)i0 * P + idx) * 2) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_obs(0);
    // poses of the first PCG groups: one thread per image
    if (tid < PCG * G) stage_poses(0, tid);
    __syncthreads();
    if (LMD && args.lm_mode == 2) {
        const int skip = s_lm_skip;
        if (skip) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            if (!(skip & 1) && blockIdx.x == 0 && tid == 0) lm_decide(args.lm, args.lm_so, args.red + args.host_index, 2, 1);
            return;
        }
        VG_LM_STAMP(args.lm, 2, 0)
    }
    VG_PC_DECL

    // ---- A: one thread per (image, corner) -- the camera model, then the rows into the staging area -----------
    // When a group fits one pass over the CTA (always, unless the board has more points than the CTA threads) a
    // thread keeps its (image, corner) pair and its board point for the whole launch.
    const bool single_pass = G * P <= (int)blockDim.x;
    const int g_t = tid / P, c_t = tid - g_t * P;
    double bx_t = 0.0, by_t = 0.0, bz_t = 0.0;
    if (single_pass && tid < G * P) {
        bx_t = __ldg(args.board + 3 * c_t); by_t = __ldg(args.board + 3 * c_t + 1); bz_t = __ldg(args.board + 3 * c_t + 2);
    }
    const double *pose_base = st.pose;          // the current group's pose records
    auto corner = [&](const int idx, const int g, const int c, const double bx, const double by, const double bz) {
        const double *ps = pose_base + (size_t)g * LY::POSE;
#if VG_CAM_SMEM
        double intr[K];
#pragma unroll
        for (int i = 0; i < K; i++) intr[i] = st.cam[i];
        const typename CAM::Consts cc = CAM::load(intr, st.cam + K);
#endif
        const double2 ob = *reinterpret_cast<const double2 *>(st.obuf + 2 * idx);
        const double X0 = fma(ps[2], bz, fma(ps[1], by, fma(ps[0], bx, ps[9])));
        const double X1 = fma(ps[5], bz, fma(ps[4], by, fma(ps[3], bx, ps[10])));
        const double X2 = fma(ps[8], bz, fma(ps[7], by, fma(ps[6], bx, ps[11])));
        double u, v, Pu[3], Pv[3], Ju[K], Jv[K];
        const bool ok = CAM::eval(intr, cc, X0, X1, X2, u, v, Pu, Pv, Ju, Jv);
        // rows into the staging area, in exactly the Ceres block layout
        const size_t row = (size_t)g * 2 * P + 2 * c;
        *reinterpret_cast<double2 *>(st.rs + row) = make_double2(u - ob.x, v - ob.y);
        store_pair<K>(st.Jas + row * K, Ju, Jv, hi);
        // rows of dP/dX in the order this lane stores them (see store_pair)
        double Pa[3], Pb[3];
#pragma unroll
        for (int q = 0; q < 3; q++) { Pa[q] = hi ? Pv[q] : Pu[q]; Pb[q] = hi ? Pu[q] : Pv[q]; }
        const int ra = hi ? 6 : 0, rb = hi ? 0 : 6;
#pragma unroll
        for (int e = 0; e < L; e++) {
            const double *pe = ps + 12 + 21 * e;
            const double w0 = X0 - pe[18], w1 = X1 - pe[19], w2 = X2 - pe[20];
            // (w x p)^T M12  ==  -p^T hat(w) M12   (jacobian.h:165,170)
            const double ca0 = w1 * Pa[2] - w2 * Pa[1], ca1 = w2 * Pa[0] - w0 * Pa[2],
                         ca2 = w0 * Pa[1] - w1 * Pa[0];
            const double cb0 = w1 * Pb[2] - w2 * Pb[1], cb1 = w2 * Pb[0] - w0 * Pb[2],
                         cb2 = w0 * Pb[1] - w1 * Pb[0];
            double ja[6], jb_[6];
            if (e == 0 && first_direct) {
#pragma unroll
                for (int q = 0; q < 3; q++) { ja[q] = Pa[q]; jb_[q] = Pb[q]; }
            } else {
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    ja[q] = fma(Pa[2], pe[6 + q], fma(Pa[1], pe[3 + q], Pa[0] * pe[q]));
                    jb_[q] = fma(Pb[2], pe[6 + q], fma(Pb[1], pe[3 + q], Pb[0] * pe[q]));
                }
            }
#pragma unroll
            for (int q = 0; q < 3; q++) {
                ja[3 + q] = fma(ca2, pe[15 + q], fma(ca1, pe[12 + q], ca0 * pe[9 + q]));
                jb_[3 + q] = fma(cb2, pe[15 + q], fma(cb1, pe[12 + q], cb0 * pe[9 + q]));
            }
            double *pj = st.Jes[e] + row * 6;
            store_row<6>(pj + ra, ja, true);
            store_row<6>(pj + rb, jb_, true);
        }
        if (!ok) {
            // failed projection (rare): the 1e15 sentinel and zero Jacobian rows over what was just written
            // (calib_cost_functions.cpp:66-70, eucm.h:141-150,198-206)
            *reinterpret_cast<double2 *>(st.rs + row) = make_double2(DOUBLE_BIG, DOUBLE_BIG);
            for (int i = 0; i < 2 * K; i++) st.Jas[row * K + i] = 0.0;
#pragma unroll
            for (int e = 0; e < L; e++)
                for (int i = 0; i < 12; i++) st.Jes[e][row * 6 + i] = 0.0;
        }
    };

    for (int gi = 0; gi < my_groups; gi++) {
        const int jb = gi % PCG;
        const int img0 = group_first(gi);
        const int nv = min(G, args.n_img - img0);
        const int ncorn = nv * P;
        pose_base = st.pose + (size_t)jb * G * LY::POSE;

        asm volatile("cp.async.wait_group 0;" ::: "memory");     // this thread's own observations have landed
        if (single_pass) {
            if (tid < ncorn) corner(tid, g_t, c_t, bx_t, by_t, bz_t);
        } else {
            for (int idx = tid; idx < ncorn; idx += blockDim.x) {
                const int g = idx / P, c = idx - g * P;
                corner(idx, g, c, __ldg(args.board + 3 * c), __ldg(args.board + 3 * c + 1), __ldg(args.board + 3 * c + 2));
            }
        }
        VG_PC(0)
        // the previous group's packed blocks must have been read before this group's Gram phase rewrites them
        if (tid == T0) bulk_wait_read_all();
        fetch_obs(gi + 1);          // the next group's observations, into the elements just consumed
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA engine
        VG_PC(1)
        __syncthreads();            // ---- B2
        VG_PC(2)

        const bool restage = jb == PCG - 1 && gi + 1 < my_groups;
        if (warp == nw - 1) {
            // ---- S: the Ceres-layout blocks leave with TMA bulk copies; next batch of poses; staging area read ----
            if (lane == 0) {
                const size_t rows = (size_t)nv * 2 * P;
                bool issued = false;
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
            // every warp has left the corner phase of this batch's last group: its pose records are free
            if (restage) stage_poses(gi + 1, lane);
            if (lane == 0) bulk_wait_read_all();      // the next corner phase may overwrite the staging area
        }
        // ---- B: per-image normal-equation blocks, one warp per image -----------------------------------
        if (args.H) {
            for (int g = warp; g < nv && warp < ngw; g += ngw) {
                if constexpr (use_parity_gram<MODEL, L>()) {
                    const double *Jes_g[L];
#pragma unroll
                    for (int e = 0; e < L; e++) Jes_g[e] = st.Jes[e] + (size_t)g * 2 * P * 6;
                    GramFrag<MODEL, L> v;
                    gram_slot_parity<MODEL, L>(st.rs + (size_t)g * 2 * P, st.Jas + (size_t)g * 2 * P * K, Jes_g, st.zero, lane, P, v);
                    gram_frag_emit<MODEL, L, typename LY::map_t>(v, st.map, st.Hs + (size_t)g * LY::NE, lane);
                } else {
                    gram_image<MODEL, L>(st, g, lane, P);
                }
                if (args.loss_b > 0.0) {       // SoftLOneLoss on this image's block (see EvalArgs::loss_b)
                    double *Hg = st.Hs + (size_t)g * LY::NE;
                    __syncwarp();
                    const double q = sqrt(1.0 + Hg[LY::NE - 1] / args.loss_b);
                    const double w = 1.0 / q;
                    __syncwarp();
                    for (int e = lane; e < LY::NE - 1; e += 32) Hg[e] *= w;
                    if (lane == 0) Hg[LY::NE - 1] = 2.0 * args.loss_b * (q - 1.0);
                }
            }
            fence_proxy_async_smem();
        }
        VG_PC(3)
        __syncthreads();            // ---- B3: packed blocks complete, staging area read, next poses staged
        VG_PC(4)
        if (args.H) {
            // the group's packed blocks are contiguous in global memory: one more bulk copy (a ragged last
            // group, whose byte count may not be a multiple of 16, goes through ordinary stores)
            double *Hb = args.H;
            if (LMD && args.lm_mode == 2) {        // candidates alternate between the two buffers (LmState::hcur)
                int hc;
                asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(hc) : "l"(&args.lm->hcur));
                if (hc) Hb = args.H_alt;
            }
            double *Hg = Hb + (size_t)img0 * LY::NE;
            const size_t hbytes = (size_t)nv * LY::NE * 8;
            const bool h_bulk = (hbytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(Hg) & 15) == 0);
            if (h_bulk && tid == T0) { bulk_store(Hg, st.Hs, (uint32_t)hbytes); bulk_commit(); }
#pragma unroll
            for (int q = 0; q < LY::NPART; q++) {
                const int e = tid + q * blockDim.x;
                if (e < LY::NE) {
                    double sum = 0.0;
                    for (int g = 0; g < nv; g++) {
                        const double val = st.Hs[(size_t)g * LY::NE + e];
                        if (!h_bulk) Hg[(size_t)g * LY::NE + e] = val;
                        sum += val;
                    }
                    part[q] += sum;
                }
            }
        }
        VG_PC(5)
    }
    if (tid == T0) bulk_wait_read_all();        // shared memory must outlive the TMA reads
    VG_PC_FLUSH_ALL
    // (late wait, see the head: from here on the launch ahead must be over -- the reduction's scratch is shared with it,
    // and a grid must not complete before the grid ahead of it in the stream has)
    if (!wait_at_head) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (args.H && args.cta_partial) {
#pragma unroll
        for (int q = 0; q < LY::NPART; q++) {
            const int e = tid + q * blockDim.x;
            if (e < LY::NE) args.cta_partial[(size_t)blockIdx.x * LY::NE + e] = part[q];
        }
        if (args.tickets) fused_reduce<LY::NE, LMD>(args, st.rs, G * 2 * P * (1 + K + 6 * L));
    }
}

template <int MODEL, int L, int PC, bool LMD = false>
cudaError_t launch_fixed(const EvalArgs &args, const LaunchPlan &pl, cudaStream_t stream, unsigned long long *launches,
                         int *grid_out, bool query_only)
{
    static int configured_bytes[64];    // per instantiation and device; zero-initialised
    static int blocks_per_sm[64];
    static int sm_count[64];
    static int planned_threads[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (pl.smem > configured_bytes[dev] || planned_threads[dev] != pl.threads) {
        cudaError_t e = cudaFuncSetAttribute(reproj_eval_kernel<MODEL, L, PC, LMD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return e;
        configured_bytes[dev] = (int)pl.smem;
        planned_threads[dev] = pl.threads;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[dev], reproj_eval_kernel<MODEL, L, PC, LMD>,
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
    // launched with programmatic stream serialization (developer knob VG_PDL=0: plain launch)
    static const bool pdl = [] { const char *e = getenv("VG_PDL"); return !(e && e[0] == '0'); }();
    static const bool late = [] { const char *e = getenv("VG_LATE_WAIT"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(pl.threads); cfg.dynamicSmemBytes = (size_t)pl.smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    EvalArgs launch_args = args;
    launch_args.late_wait = pdl && late && !LMD && !args.wait_at_head ? 1 : 0;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, reproj_eval_kernel<MODEL, L, PC, LMD>, launch_args, pl.G, pl.PCG);
    if (launches) count_launch(launches);
    return le != cudaSuccess ? le : cudaGetLastError();
}

// boards with a kernel of their own (the 9 x 6 board of every BASELINE configuration), for chains of one or two
// transforms; anything else runs the run-time-sized kernel
template <int MODEL, int L>
cudaError_t launch_one(const EvalArgs &args, cudaStream_t stream, unsigned long long *launches, int *grid_out,
                       bool query_only)
{
    LaunchPlan pl;
    if (!plan_eval(MODEL, L, args.P, &pl)) return cudaErrorInvalidValue;
#ifndef VG_NO_FIXED_BOARD
    if constexpr (L <= 2) {
        constexpr int PC = 54;
        if (args.P == PC && pl.G == plan_group(PC) && pl.PCG == plan_pcg(Layout<MODEL, L>::POSE, plan_group(PC))) {
            if constexpr (L == 1)
                if (args.lm_mode) return launch_fixed<MODEL, L, PC, true>(args, pl, stream, launches, grid_out, query_only);
            return launch_fixed<MODEL, L, PC>(args, pl, stream, launches, grid_out, query_only);
        }
    }
#endif
    if (args.lm_mode) {
        if constexpr (L == 1) return launch_fixed<MODEL, L, 0, true>(args, pl, stream, launches, grid_out, query_only);
        return cudaErrorInvalidValue;
    }
    return launch_fixed<MODEL, L, 0>(args, pl, stream, launches, grid_out, query_only);
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
