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

constexpr int round_up2(int x) { return (x + 1) & ~1; }

template <int MODEL, int L> struct Layout {
    static constexpr int K = Camera<MODEL>::K;
    static constexpr int NBI = (K > 6) ? 2 : 1;   // intrinsic sub-blocks of <= 6 columns
    static constexpr int KA = K / NBI;
    static constexpr int NB = NBI + L;            // column blocks (residual column rides along)
    static constexpr int D = K + 6 * L;
    static constexpr int W = D + 1;
    static constexpr int NE = W * (W + 1) / 2;
    static constexpr int NT = NB * (NB + 1) / 2;  // block-pair tiles
    static constexpr int POSE = round_up2(12 + 21 * L);
    // doubles of shared memory for G images of P corners
    __host__ __device__ static constexpr long long smem_doubles(int G, int P)
    {
        return (long long)G * POSE + (long long)G * 2 * P * (1 + K + 6 * L) + (long long)G * NE;
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
    __device__ Stage(double *base, int G, int P)
    {
        pose = base;            base += (size_t)G * LY::POSE;
        rs = base;              base += (size_t)G * 2 * P;
        Jas = base;             base += (size_t)G * 2 * P * LY::K;
#pragma unroll
        for (int e = 0; e < L; e++) { Jes[e] = base; base += (size_t)G * 2 * P * 6; }
        Hs = base;
    }
};

// ---- phase B: one block-pair tile -------------------------------------------------
template <int MODEL, int L, int T>
__device__ __forceinline__ void run_tile(const Stage<MODEL, L> &st, const int g, const int s, const int S,
                                         const int P, const bool valid)
{
    using LY = Layout<MODEL, L>;
    // decode T -> (bi,bj), bi <= bj, enumerated row by row
    constexpr int NB = LY::NB;
    constexpr int bi = [] { int t = T, i = 0; while (t >= NB - i) { t -= NB - i; i++; } return i; }();
    constexpr int bj = [] { int t = T, i = 0; while (t >= NB - i) { t -= NB - i; i++; } return i + t; }();
    constexpr bool DIAG = (bi == bj);
    constexpr bool LAST = DIAG && (bi == NB - 1);
    constexpr bool I_INTR = bi < LY::NBI, J_INTR = bj < LY::NBI;
    constexpr int NI = I_INTR ? LY::KA : 6, NJ = J_INTR ? LY::KA : 6;
    constexpr int RSI = I_INTR ? LY::K : 6, RSJ = J_INTR ? LY::K : 6;
    constexpr int COI = I_INTR ? bi * LY::KA : 0, COJ = J_INTR ? bj * LY::KA : 0;
    constexpr int GCI = I_INTR ? bi * LY::KA : LY::K + 6 * (bi - LY::NBI);
    constexpr int GCJ = J_INTR ? bj * LY::KA : LY::K + 6 * (bj - LY::NBI);
    const double *regi = I_INTR ? st.Jas : st.Jes[I_INTR ? 0 : bi - LY::NBI];
    const double *regj = J_INTR ? st.Jas : st.Jes[J_INTR ? 0 : bj - LY::NBI];
    const double *pi = regi + (size_t)g * 2 * P * RSI + COI;
    const double *pj = regj + (size_t)g * 2 * P * RSJ + COJ;
    const double *pr = st.rs + (size_t)g * 2 * P;

    constexpr int NACC = DIAG ? NI * (NI + 1) / 2 : NI * NJ;
    double acc[NACC];
    double accr[NI];
    double rr = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NI; i++) accr[i] = 0.0;

    if (valid) {
        for (int k = s; k < 2 * P; k += S) {
            double a[NI];
            load_row<NI, (RSI % 2 == 0), COI % 2>(pi + (size_t)k * RSI, a);
            if constexpr (DIAG) {
                const double r = pr[k];
                int q = 0;
#pragma unroll
                for (int i = 0; i < NI; i++) {
#pragma unroll
                    for (int j = i; j < NI; j++) { acc[q] = fma(a[i], a[j], acc[q]); q++; }
                    accr[i] = fma(a[i], r, accr[i]);
                }
                if constexpr (LAST) rr = fma(r, r, rr);
            } else {
                double b[NJ];
                load_row<NJ, (RSJ % 2 == 0), COJ % 2>(pj + (size_t)k * RSJ, b);
#pragma unroll
                for (int i = 0; i < NI; i++)
#pragma unroll
                    for (int j = 0; j < NJ; j++) acc[i * NJ + j] = fma(a[i], b[j], acc[i * NJ + j]);
            }
        }
    }
    // combine the S row-splits (lanes g*S .. g*S+S-1)
    for (int off = S >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
        if constexpr (DIAG) {
#pragma unroll
            for (int i = 0; i < NI; i++) accr[i] += __shfl_xor_sync(0xffffffffu, accr[i], off);
            if constexpr (LAST) rr += __shfl_xor_sync(0xffffffffu, rr, off);
        }
    }
    if (valid && s == 0) {
        double *h = st.Hs + (size_t)g * LY::NE;
        if constexpr (DIAG) {
            int q = 0;
#pragma unroll
            for (int i = 0; i < NI; i++) {
#pragma unroll
                for (int j = i; j < NI; j++) h[pk(GCI + i, GCI + j, LY::W)] = acc[q++];
                h[pk(GCI + i, LY::D, LY::W)] = accr[i];
            }
            if constexpr (LAST) h[pk(LY::D, LY::D, LY::W)] = rr;
        } else {
#pragma unroll
            for (int i = 0; i < NI; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) h[pk(GCI + i, GCJ + j, LY::W)] = acc[i * NJ + j];
        }
    }
}

template <int MODEL, int L, int T>
__device__ __forceinline__ void dispatch_tile(const int t, const Stage<MODEL, L> &st, const int g, const int s,
                                              const int S, const int P, const bool valid)
{
    if constexpr (T < Layout<MODEL, L>::NT) {
        if (t == T) run_tile<MODEL, L, T>(st, g, s, S, P, valid);
        else dispatch_tile<MODEL, L, T + 1>(t, st, g, s, S, P, valid);
    }
}

// ---- the kernel -------------------------------------------------------------------
template <int MODEL, int L>
__global__ void __launch_bounds__(256, 2)
reproj_eval_kernel(const EvalArgs args, const int G)
{
    using LY = Layout<MODEL, L>;
    using CAM = Camera<MODEL>;
    constexpr int K = LY::K;
    extern __shared__ __align__(16) double smem[];
    const int P = args.P;
    const Stage<MODEL, L> st(smem, G, P);
    const int tid = threadIdx.x;
    const int img0 = blockIdx.x * G;
    const int nv = min(G, args.n_img - img0);

    // ---- phase 0: per-image transform chain -------------------------------------
    if (tid < nv) {
        const int img = img0 + tid;
        double *ps = st.pose + (size_t)tid * LY::POSE;
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
    double intr[K];
#pragma unroll
    for (int i = 0; i < K; i++) intr[i] = __ldg(args.intr + i);
    const bool first_direct = (args.inverse[0] == 0);   // R12 of element 0 is the identity
    __syncthreads();

    // ---- phase A: one thread per (image, corner) ----------------------------------
    for (int idx = tid; idx < nv * P; idx += blockDim.x) {
        const int g = idx / P;
        const int c = idx - g * P;
        const int img = img0 + g;
        const double *ps = st.pose + (size_t)g * LY::POSE;
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
            const double cu0 = w1 * Pu[2] - w2 * Pu[1], cu1 = w2 * Pu[0] - w0 * Pu[2], cu2 = w0 * Pu[1] - w1 * Pu[0];
            const double cv0 = w1 * Pv[2] - w2 * Pv[1], cv1 = w2 * Pv[0] - w0 * Pv[2], cv2 = w0 * Pv[1] - w1 * Pv[0];
            double ju[6], jv[6];
            if (e == 0 && first_direct) {
#pragma unroll
                for (int j = 0; j < 3; j++) { ju[j] = Pu[j]; jv[j] = Pv[j]; }
            } else {
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    ju[j] = fma(Pu[2], pe[6 + j], fma(Pu[1], pe[3 + j], Pu[0] * pe[j]));
                    jv[j] = fma(Pv[2], pe[6 + j], fma(Pv[1], pe[3 + j], Pv[0] * pe[j]));
                }
            }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                ju[3 + j] = fma(cu2, pe[15 + j], fma(cu1, pe[12 + j], cu0 * pe[9 + j]));
                jv[3 + j] = fma(cv2, pe[15 + j], fma(cv1, pe[12 + j], cv0 * pe[9 + j]));
            }
            store_row<6>(st.Jes[e] + row * 6, ju, true);
            store_row<6>(st.Jes[e] + (row + 1) * 6, jv, true);
        }
    }
    fence_proxy_async_smem();   // make the generic-proxy smem writes visible to the TMA engine
    __syncthreads();

    // ---- phase S: stream the Ceres-layout blocks out with TMA bulk copies ----------
    bool issued = false;
    if (tid == 0) {
        const size_t rows = (size_t)nv * 2 * P;
        if (args.r) { bulk_store_chunked(args.r + (size_t)img0 * 2 * P, st.rs, rows * 8); issued = true; }
        if (args.Ja) { bulk_store_chunked(args.Ja + (size_t)img0 * 2 * P * K, st.Jas, rows * K * 8); issued = true; }
#pragma unroll
        for (int e = 0; e < L; e++)
            if (args.Je[e]) { bulk_store_chunked(args.Je[e] + (size_t)img0 * 2 * P * 6, st.Jes[e], rows * 48); issued = true; }
        if (issued) bulk_commit();
    }

    // ---- phase B: per-image normal-equation blocks ---------------------------------
    if (args.H) {
        const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
        const int S = 32 / G;
        const int g = lane / S, s = lane - g * S;
        for (int t = warp; t < LY::NT; t += nw) dispatch_tile<MODEL, L, 0>(t, st, g, s, S, P, g < nv);
        __syncthreads();
        double *Hg = args.H + (size_t)img0 * LY::NE;
        for (int i = tid; i < nv * LY::NE; i += blockDim.x) Hg[i] = st.Hs[i];
    }
    if (issued) bulk_wait_read_all();   // shared memory must outlive the TMA reads
}

template <int MODEL, int L>
cudaError_t launch_one(const EvalArgs &args, cudaStream_t stream, unsigned long long *launches)
{
    using LY = Layout<MODEL, L>;
    int G, threads;
    const long long bytes = eval_smem_bytes(MODEL, L, args.P, &G, &threads);
    if (bytes < 0) return cudaErrorInvalidValue;
    static int configured_bytes[64];    // per instantiation and device; zero-initialised
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (bytes > configured_bytes[dev]) {
        cudaError_t e = cudaFuncSetAttribute(reproj_eval_kernel<MODEL, L>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        configured_bytes[dev] = (int)bytes;
    }
    if (args.n_img <= 0) return cudaSuccess;
    const int grid = (args.n_img + G - 1) / G;
    reproj_eval_kernel<MODEL, L><<<grid, threads, (size_t)bytes, stream>>>(args, G);
    if (launches) (*launches)++;
    (void)sizeof(LY);
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

template <int MODEL> long long smem_for(int L, int G, int P)
{
    switch (L) {
    case 1: return Layout<MODEL, 1>::smem_doubles(G, P) * 8;
    case 2: return Layout<MODEL, 2>::smem_doubles(G, P) * 8;
    case 3: return Layout<MODEL, 3>::smem_doubles(G, P) * 8;
    case 4: return Layout<MODEL, 4>::smem_doubles(G, P) * 8;
    case 5: return Layout<MODEL, 5>::smem_doubles(G, P) * 8;
    default: return -1;
    }
}

}  // namespace

long long eval_smem_bytes(int model, int chain_len, int P, int *images_per_cta, int *threads)
{
    if (P < 1 || chain_len < 1 || chain_len > MAX_CHAIN) return -1;
    const long long LIMIT = 227 * 1024;
    for (int G = 8; G >= 1; G >>= 1) {
        if (G > 1 && G * P > 256) continue;
        long long b;
        switch (model) {
        case MODEL_EUCM: b = smem_for<MODEL_EUCM>(chain_len, G, P); break;
        case MODEL_UCM: b = smem_for<MODEL_UCM>(chain_len, G, P); break;
        case MODEL_MEI: b = smem_for<MODEL_MEI>(chain_len, G, P); break;
        default: return -1;
        }
        if (b < 0) return -1;
        if (b > LIMIT) continue;
        int t = ((G * P + 31) / 32) * 32;
        if (t < 64) t = 64;
        if (t > 256) t = 256;
        if (images_per_cta) *images_per_cta = G;
        if (threads) *threads = t;
        return b;
    }
    return -1;
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
