// vg_solver_fast.cu -- the Levenberg-Marquardt step for the plain calibration structure: one rank, one dataset
// whose image i is the only block that touches free pose i (every monocular configuration), no coupled or constant
// sequence elements, at most FAST_MAX_KS free shared parameters.  Same arithmetic as vg_solver_kernels.cu
// (pose_factor / finalize_gram / reduced_solve / pose_backsub, i.e. the elimination Ceres performs inside
// ceres::Solve, unified_calibration.cpp:39-53) in two launches without a serial tail:
//   fast_factor   32 poses per block: the block's packed blocks arrive with one coalesced pass (no gather through the
//                 pose list), damped 6x6 Cholesky, Z = L^-1 E^T, z = L^-1 b, the block's Schur terms from shared
//                 memory; the last block of every group of FAST_GROUP blocks folds the group's rows into one
//   fast_backsub  every block first adds the few group rows up and solves the small reduced system itself (a few
//                 hundred flops, redundantly: nobody waits for a last block), then back-substitutes its poses;
//                 block 0 also leaves the candidate shared parameters and the scalars of the accept / reject decision
// Reciprocals and square roots are the hardware seed + Newton steps of vg_math.cuh (1-2 ulp): the IEEE division /
// square-root sequences were half of the instructions of these latency-bound kernels.
// All sums have a fixed order: bit-reproducible run to run.
#include "vg_solver_kernels.cuh"
#include "vg_math.cuh"

#include <cstring>
#include <type_traits>

namespace vg {

namespace {

__host__ __device__ inline int pk(int a, int b, int W) { return a * W - a * (a - 1) / 2 + (b - a); }
__device__ __forceinline__ int pks(int a, int b, int W) { return a <= b ? pk(a, b, W) : pk(b, a, W); }
__device__ __forceinline__ constexpr int lt(int i, int j) { return i * (i + 1) / 2 + j; }


constexpr int LANES = 8;                    // lanes per pose
constexpr int F_THREADS = 256, F_POSES = F_THREADS / LANES;
constexpr int B_THREADS = 128, B_POSES = B_THREADS / LANES;

__device__ __forceinline__ double sum_rows_batched(const double *p, const size_t stride, const int r0, const int r1, const int rstep)
{
    constexpr int B = 24;
    double s = 0.0;
    for (int b = r0; b < r1; b += B * rstep) {
        double v[B];
#pragma unroll
        for (int i = 0; i < B; i++) v[i] = b + i * rstep < r1 ? __ldcg(p + (size_t)(b + i * rstep) * stride) : 0.0;
#pragma unroll
        for (int i = 0; i < B; i++) s += v[i];
    }
    return s;
}

// max of p[r * stride], r in [r0, r1): every load in flight before the first comparison
__device__ __forceinline__ double max_rows_batched(const double *p, const size_t stride, const int r0, const int r1)
{
    constexpr int B = 16;
    double m = 0.0;
    for (int b = r0; b < r1; b += B) {
        double v[B];
#pragma unroll
        for (int i = 0; i < B; i++) v[i] = b + i < r1 ? __ldcg(p + (size_t)(b + i) * stride) : 0.0;
#pragma unroll
        for (int i = 0; i < B; i++) m = fmax(m, v[i]);
    }
    return m;
}

__device__ __forceinline__ unsigned int take_ticket(unsigned int *counter)
{
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
}

// entry t of a row of Schur terms -> (a, b): t < Ks (Ks + 1) / 2 the upper triangle of Z^T Z row by row, then Z^T z
__device__ __forceinline__ void pair_of(int t, int Ks, int &a, int &b)
{
    const int ntri = Ks * (Ks + 1) / 2;
    if (t < ntri) {
        a = 0;
        int rem = t;
        while (rem >= Ks - a) { rem -= Ks - a; a++; }
        b = a + rem;
    } else {
        a = t - ntri; b = Ks;
    }
}

// ---- factorisation -----------------------------------------------------------------------------
// dynamic shared memory: [packed blocks of the block's poses: F_POSES x ne | Z, z of them: F_POSES x (6 Ks + 6)]
// WC / PCC > 0: the block width and the pose block's first column are compiled in (every index into a packed block is
// then a constant): the three camera models with a chain of one transform; 0: read from the descriptor.
template <int WC, int PCC>
__global__ void __launch_bounds__(F_THREADS)
fast_factor_kernel(const FastDesc d, const int n_pose, const int Ks, double *scale, const LmConsts lm_in, double *ws,
                   double *rows, double *grp_rows, double *gmax_rows, unsigned int *tickets, int *fail_flag, const FastLm flm)
{
    // (programmatic dependent launch: see launch_fast_step.  This kernel does not release its own dependents early:
    // measured, the evaluation's persistent CTAs waiting in the SM slots cost this grid more -- 65.7 -> 73.3 us per
    // iteration at C2 -- than the hidden launch latency gives back)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    LmConsts lm = lm_in;
    const double *Hsrc = d.H;
    const int nblk = flm.st ? (int)gridDim.x - 1 : (int)gridDim.x;      // (with the loop's state on the device: one extra block)
    if (flm.st) {
        // the loop's state is on the device (vg_lm_dev.cuh): over -> nothing to do; else radius and buffers from there,
        // and an accepted candidate's poses, slab and reduced system become the current ones
        LmState *st = flm.st;
        const int4 f = __ldcg(reinterpret_cast<const int4 *>(st));           // done, limits, migrate, hcur
        const int isc = __ldcg(&st->init_scale);
        const double rad = __ldcg(&st->radius);
        if ((int)blockIdx.x == nblk) {
            // the extra block: the previous launch's decision goes to the host ring from here -- writes that cross PCIe,
            // and the fence behind them, under this kernel's own work instead of at the end of the evaluation
            if (threadIdx.x < 32) {
                const unsigned long long rcd = __ldcg(&st->records), pub = __ldcg(&st->published);
                if (pub < rcd) {
                    const int lane = threadIdx.x;
                    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&st->rec);
                    unsigned long long *dst = reinterpret_cast<unsigned long long *>(flm.ring + ((rcd - 1) % LM_RING));
                    for (int i = 1 + lane; i < (int)(sizeof(LmRecord) / 8); i += 32) dst[i] = __ldcg(src + i);
                    if (f.x) {          // the solve is over: the final shared parameters (set C if the last step was accepted)
                        const double *sl = f.z ? flm.slab_c : flm.slab_a;
                        for (int i = lane; i < flm.slab_n; i += 32) flm.final_out[i] = __ldcg(sl + i);
                    }
                    __threadfence_system();
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(dst), "l"(rcd) : "memory");
                        st->published = rcd;
                    }
                }
            }
            return;
        }
        if (f.x) return;
        VG_LM_STAMP(st, 0, 0)
        VG_LM_PHASE(st, 0, 0)
        lm.radius = rad;
        lm.init_scale = isc;
        if (f.w) Hsrc = flm.H_alt;
        if (f.z) {
            const int q0 = blockIdx.x * F_POSES * 6, qn = min(F_POSES, n_pose - (int)blockIdx.x * F_POSES) * 6;
            for (int i = threadIdx.x; i < qn; i += F_THREADS) flm.pose_a[q0 + i] = __ldcg(flm.pose_c + q0 + i);
            if (blockIdx.x == 0) {
                for (int i = threadIdx.x; i < flm.slab_n; i += F_THREADS) flm.slab_a[i] = __ldcg(flm.slab_c + i);
                for (int i = threadIdx.x; i < flm.red_n; i += F_THREADS) flm.red_a[i] = __ldcg(flm.red_c + i);
            }
        }
    }
    extern __shared__ double sm[];
    __shared__ int s_lc[FAST_MAX_KS + 1];
    __shared__ double s_red[F_THREADS];
    __shared__ int s_last;
    const int tid = threadIdx.x, sub = tid & (LANES - 1), lp = tid / LANES;
    const int p0 = blockIdx.x * F_POSES, np = min(F_POSES, n_pose - p0), p = p0 + lp;
    const int W = WC ? WC : d.W, pc = WC ? PCC : d.pose_col, ne = WC ? WC * (WC + 1) / 2 : d.ne;
    double *Hs = sm, *zs = sm + (size_t)F_POSES * ne;
    const int zstride = 6 * Ks + 6, npair = Ks * (Ks + 1) / 2 + Ks;
    // local column of shared parameter j (the residual column for j == Ks); -1: the dataset does not touch it
    if (tid <= Ks) {
        int lc = tid == Ks ? W - 1 : -1;
#pragma unroll
        for (int q = 0; q < FAST_MAX_KS; q++)
            if (q < d.n_sl && d.sl_idx[q] == tid) lc = d.sl_col[q];
        s_lc[tid] = lc;
    }
    // the Jacobi scaling of this thread's pose (not at iteration 0, when it is computed below): in flight with the blocks
    double scl[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
    if (!lm.init_scale && p < n_pose) {
#pragma unroll
        for (int k = 0; k < 6; k++) scl[k] = __ldg(scale + (size_t)p * 6 + k);
    }
    {
        const double *src = Hsrc + (size_t)p0 * ne;
        for (int i = tid; i < np * ne; i += F_THREADS) Hs[i] = __ldcs(src + i);
    }
    __syncthreads();
    VG_LM_PHASE(flm.st, 0, 1)
    double gmax = 0.0;
    if (p < n_pose) {
        const double *Hp = Hs + (size_t)lp * ne;
        double Lm[21];
        static_for<0, 6>([&](auto ic) {
            static_for<0, VG_IDX(ic) + 1>([&](auto jc) {
                constexpr int i = VG_IDX(ic), j = VG_IDX(jc);
                Lm[lt(i, j)] = Hp[pk(pc + j, pc + i, W)];
            });
        });
        double *w = ws + (size_t)p * pose_ws_stride(Ks);
        double lam[6], invd[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
        bool empty = true;
        static_for<0, 6>([&](auto kc) {
            constexpr int k = VG_IDX(kc);
            const double ckk = Lm[lt(k, k)];
            if (ckk != 0.0) empty = false;
            double sc;
            if (lm.init_scale) {
                sc = lm.jacobi_scaling ? fast_rcp(1.0 + (ckk > 0.0 ? ckk * fast_rsqrt(ckk) : 0.0)) : 1.0;
                if (sub == 0) scale[(size_t)p * 6 + k] = sc;
            } else {
                sc = scl[k];
            }
            const double s2 = sc * sc;
            lam[k] = fmin(fmax(s2 * ckk, lm.min_diag), lm.max_diag) * fast_rcp(lm.radius * s2);
        });
        bool ok = true;
        if (empty) {
            // a pose nothing observes: identity factor, zero step
            static_for<0, 6>([&](auto ic) {
                static_for<0, VG_IDX(ic) + 1>([&](auto jc) {
                    constexpr int i = VG_IDX(ic), j = VG_IDX(jc);
                    Lm[lt(i, j)] = (i == j) ? 1.0 : 0.0;
                });
                lam[VG_IDX(ic)] = 0.0;
            });
        } else {
            static_for<0, 6>([&](auto kc) { Lm[lt(VG_IDX(kc), VG_IDX(kc))] += lam[VG_IDX(kc)]; });
            static_for<0, 6>([&](auto jc) {
                constexpr int j = VG_IDX(jc);
                double s = Lm[lt(j, j)];
                static_for<0, j>([&](auto kc) { s = fma(-Lm[lt(j, VG_IDX(kc))], Lm[lt(j, VG_IDX(kc))], s); });
                if (!(s > 0.0)) { ok = false; s = 1.0; }
                const double inv = fast_rsqrt(s);
                Lm[lt(j, j)] = s * inv;
                invd[j] = inv;
                static_for<j + 1, 6>([&](auto ic) {
                    constexpr int i = VG_IDX(ic);
                    double t = Lm[lt(i, j)];
                    static_for<0, j>([&](auto kc) { t = fma(-Lm[lt(i, VG_IDX(kc))], Lm[lt(j, VG_IDX(kc))], t); });
                    Lm[lt(i, j)] = t * inv;
                });
            });
        }
        if (sub == 0) {
            if (!ok) atomicExch(fail_flag, 1);
            // (the diagonal of the stored factor holds the RECIPROCALS of L's diagonal: all fast_backsub needs of it)
            static_for<0, 6>([&](auto ic) {
                static_for<0, VG_IDX(ic) + 1>([&](auto jc) {
                    constexpr int i = VG_IDX(ic), j = VG_IDX(jc);
                    w[lt(i, j)] = i == j ? invd[i] : Lm[lt(i, j)];
                });
                w[21 + VG_IDX(ic)] = lam[VG_IDX(ic)];
            });
        }
        double *zrow = zs + (size_t)lp * zstride;
        // the group's lanes take the columns: col < Ks -> column col of E^T (Z = L^-1 E^T), col == Ks -> the gradient
        for (int col = sub; col <= Ks; col += LANES) {
            const int lc = s_lc[col];
            double e[6];
            static_for<0, 6>([&](auto ic) { e[VG_IDX(ic)] = lc >= 0 ? Hp[pks(lc, pc + VG_IDX(ic), W)] : 0.0; });
            if (col == Ks) static_for<0, 6>([&](auto kc) { gmax = fmax(gmax, fabs(e[VG_IDX(kc)])); });
            static_for<0, 6>([&](auto ic) {           // e <- L^-1 e
                constexpr int i = VG_IDX(ic);
                double s = e[i];
                static_for<0, i>([&](auto kc) { s = fma(-Lm[lt(i, VG_IDX(kc))], e[VG_IDX(kc)], s); });
                e[i] = s * invd[i];
            });
            if (col == Ks) {
                static_for<0, 6>([&](auto kc) { w[27 + VG_IDX(kc)] = e[VG_IDX(kc)]; zrow[6 * Ks + VG_IDX(kc)] = e[VG_IDX(kc)]; });
            } else {
                static_for<0, 6>([&](auto kc) { w[33 + VG_IDX(kc) * Ks + col] = e[VG_IDX(kc)]; zrow[VG_IDX(kc) * Ks + col] = e[VG_IDX(kc)]; });
            }
        }
    }
    // max |g| of the block
    for (int off = 16; off > 0; off >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
    if ((tid & 31) == 0) s_red[tid >> 5] = gmax;
    __syncthreads();                          // also: the block's Z, z are in shared memory
    VG_LM_PHASE(flm.st, 0, 2)
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < F_THREADS / 32; i++) m = fmax(m, s_red[i]);
        gmax_rows[blockIdx.x] = m;
    }
    __syncthreads();
    // the block's row of Schur terms: thread = (entry t, slice q of the block's poses), slices added in slice order
    if (npair > 0) {
        const int Q = max(1, F_THREADS / npair), per = (F_POSES + Q - 1) / Q;
        for (int t0 = 0; t0 < npair; t0 += F_THREADS) {
            const int q = Q > 1 ? tid / npair : 0;
            const int t = Q > 1 ? tid - q * npair : t0 + tid;
            double s = 0.0;
            if (t < npair && q < Q) {
                int a, b;
                pair_of(t, Ks, a, b);
                const int q0 = q * per, q1 = min(np, q0 + per);
                for (int l = q0; l < q1; l++) {
                    const double *Z = zs + (size_t)l * zstride, *z = Z + 6 * Ks;
#pragma unroll
                    for (int k = 0; k < 6; k++) s = fma(Z[k * Ks + a], (b < Ks) ? Z[k * Ks + b] : z[k], s);
                }
            }
            if (Q > 1) {
                s_red[tid] = s;
                __syncthreads();
                if (tid < npair) {
                    double tot = 0.0;
                    for (int qq = 0; qq < Q; qq++) tot += s_red[qq * npair + tid];
                    rows[(size_t)blockIdx.x * npair + tid] = tot;
                }
                break;
            }
            if (t < npair) rows[(size_t)blockIdx.x * npair + t] = s;
        }
    }
    // the last block of the group folds the group's rows (and its max |g|) into one
    const int grp = blockIdx.x / FAST_GROUP, b0 = grp * FAST_GROUP, b1 = min(nblk, b0 + FAST_GROUP);
    __syncthreads();
    VG_LM_PHASE(flm.st, 0, 3)
    if (tid == 0) s_last = take_ticket(tickets + grp) == (unsigned)(b1 - b0 - 1);
    __syncthreads();
    VG_LM_PHASE(flm.st, 0, 4)
    VG_LM_STAMP(flm.st, 0, 1)
    if (!s_last) return;
    for (int t = tid; t < npair; t += F_THREADS) grp_rows[(size_t)grp * (npair + 1) + t] = sum_rows_batched(rows + t, npair, b0, b1, 1);
    if (tid == F_THREADS - 1) {
        grp_rows[(size_t)grp * (npair + 1) + npair] = max_rows_batched(gmax_rows, 1, b0, b1);
        tickets[grp] = 0;
    }
    __syncthreads();
    VG_LM_STAMP(flm.st, 0, 1)
}

// ---- reduced solve + back substitution ------------------------------------------------------------
// static shared memory of the small solve: Ks <= FAST_MAX_KS
struct SolveSm {
    double S[FAST_MAX_KS * FAST_MAX_KS], A[FAST_MAX_KS * FAST_MAX_KS], Sred[FAST_MAX_KS * FAST_MAX_KS];
    double rhs[FAST_MAX_KS], da[FAST_MAX_KS], g[FAST_MAX_KS], v[FAST_MAX_KS], xs[FAST_MAX_KS], lo[FAST_MAX_KS], hi[FAST_MAX_KS],
        sc[FAST_MAX_KS], xn[FAST_MAX_KS], Ad[FAST_MAX_KS];
    double gmax;
    int ok;
};

template <int KSC>                                    // Ks <= KSC: what the register arrays over the shared parameters are unrolled to
__global__ void __launch_bounds__(B_THREADS, 6)       // every block of a C2-sized problem resident at once
fast_backsub_kernel(const int n_pose, const int Ks, const int n_grp, const double *grp_rows, const SolveArgs sa,
                    const LmConsts lm_in, const double *seq_cur, double *seq_cand, const double *ws, double *partial,
                    unsigned int *ticket, int *fail_flag, double *host_out, const double *fail_src, const LmState *st,
                    const FastShared fs)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    LmConsts lm = lm_in;
    // the loop's state on the device (vg_lm_dev.cuh): in flight with everything else this block reads first; nothing is
    // written before "done" has been looked at
    int st_done = 0;
    if (st) {
        st_done = __ldcg(&st->done);
        lm.init_scale = __ldcg(&st->init_scale);
        lm.radius = __ldcg(&st->radius);
    }
    // this lane's part of its pose's rows, in flight while the block solves the reduced system: lane k < 6 takes
    // z_k, row k of Z, the factor and its own component of the pose
    const int sub = threadIdx.x & (LANES - 1);
    const int p = blockIdx.x * B_POSES + threadIdx.x / LANES;
    const bool active = p < n_pose && sub < 6;
    const double *w = ws + (size_t)(p < n_pose ? p : 0) * pose_ws_stride(Ks);
    double zk = 0.0, lamk = 0.0, xk = 0.0, Zr[KSC];
#pragma unroll
    for (int a = 0; a < KSC; a++) Zr[a] = 0.0;
    if (active) {
        zk = __ldcg(w + 27 + sub); lamk = __ldcg(w + 21 + sub); xk = __ldcg(seq_cur + (size_t)p * 6 + sub);
#pragma unroll
        for (int a = 0; a < KSC; a++) if (a < Ks) Zr[a] = __ldcg(w + 33 + sub * Ks + a);
    }
    __shared__ SolveSm q;
    __shared__ double sh[3][B_THREADS / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const int npair = Ks * (Ks + 1) / 2 + Ks, ntri = Ks * (Ks + 1) / 2;
    const double *A = sa.red_cur + red_off_A(Ks), *ga = sa.red_cur + red_off_g(Ks);
    // ---- the reduced system, by every block: the three kinds of loads on different threads (for the usual sizes), so that
    // their round trips to L2 overlap instead of following one another ----
    for (int t = tid; t <= npair; t += B_THREADS) {
        if (t < npair) {
            const double s = sum_rows_batched(grp_rows + t, npair + 1, 0, n_grp, 1);
            if (t < ntri) {
                int a, b;
                pair_of(t, Ks, a, b);
                q.Sred[a * Ks + b] = s;
                q.Sred[b * Ks + a] = s;
            } else {
                q.v[t - ntri] = s;
            }
        } else {
            q.gmax = max_rows_batched(grp_rows + npair, npair + 1, 0, n_grp);
        }
    }
    for (int i = (tid + 3 * B_THREADS / 4) % B_THREADS; i < Ks * Ks; i += B_THREADS) q.A[i] = __ldcg(A + i);       // from thread 32 on
    for (int j = (tid + B_THREADS / 4) % B_THREADS; j < Ks; j += B_THREADS) {                                   // from thread 96 on
        const double gj = __ldcg(ga + j), xj = __ldcg(sa.slab_cur + fs.off[j]), sj = __ldcg(sa.scale_a + j);
        q.g[j] = gj;
        q.xs[j] = xj;
        q.lo[j] = fs.lo[j]; q.hi[j] = fs.hi[j];
        q.sc[j] = lm.init_scale ? 0.0 : sj;
    }
    if (st_done) return;
    VG_LM_STAMP(const_cast<LmState *>(st), 1, 0)
    VG_LM_PHASE(const_cast<LmState *>(st), 1, 0)
    __syncthreads();
    VG_LM_PHASE(const_cast<LmState *>(st), 1, 1)
    if (tid < 32) {
        // One warp, lane j = row j of the reduced system in registers: right-looking Cholesky and the two substitutions
        // through shuffles (the same operations in the same order as reduced_solve_kernel's; Ks <= KSC < 32).
        constexpr unsigned FULL = 0xffffffffu;
        const int j = lane;
        const bool row = j < Ks;
        double Srow[KSC], rhs = 0.0;
        {
            const double ajj = row ? q.A[j * Ks + j] : 1.0;
            const double sc = lm.init_scale ? (lm.jacobi_scaling ? fast_rcp(1.0 + (ajj > 0.0 ? ajj * fast_rsqrt(ajj) : 0.0)) : 1.0)
                                            : (row ? q.sc[j] : 1.0);
            if (row) q.sc[j] = sc;
            const double s2 = sc * sc;
            const double damp = fmin(fmax(s2 * ajj, lm.min_diag), lm.max_diag) * fast_rcp(lm.radius * s2);
#pragma unroll
            for (int k = 0; k < KSC; k++)
                Srow[k] = (row && k < Ks) ? q.A[j * Ks + k] - q.Sred[j * Ks + k] + (k == j ? damp : 0.0) : 0.0;
            if (row) rhs = -(q.g[j] - q.v[j]);
        }
        VG_LM_PHASE(const_cast<LmState *>(st), 1, 5)
        bool ok = true;
#pragma unroll
        for (int c = 0; c < KSC; c++) {
            if (c < Ks) {
                double piv = __shfl_sync(FULL, Srow[c], c);
                if (!(piv > 0.0)) { ok = false; piv = 1.0; }
                const double inv = fast_rsqrt(piv);
                const double lic = Srow[c] * inv;                   // L_jc (rows j > c)
#pragma unroll
                for (int k = c + 1; k < KSC; k++) {
                    if (k < Ks) {
                        const double lkc = __shfl_sync(FULL, lic, k);
                        Srow[k] = fma(-lic, lkc, Srow[k]);
                    }
                }
                Srow[c] = (j == c) ? inv : lic;                     // the diagonal holds the reciprocal of the pivot
            }
        }
        VG_LM_PHASE(const_cast<LmState *>(st), 1, 6)
        // L y = rhs, column by column: lane c's entry is final when its turn comes
        double y = rhs;
#pragma unroll
        for (int c = 0; c < KSC; c++) {
            if (c < Ks) {
                const double yc = __shfl_sync(FULL, y * Srow[c], c);
                if (j == c) y = yc;
                else if (j > c) y = fma(-Srow[c], yc, y);
            }
        }
        VG_LM_PHASE(const_cast<LmState *>(st), 1, 7)
        // L^T x = y: lane j needs column j of L -- through shared memory, once
#pragma unroll
        for (int k = 0; k < KSC; k++)
            if (row && k <= j) q.S[j * Ks + k] = Srow[k];
        __syncwarp();
        double Lt[KSC], x[KSC];
#pragma unroll
        for (int k = 0; k < KSC; k++) { Lt[k] = (row && k > j && k < Ks) ? q.S[k * Ks + j] : 0.0; x[k] = 0.0; }
#pragma unroll
        for (int i = KSC - 1; i >= 0; i--) {
            if (i < Ks) {
                double t = y;
#pragma unroll
                for (int k = i + 1; k < KSC; k++)
                    if (k < Ks) t = fma(-Lt[k], x[k], t);
                x[i] = __shfl_sync(FULL, t * Srow[i], i);            // (lane i's own value is the one that counts)
            }
        }
        VG_LM_PHASE(const_cast<LmState *>(st), 1, 8)
        if (lane == 0) {
            q.ok = ok ? 1 : 0;
#pragma unroll
            for (int k = 0; k < KSC; k++)
                if (k < Ks) q.da[k] = ok ? x[k] : 0.0;
        }
    }
    __syncthreads();
    VG_LM_PHASE(const_cast<LmState *>(st), 1, 2)
    // ---- block 0: what the reduced solve leaves for the host and for the candidate's evaluation ----
    if (blockIdx.x == 0) {
        double *out = sa.red_cand + red_size(Ks, sa.nranks);
        for (int i = tid; i < sa.slab_n; i += B_THREADS) {       // constants and padding carry over
            const double x = sa.slab_cur[i];
            sa.slab_cand[i] = x;
            out[SOLVE_OUT + i] = x;
        }
        __syncthreads();
        for (int j = tid; j < Ks; j += B_THREADS) {
            double t = 0.0;
            for (int k = 0; k < Ks; k++) t = fma(q.A[j * Ks + k], q.da[k], t);
            q.Ad[j] = t;
            sa.delta_a[j] = q.da[j];
            if (lm.init_scale) sa.scale_a[j] = q.sc[j];
            const double xn = fmin(fmax(q.xs[j] + q.da[j], q.lo[j]), q.hi[j]);      // box bounds by projection
            q.xn[j] = xn;
            sa.slab_cand[sa.sh_off[j]] = xn;
            out[SOLVE_OUT + sa.sh_off[j]] = xn;
        }
        __syncthreads();
        if (tid == 0) {
            double gd = 0.0, dHd = 0.0, step2 = 0.0, x2 = 0.0, gmax = q.gmax;
            for (int j = 0; j < Ks; j++) {
                gd = fma(q.g[j], q.da[j], gd);
                dHd = fma(q.da[j], q.Ad[j], dHd);
                x2 = fma(q.xs[j], q.xs[j], x2);
                step2 = fma(q.xn[j] - q.xs[j], q.xn[j] - q.xs[j], step2);
                gmax = fmax(gmax, fabs(q.xs[j] - fmin(fmax(q.xs[j] - q.g[j], q.lo[j]), q.hi[j])));
            }
            // failed pose factorisations: this rank's flag, or (several ranks) the count the exchange summed
            const double failed = fail_src ? __ldcg(fail_src) : (double)atomicExch(fail_flag, 0);
            out[0] = gmax;
            out[1] = failed;
            out[2] = (double)q.ok;
            out[3] = gd; out[4] = dHd; out[5] = step2; out[6] = x2; out[7] = 0.0;
        }
    }
    // ---- back substitution: lane k < 6 of a pose's group: w_k = z_k + Z_k . delta_a, delta = -L^-T w ----
    double m = 0.0, st2 = 0.0, x2 = 0.0, wk = 0.0;
    if (active) {
        double s = zk;
#pragma unroll
        for (int a = 0; a < KSC; a++) if (a < Ks) s = fma(Zr[a], q.da[a], s);
        wk = s;
        m = -0.5 * s * s;
    }
    double wv[6];
#pragma unroll
    for (int j = 0; j < 6; j++) wv[j] = __shfl_sync(0xffffffffu, wk, (lane & ~(LANES - 1)) + j);
    if (active) {
        double Lm[21];
#pragma unroll
        for (int i = 0; i < 21; i++) Lm[i] = __ldcg(w + i);
#pragma unroll
        for (int i = 5; i >= 0; i--) {              // wv <- L^-T wv
            double s = wv[i];
#pragma unroll
            for (int k = i + 1; k < 6; k++) s = fma(-Lm[lt(k, i)], wv[k], s);
            wv[i] = s * Lm[lt(i, i)];           // (reciprocal diagonal, see fast_factor)
        }
        double yk = 0.0;
#pragma unroll
        for (int j = 0; j < 6; j++) yk = (j == sub) ? wv[j] : yk;
        const double dlt = -yk;
        seq_cand[(size_t)p * 6 + sub] = xk + dlt;
        m = fma(-0.5 * lamk * dlt, dlt, m);
        st2 = dlt * dlt;
        x2 = xk * xk;
    }
    // block sums in a fixed order: lanes of a warp through shuffles (xor tree), warps in index order
    for (int off = 16; off > 0; off >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, off);
        st2 += __shfl_xor_sync(0xffffffffu, st2, off);
        x2 += __shfl_xor_sync(0xffffffffu, x2, off);
    }
    if (lane == 0) { sh[0][tid >> 5] = m; sh[1][tid >> 5] = st2; sh[2][tid >> 5] = x2; }
    __syncthreads();
    VG_LM_PHASE(const_cast<LmState *>(st), 1, 3)
    if (tid < 3) {
        double s = 0.0;
        for (int i = 0; i < B_THREADS / 32; i++) s += sh[tid][i];
        partial[(size_t)blockIdx.x * 3 + tid] = s;
    }
    // (the loop's state on the device: the candidate's evaluation adds the rows up at its head, vg_eval_impl.cuh)
    if (st) return;
    // the last block adds the rows up
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = take_ticket(ticket) == gridDim.x - 1;
    __syncthreads();
    VG_LM_PHASE(const_cast<LmState *>(st), 1, 4)
    VG_LM_STAMP(const_cast<LmState *>(st), 1, 1)
    if (!s_last) return;
    if (tid == 0) *ticket = 0;
    const int qn = tid >> 5;
    if (qn < 3) {
        const double s = sum_rows_batched(partial + qn, 3, lane, gridDim.x, 32);
        double tot = 0.0;
        for (int l = 0; l < 32; l++) tot += __shfl_sync(0xffffffffu, s, l);
        if (lane == 0) sa.red_cand[red_off_model(Ks) + qn] = tot;       // (summed across the ranks by the candidate's evaluation)
    }
    // a polling host (no copy, no stream sync): block 0's scalars and the candidate slab, straight into host-mapped memory.
    // Done here, by the block that finishes last, so that no block waits on a fence behind writes that cross PCIe.
    if (host_out) {
        const double *out = sa.red_cand + red_size(Ks, sa.nranks);
        for (int i = tid; i < SOLVE_OUT + sa.slab_n; i += B_THREADS)
            host_out[i < SOLVE_OUT ? i : FAST_HOST_SLAB + (i - SOLVE_OUT)] = __ldcg(out + i);
    }
    __syncthreads();
    VG_LM_STAMP(const_cast<LmState *>(st), 1, 1)
}

// several ranks: this rank's Schur terms (its group rows folded into one), its max |g| and its failed factorisations
// summed / gathered across the ranks over peer memory (vg_peer.cuh) -> one row [S, v (npair) | max |g| | failures] that
// fast_backsub reads as if it were a single group row.  One block.
__global__ void __launch_bounds__(256)
fast_exchange_kernel(const int Ks, const int n_grp, const double *grp_rows, double *xbuf, double *row_out, int *fail_flag,
                     const PeerCtx pc_in, const LmState *st)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    PeerCtx pc = pc_in;
    if (st) {                                   // (every rank takes the same decisions: all of them leave, or none)
        const int dn = __ldcg(&st->done);
        const unsigned long long ep = __ldcg(&st->epoch);
        if (dn) return;
        pc.epoch = ep;
    }
    __shared__ double scratch[2048];
    const int tid = threadIdx.x, npair = Ks * (Ks + 1) / 2 + Ks;
    for (int t = tid; t < npair; t += blockDim.x) xbuf[t] = sum_rows_batched(grp_rows + t, npair + 1, 0, n_grp, 1);
    if (tid == blockDim.x - 1) {
        const double m = max_rows_batched(grp_rows + npair, npair + 1, 0, n_grp);
        for (int r = 0; r < pc.n; r++) xbuf[npair + r] = r == pc.rank ? m : 0.0;
        xbuf[npair + pc.n] = (double)atomicExch(fail_flag, 0);
    }
    __syncthreads();
    peer_allreduce(xbuf, npair + pc.n + 1, pc, scratch, 2048);
    __syncthreads();
    for (int t = tid; t < npair; t += blockDim.x) row_out[t] = xbuf[t];
    if (tid == 0) {
        double m = 0.0;
        for (int r = 0; r < pc.n; r++) m = fmax(m, xbuf[npair + r]);
        row_out[npair] = m;
        row_out[npair + 1] = xbuf[npair + pc.n];
    }
}

}  // namespace

int fast_factor_blocks(int n_pose) { return (n_pose + F_POSES - 1) / F_POSES; }
int fast_groups(int n_pose) { return (fast_factor_blocks(n_pose) + FAST_GROUP - 1) / FAST_GROUP; }
int fast_backsub_blocks(int n_pose) { return (n_pose + B_POSES - 1) / B_POSES; }

// where launch_fast_step keeps the back-substitution's rows of three sums inside its scratch
const double *fast_partial_rows(const double *scratch, int n_pose, int Ks)
{
    const size_t npair = (size_t)Ks * (Ks + 1) / 2 + Ks;
    const size_t nb = fast_factor_blocks(n_pose), ng = fast_groups(n_pose);
    return scratch + nb * npair + nb + ng * (npair + 1);
}

size_t fast_scratch(int n_pose, int Ks)
{
    const size_t npair = (size_t)Ks * (Ks + 1) / 2 + Ks;
    return (size_t)fast_factor_blocks(n_pose) * (npair + 1) + (size_t)fast_groups(n_pose) * (npair + 1) +
           3 * (size_t)fast_backsub_blocks(n_pose) + 2 * (npair + PEER_MAX_RANKS + 4) + 16;
}

template <int WC, int PCC>
static cudaError_t launch_factor(const FastDesc &d, int n_pose, int Ks, double *scale, const LmConsts &lm, double *ws, double *rows,
                                 double *grp_rows, double *gmax_rows, unsigned int *tickets, int *fail_flag, size_t smem,
                                 cudaStream_t stream, const FastLm &flm)
{
    static size_t configured[64];           // per instantiation and device; zero-initialised
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fast_factor_kernel<WC, PCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fast_factor_blocks(n_pose) + (flm.st ? 1 : 0)); cfg.blockDim = dim3(F_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fast_factor_kernel<WC, PCC>, d, n_pose, Ks, scale, lm, ws, rows, grp_rows, gmax_rows, tickets,
                              fail_flag, flm);
}

cudaError_t launch_fast_step(const FastDesc &d, int n_pose, int Ks, double *scale, LmConsts lm, double *ws, double *scratch,
                             unsigned int *tickets, int *fail_flag, const SolveArgs &sa, const double *seq_cur,
                             double *seq_cand, bool backsub, SolverLaunch sl, cudaEvent_t between, double *host_out,
                             const PeerCtx *peer, const FastLm *flm_in, const FastShared &fs)
{
    FastLm flm;
    memset(&flm, 0, sizeof flm);
    if (flm_in) flm = *flm_in;
    if (Ks < 1 || Ks > FAST_MAX_KS || n_pose < 1 || (sa.nranks != 1 && !peer)) return cudaErrorInvalidValue;
    const int nb = fast_factor_blocks(n_pose), ng = fast_groups(n_pose), npair = Ks * (Ks + 1) / 2 + Ks;
    double *rows = scratch, *gmax_rows = rows + (size_t)nb * npair, *grp_rows = gmax_rows + nb,
           *partial = grp_rows + (size_t)ng * (npair + 1), *xbuf = partial + 3 * (size_t)fast_backsub_blocks(n_pose),
           *xrow = xbuf + (npair + PEER_MAX_RANKS + 4);
    const size_t smem = sizeof(double) * ((size_t)F_POSES * d.ne + (size_t)F_POSES * (6 * Ks + 6));
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    // Both kernels are launched with programmatic stream serialization (they start with griddepcontrol.wait): their
    // launch latency hides under the tail of the kernel before them.
    cudaError_t e;
    if (d.W == 13 && d.pose_col == 6) e = launch_factor<13, 6>(d, n_pose, Ks, scale, lm, ws, rows, grp_rows, gmax_rows, tickets, fail_flag, smem, sl.stream, flm);
    else if (d.W == 12 && d.pose_col == 5) e = launch_factor<12, 5>(d, n_pose, Ks, scale, lm, ws, rows, grp_rows, gmax_rows, tickets, fail_flag, smem, sl.stream, flm);
    else if (d.W == 17 && d.pose_col == 10) e = launch_factor<17, 10>(d, n_pose, Ks, scale, lm, ws, rows, grp_rows, gmax_rows, tickets, fail_flag, smem, sl.stream, flm);
    else e = launch_factor<0, 0>(d, n_pose, Ks, scale, lm, ws, rows, grp_rows, gmax_rows, tickets, fail_flag, smem, sl.stream, flm);
    if (sl.launches) count_launch(sl.launches);
    if (e != cudaSuccess) return e;
    if (between) cudaEventRecord(between, sl.stream);
    const double *rows_in = grp_rows, *fail_src = nullptr;
    int n_rows_in = ng;
    if (peer && peer->n > 1) {
        cudaLaunchConfig_t xcfg = {};
        xcfg.gridDim = dim3(1); xcfg.blockDim = dim3(256); xcfg.dynamicSmemBytes = 0; xcfg.stream = sl.stream;
        cudaLaunchAttribute xattr[1];
        xattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        xattr[0].val.programmaticStreamSerializationAllowed = 1;
        xcfg.attrs = xattr; xcfg.numAttrs = between ? 0 : 1;
        e = cudaLaunchKernelEx(&xcfg, fast_exchange_kernel, Ks, ng, (const double *)grp_rows, xbuf, xrow, fail_flag, *peer,
                               (const LmState *)flm.st);
        if (sl.launches) count_launch(sl.launches);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        rows_in = xrow; n_rows_in = 1; fail_src = xrow + npair + 1;
    }
    // (backsub == false: only the gradient test is still due -- the kernel still leaves the scalars; the candidate
    // poses it writes are never looked at)
    (void)backsub;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fast_backsub_blocks(n_pose)); cfg.blockDim = dim3(B_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = sl.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = between ? 0 : 1;
    if (Ks <= 6)
        e = cudaLaunchKernelEx(&cfg, fast_backsub_kernel<6>, n_pose, Ks, n_rows_in, rows_in, sa, lm, seq_cur, seq_cand,
                               (const double *)ws, partial, tickets + ng, fail_flag, host_out, fail_src, (const LmState *)flm.st, fs);
    else
        e = cudaLaunchKernelEx(&cfg, fast_backsub_kernel<FAST_MAX_KS>, n_pose, Ks, n_rows_in, rows_in, sa, lm, seq_cur, seq_cand,
                               (const double *)ws, partial, tickets + ng, fail_flag, host_out, fail_src, (const LmState *)flm.st, fs);
    if (sl.launches) count_launch(sl.launches);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace vg
