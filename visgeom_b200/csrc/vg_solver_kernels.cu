// vg_solver_kernels.cu -- see vg_solver_kernels.cuh.  All reductions are two-stage
// (per-block partials in a fixed layout, then one block summing them in index order),
// so results are bit-reproducible run to run for a fixed problem.
#include "vg_solver_kernels.cuh"
#include "vg_math.cuh"

#include <cstring>
#include <vector>

namespace vg {

namespace {

__host__ __device__ inline int pk(int a, int b, int W) { return a * W - a * (a - 1) / 2 + (b - a); }
__device__ __forceinline__ int pks(int a, int b, int W) { return a <= b ? pk(a, b, W) : pk(b, a, W); }

constexpr int SOLVE_THREADS = 256;
// A pose is worked on by a group of 8 lanes: with one thread per pose 10 000 poses are 2 warps per SM, each
// running ~5 000 dependent instructions -- pure issue latency (ncu: 3 % of the warp slots active, 34 us).  The
// 6x6 factor is computed by every lane of the group, the Ks + 1 right-hand-side columns are spread over them.
constexpr int POSE_LANES = 8;
constexpr int POSE_THREADS = 128;     // pose_backsub: 16 poses per block
constexpr int BACK_POSES = POSE_THREADS / POSE_LANES;
constexpr int FACTOR_THREADS = 256;   // pose_factor: 32 poses per block
constexpr int FACTOR_POSES = FACTOR_THREADS / POSE_LANES;

// sum of p[r * stride], r in [r0, r1) step rstep, added in index order; the (L2) loads of 16 rows are in flight together
// (a plain `s += __ldcg(..)` loop issues one load per round trip: the block that finishes last would spend tens of
// microseconds on a few hundred rows)
__device__ __forceinline__ double sum_rows_batched(const double *p, const size_t stride, const int r0, const int r1, const int rstep)
{
    constexpr int B = 16;
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

// ---- per-pose factorisation ------------------------------------------------------------
// lower-triangular packed index (i >= j)
__device__ __forceinline__ constexpr int lt(int i, int j) { return i * (i + 1) / 2 + j; }

// x <- L^-1 x; invd = reciprocals of L's diagonal (one division per pivot instead of one per solve)
__device__ __forceinline__ void forward_subst(const double (&Lm)[21], const double (&invd)[6], double (&x)[6])
{
    static_for<0, 6>([&](auto ic) {
        constexpr int i = VG_IDX(ic);
        double s = x[i];
        static_for<0, i>([&](auto kc) { s = fma(-Lm[lt(i, VG_IDX(kc))], x[VG_IDX(kc)], s); });
        x[i] = s * invd[i];
    });
}

__device__ __forceinline__ void backward_subst(const double (&Lm)[21], double (&x)[6])
{
    static_for<0, 6>([&](auto rc) {
        constexpr int i = 5 - VG_IDX(rc);
        double s = x[i];
        static_for<i + 1, 6>([&](auto kc) { s = fma(-Lm[lt(VG_IDX(kc), i)], x[VG_IDX(kc)], s); });
        x[i] = s / Lm[lt(i, i)];
    });
}

// ---- reduced (shared-block) system on the device ----------------------------------------------------------
// (A + D_a - S_red) delta_a = -(g_a - v_red) of the current parameter set, the candidate shared parameters
// Pi(x + delta_a) (box bounds by projection), and the scalars the host's accept / reject decision needs -- so that
// an LM iteration costs ONE host synchronisation (after the candidate's evaluation) instead of two.  One block;
// the Cholesky factorisation is right-looking with the trailing update spread over the threads.

// pose_factor's fused tail: ticket == nullptr -> the caller launches finalize_gram / reduced_solve itself
struct FusedTail {
    unsigned int *ticket;
    int n_gmax;
    double *red;
    SolveArgs solve;
};

// sm: 2 Ks^2 + 9 Ks doubles of shared memory
__device__ __forceinline__ void reduced_solve_body(int Ks, const SolveArgs &sa, const LmConsts &lm, double *sm)
{
    const int slab_n = sa.slab_n, nranks = sa.nranks;
    const double *red_cur = sa.red_cur, *slab_cur = sa.slab_cur, *sh_lo = sa.sh_lo, *sh_hi = sa.sh_hi;
    double *red_cand = sa.red_cand, *slab_cand = sa.slab_cand, *delta_a = sa.delta_a, *scale_a = sa.scale_a;
    const int *sh_off = sa.sh_off;
    // everything the serial parts touch is staged in shared memory first (a lone thread's global loads cost an L2
    // round trip each)
    double *S = sm, *As = S + Ks * Ks, *rhs = As + Ks * Ks, *da = rhs + Ks, *Ad = da + Ks, *gs = Ad + Ks, *xs = gs + Ks,
           *los = xs + Ks, *his = los + Ks, *xn = his + Ks;
    __shared__ int s_ok;
    __shared__ double s_pose[2];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double *A = red_cur + red_off_A(Ks), *ga = red_cur + red_off_g(Ks);
    const double *Sred = red_cur + red_off_S(Ks), *v = red_cur + red_off_v(Ks);
    double *out = red_cand + red_size(Ks, nranks);
    if (tid == 0) s_ok = 1;
    if (tid == 32) {
        double g = 0.0;
        for (int r = 0; r < nranks; r++) g = fmax(g, __ldcg(red_cur + red_off_gmax(Ks) + r));
        s_pose[0] = g;
        s_pose[1] = __ldcg(red_cur + red_off_fail(Ks));
    }
    for (int i = tid; i < Ks * Ks; i += nt) { const double a = __ldcg(A + i); As[i] = a; S[i] = a - __ldcg(Sred + i); }
    for (int j = tid; j < Ks; j += nt) {
        gs[j] = __ldcg(ga + j);
        rhs[j] = -(gs[j] - __ldcg(v + j));
        xs[j] = slab_cur[sh_off[j]];
        los[j] = sh_lo[j]; his[j] = sh_hi[j];
    }
    for (int i = tid; i < slab_n; i += nt) slab_cand[i] = slab_cur[i];      // constants and padding carry over
    __syncthreads();
    for (int j = tid; j < Ks; j += nt) {
        double sc;
        if (lm.init_scale) { sc = lm.jacobi_scaling ? 1.0 / (1.0 + sqrt(As[j * Ks + j])) : 1.0; scale_a[j] = sc; }
        else sc = scale_a[j];
        const double s2 = sc * sc;
        S[j * Ks + j] += fmin(fmax(s2 * As[j * Ks + j], lm.min_diag), lm.max_diag) / (lm.radius * s2);
    }
    __syncthreads();
    // right-looking Cholesky: pivot, column scaling, trailing update spread over the threads
    for (int j = 0; j < Ks; j++) {
        if (tid == 0) {
            double piv = S[j * Ks + j];
            if (!(piv > 0.0)) { s_ok = 0; piv = 1.0; }
            S[j * Ks + j] = sqrt(piv);
        }
        __syncthreads();
        const double inv = 1.0 / S[j * Ks + j];
        for (int i = j + 1 + tid; i < Ks; i += nt) S[i * Ks + j] *= inv;
        __syncthreads();
        const int m = Ks - j - 1;
        for (int e = tid; e < m * m; e += nt) {
            const int i = j + 1 + e / m, k = j + 1 + e % m;
            if (k <= i) S[i * Ks + k] = fma(-S[i * Ks + j], S[k * Ks + j], S[i * Ks + k]);
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int i = 0; i < Ks; i++) {
            double s = rhs[i];
            for (int k = 0; k < i; k++) s = fma(-S[i * Ks + k], da[k], s);
            da[i] = s / S[i * Ks + i];
        }
        for (int i = Ks - 1; i >= 0; i--) {
            double s = da[i];
            for (int k = i + 1; k < Ks; k++) s = fma(-S[k * Ks + i], da[k], s);
            da[i] = s / S[i * Ks + i];
        }
        if (!s_ok)
            for (int i = 0; i < Ks; i++) da[i] = 0.0;
    }
    __syncthreads();
    for (int j = tid; j < Ks; j += nt) {
        double t = 0.0;
        for (int k = 0; k < Ks; k++) t = fma(As[j * Ks + k], da[k], t);
        Ad[j] = t;
        delta_a[j] = da[j];
        xn[j] = fmin(fmax(xs[j] + da[j], los[j]), his[j]);      // box bounds by projection
        slab_cand[sh_off[j]] = xn[j];
        out[SOLVE_OUT + sh_off[j]] = xn[j];
    }
    for (int i = tid; i < slab_n; i += nt) {                    // the host reads the candidate from the copy after `out`
        bool shared = false;
        for (int j = 0; j < Ks; j++) shared = shared || sh_off[j] == i;
        if (!shared) out[SOLVE_OUT + i] = slab_cur[i];
    }
    __syncthreads();
    if (tid == 0) {
        double gd = 0.0, dHd = 0.0, step2 = 0.0, x2 = 0.0, gmax = s_pose[0];
        for (int j = 0; j < Ks; j++) {
            gd = fma(gs[j], da[j], gd);
            dHd = fma(da[j], Ad[j], dHd);
            x2 = fma(xs[j], xs[j], x2);
            step2 = fma(xn[j] - xs[j], xn[j] - xs[j], step2);
            // max-norm of the projected gradient
            gmax = fmax(gmax, fabs(xs[j] - fmin(fmax(xs[j] - gs[j], los[j]), his[j])));
        }
        out[0] = gmax;
        out[1] = s_pose[1];
        out[2] = (double)s_ok;
        out[3] = gd; out[4] = dHd; out[5] = step2; out[6] = x2; out[7] = 0.0;
    }
}

__global__ void __launch_bounds__(SOLVE_THREADS) reduced_solve_kernel(int Ks, SolveArgs sa, LmConsts lm)
{
    extern __shared__ double dyn_sm[];
    reduced_solve_body(Ks, sa, lm, dyn_sm);
}

// ---- Schur complement terms: S_red = sum Z^T Z, v_red = sum Z^T z -----------------------
// Tail of pose_factor: the block's poses are folded into one row of partial sums.
// thread = (entry t of the upper triangle of Z^T Z plus the Z^T z column, slice q of the block's poses); the
// slices of an entry are added in slice order afterwards, so the sum order is fixed.  sh: blockDim.x doubles.
// zs (nullable): this block's Z and z in shared memory, (6 Ks + 6) doubles per pose [Z row-major 6 x Ks | z], left
// there by the factorisation above -- the block's own rows do not come back through L2
__device__ __forceinline__ void block_gram(int n_pose, int Ks, const double *ws, double *partial, double *sh, const double *zs)
{
    const int npair = Ks * (Ks + 1) / 2 + Ks;
    const int p0 = blockIdx.x * FACTOR_POSES, p1 = min(n_pose, p0 + FACTOR_POSES);
    const int stride = pose_ws_stride(Ks);
    if (npair == 0) return;                       // no free shared parameter (poses-only refinement)
    const int Q = max(1, (int)blockDim.x / npair);
    const int per = (FACTOR_POSES + Q - 1) / Q;
    for (int t0 = 0; t0 < npair; t0 += blockDim.x) {
        const int q = Q > 1 ? threadIdx.x / npair : 0;
        const int t = Q > 1 ? threadIdx.x - q * npair : t0 + threadIdx.x;
        double s = 0.0;
        const bool active = t < npair && q < Q;
        if (active) {
            int a, b;   // b == Ks -> the z column
            if (t < Ks * (Ks + 1) / 2) {
                a = 0; int rem = t;
                while (rem >= Ks - a) { rem -= Ks - a; a++; }
                b = a + rem;
            } else {
                a = t - Ks * (Ks + 1) / 2; b = Ks;
            }
            const int q0 = p0 + q * per, q1 = min(p1, q0 + per);
            for (int p = q0; p < q1; p++) {
                if (zs) {
                    const double *Z = zs + (size_t)(p - p0) * (6 * Ks + 6), *z = Z + 6 * Ks;
#pragma unroll
                    for (int k = 0; k < 6; k++) s = fma(Z[k * Ks + a], (b < Ks) ? Z[k * Ks + b] : z[k], s);
                } else {
                    const double *w = ws + (size_t)p * stride;
                    const double *Z = w + 33, *z = w + 27;
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        const double za = __ldcg(Z + k * Ks + a);
                        const double zb = (b < Ks) ? __ldcg(Z + k * Ks + b) : __ldcg(z + k);
                        s = fma(za, zb, s);
                    }
                }
            }
        }
        if (Q > 1) {
            __syncthreads();
            sh[threadIdx.x] = s;
            __syncthreads();
            if (threadIdx.x < npair) {
                double tot = 0.0;
                for (int qq = 0; qq < Q; qq++) tot += sh[qq * npair + threadIdx.x];
                partial[(size_t)blockIdx.x * npair + threadIdx.x] = tot;
            }
            break;
        }
        if (active) partial[(size_t)blockIdx.x * npair + t] = s;
    }
}

constexpr int FIN_THREADS = 1024;    // finalize_gram: entries x row slices; the more slices, the shorter each thread's chain of loads

// sh: blockDim.x doubles.  Rows may have been written by other blocks of the same launch (fused tail): they are
// read past L1.
__device__ __forceinline__ void finalize_gram_body(int Ks, int n_blocks, const double *partial, int n_gmax,
                                                   const double *partial_gmax, double *red, int *fail_flag, int rank,
                                                   int nranks, double *sh)
{
    // thread = (entry t, slice q of the blocks); slices combined in slice order (fixed sum order)
    const int npair = Ks * (Ks + 1) / 2 + Ks;
    double *S = red + red_off_S(Ks), *v = red + red_off_v(Ks);
    const int Q = npair > 0 ? max(1, (int)blockDim.x / npair) : 1;
    const int per = (n_blocks + Q - 1) / Q;
    for (int t0 = 0; t0 < npair; t0 += blockDim.x) {
        const int q = Q > 1 ? threadIdx.x / npair : 0;
        const int t = Q > 1 ? threadIdx.x - q * npair : t0 + threadIdx.x;
        double s = 0.0;
        if (t < npair && q < Q) {
            const int b0 = q * per, b1 = min(n_blocks, b0 + per);
            s = sum_rows_batched(partial + t, npair, b0, b1, 1);
        }
        if (Q > 1) {
            sh[threadIdx.x] = s;
            __syncthreads();
            s = 0.0;
            if (threadIdx.x < npair)
                for (int qq = 0; qq < Q; qq++) s += sh[qq * npair + threadIdx.x];
        }
        if (t < npair && (Q == 1 || threadIdx.x < npair)) {
            if (t < Ks * (Ks + 1) / 2) {
                int a = 0, rem = t;
                while (rem >= Ks - a) { rem -= Ks - a; a++; }
                const int b = a + rem;
                S[a * Ks + b] = s;
                S[b * Ks + a] = s;
            } else {
                v[t - Ks * (Ks + 1) / 2] = s;
            }
        }
        if (Q > 1) break;
    }
    if (threadIdx.x >= blockDim.x - 32) {        // the last warp: max |g| over the blocks' / segments' slots
        const int lane = threadIdx.x & 31;
        double m = 0.0;
        for (int i0 = lane; i0 < n_gmax; i0 += 32 * 8) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = i0 + 32 * k < n_gmax ? __ldcg(partial_gmax + i0 + 32 * k) : 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) m = fmax(m, v[k]);
        }
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0) {
            for (int r = 0; r < nranks; r++) red[red_off_gmax(Ks) + r] = (r == rank) ? m : 0.0;
            red[red_off_fail(Ks)] = (double)atomicExch(fail_flag, 0);
        }
    }
}

__global__ void __launch_bounds__(FIN_THREADS)
finalize_gram_kernel(int Ks, int n_blocks, const double *partial, int n_gmax, const double *partial_gmax,
                     double *red, int *fail_flag, int rank, int nranks)
{
    __shared__ double sh[FIN_THREADS];
    finalize_gram_body(Ks, n_blocks, partial, n_gmax, partial_gmax, red, fail_flag, rank, nranks, sh);
}

__global__ void __launch_bounds__(FACTOR_THREADS)
pose_factor_kernel(const DatasetDesc *desc_all, int n_pose, int Ks,
                   const int *pose_start, const int *contrib_ds, const int *contrib_img,
                   double *scale, LmConsts lm, double *ws, double *partial_gmax, double *partial_gram, int *fail_flag,
                   const unsigned char *chain_mask, FusedTail ft, int z_smem)
{
    __shared__ double sh_max[FACTOR_THREADS];
    extern __shared__ double dyn_sm[];
    double *zs = z_smem ? dyn_sm : nullptr;       // (the fused tail reuses the same bytes after block_gram)
    double *zrow = zs ? zs + (size_t)(threadIdx.x / POSE_LANES) * (6 * Ks + 6) : nullptr;
    const int sub = threadIdx.x & (POSE_LANES - 1);
    const int p = blockIdx.x * FACTOR_POSES + threadIdx.x / POSE_LANES;
    double gmax = 0.0;
    // chain_mask: elements the chain kernels (vg_priors.cu) have already factorised into ws
    if (p < n_pose && !(chain_mask && chain_mask[p])) {
        double Lm[21];
#pragma unroll
        for (int i = 0; i < 21; i++) Lm[i] = 0.0;
        const int c0 = pose_start[p], c1 = pose_start[p + 1];
        for (int c = c0; c < c1; c++) {
            const DatasetDesc &d = desc_all[contrib_ds[c]];
            const double *H = d.H + (size_t)contrib_img[c] * d.ne;
            const int pc = d.pose_col, W = d.W;
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) Lm[lt(i, j)] += H[pk(pc + j, pc + i, W)];
        }
        double *w = ws + (size_t)p * pose_ws_stride(Ks);
        double lam[6], invd[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
        bool empty = true;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double ckk = Lm[lt(k, k)];
            if (ckk != 0.0) empty = false;
            double sc;
            if (lm.init_scale) {
                sc = lm.jacobi_scaling ? 1.0 / (1.0 + sqrt(ckk)) : 1.0;
                if (sub == 0) scale[(size_t)p * 6 + k] = sc;
            } else {
                sc = scale[(size_t)p * 6 + k];
            }
            const double s2 = sc * sc;
            lam[k] = fmin(fmax(s2 * ckk, lm.min_diag), lm.max_diag) / (lm.radius * s2);
        }
        bool ok = true;
        if (empty) {
            // a pose nothing observes: identity factor, zero step
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) Lm[lt(i, j)] = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++) lam[k] = 0.0;
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) Lm[lt(k, k)] += lam[k];
            // Cholesky, in place (compile-time indices: a loop left rolled would send Lm to local memory)
            static_for<0, 6>([&](auto jc) {
                constexpr int j = VG_IDX(jc);
                double s = Lm[lt(j, j)];
                static_for<0, j>([&](auto kc) { s = fma(-Lm[lt(j, VG_IDX(kc))], Lm[lt(j, VG_IDX(kc))], s); });
                if (!(s > 0.0)) { ok = false; s = 1.0; }
                s = sqrt(s);
                Lm[lt(j, j)] = s;
                const double inv = 1.0 / s;
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
#pragma unroll
            for (int i = 0; i < 21; i++) w[i] = Lm[i];
#pragma unroll
            for (int k = 0; k < 6; k++) w[21 + k] = lam[k];
        }
        // the group's lanes take the columns: col < Ks -> column col of E^T (Z = L^-1 E^T), col == Ks -> the gradient
        for (int col = sub; col <= Ks; col += POSE_LANES) {
            double e[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int c = c0; c < c1; c++) {
                const DatasetDesc &d = desc_all[contrib_ds[c]];
                const double *H = d.H + (size_t)contrib_img[c] * d.ne;
                const int pc = d.pose_col, W = d.W;
                int lc = W - 1;                         // the residual column
                if (col < Ks) {
                    lc = -1;
                    for (int q = 0; q < d.n_sl; q++)
                        if (d.sl_idx[q] == col) lc = d.sl_col[q];
                    if (lc < 0) continue;               // this dataset does not touch the shared column
                }
#pragma unroll
                for (int i = 0; i < 6; i++) e[i] += H[pks(lc, pc + i, W)];
            }
            if (col == Ks) {
#pragma unroll
                for (int k = 0; k < 6; k++) gmax = fmax(gmax, fabs(e[k]));
            }
            forward_subst(Lm, invd, e);
            if (col == Ks) {
#pragma unroll
                for (int k = 0; k < 6; k++) { w[27 + k] = e[k]; if (zrow) zrow[6 * Ks + k] = e[k]; }   // z = L^-1 b
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) { w[33 + k * Ks + col] = e[k]; if (zrow) zrow[k * Ks + col] = e[k]; }
            }
        }
    } else if (zrow && p < n_pose) {
        // an element the chain kernels factorised: its rows of ws are complete in global memory
        const double *w = ws + (size_t)p * pose_ws_stride(Ks);
        for (int i = sub; i < 6 * Ks + 6; i += POSE_LANES) zrow[i] = i < 6 * Ks ? __ldcg(w + 33 + i) : __ldcg(w + 27 + (i - 6 * Ks));
    }
    sh_max[threadIdx.x] = gmax;
    __syncthreads();          // also: this block's Z, z are written (read back below by other threads of the block)
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh_max[threadIdx.x] = fmax(sh_max[threadIdx.x], sh_max[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial_gmax[blockIdx.x] = sh_max[0];
    block_gram(n_pose, Ks, ws, partial_gram, sh_max, zs);
    if (!ft.ticket) return;
    // Fused tail (one rank): the block that finishes last combines every block's row and solves the reduced
    // system, instead of two more single-block launches after this one.
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ft.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ft.ticket = 0;
    finalize_gram_body(Ks, gridDim.x, partial_gram, ft.n_gmax, partial_gmax, ft.red, fail_flag, 0, 1, sh_max);
    __threadfence();
    __syncthreads();
    reduced_solve_body(Ks, ft.solve, lm, dyn_sm);
}

// warp q < 3 sums quantity q (model decrease, step^2, x^2): lane l takes rows l, l + 32, ..., the lanes' sums are
// added in lane order (fixed order -> reproducible)
__device__ __forceinline__ void finalize_backsub_body(int Ks, int n_blocks, const double *partial, double *red)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= 3) return;
    const double s = sum_rows_batched(partial + q, 3, lane, n_blocks, 32);
    double tot = 0.0;
    for (int l = 0; l < 32; l++) tot += __shfl_sync(0xffffffffu, s, l);
    if (lane == 0) red[red_off_model(Ks) + q] = tot;
}

__global__ void __launch_bounds__(96) finalize_backsub_kernel(int Ks, int n_blocks, const double *partial, double *red)
{
    finalize_backsub_body(Ks, n_blocks, partial, red);
}

// ---- back substitution -------------------------------------------------------------------
__global__ void __launch_bounds__(POSE_THREADS)
pose_backsub_kernel(int n_pose, int Ks, const double *delta_a, const double *const *seq_cur,
                    double *const *seq_cand, const int *pose_seq, const int *pose_local,
                    const double *ws, double *partial, const unsigned char *chain_mask, double *chain_w,
                    unsigned int *ticket, double *red)
{
    // lane k < 6 of a pose's group: w_k = z_k + Z_k . delta_a (row k of Z is contiguous), then its own component of
    // delta = -L^-T w (the triangular solve is repeated by the six lanes: the others' w come through shuffles)
    __shared__ double sh[3][POSE_THREADS];
    const int sub = threadIdx.x & (POSE_LANES - 1), lane = threadIdx.x & 31;
    const int p = blockIdx.x * BACK_POSES + threadIdx.x / POSE_LANES;
    double m = 0.0, st2 = 0.0, x2 = 0.0, wk = 0.0;
    const bool active = p < n_pose && sub < 6;
    const bool chained = active && chain_mask && chain_mask[p];
    const double *w = ws + (size_t)(p < n_pose ? p : 0) * pose_ws_stride(Ks);
    if (active) {
        const double *Zk = w + 33 + sub * Ks;
        double s = w[27 + sub];
        for (int a = 0; a < Ks; a++) s = fma(Zk[a], delta_a[a], s);
        wk = s;
        m = -0.5 * s * s;
        // coupled elements: the triangular solve runs along the segment (chain_backsub_kernel, vg_priors.cu)
        if (chained) chain_w[(size_t)p * 6 + sub] = s;
    }
    double wv[6];
#pragma unroll
    for (int j = 0; j < 6; j++) wv[j] = __shfl_sync(0xffffffffu, wk, (lane & ~(POSE_LANES - 1)) + j);
    if (active && !chained) {
        double Lm[21];
#pragma unroll
        for (int i = 0; i < 21; i++) Lm[i] = w[i];
        backward_subst(Lm, wv);   // wv = L^-T w ; delta = -wv
        double yk = 0.0;
#pragma unroll
        for (int j = 0; j < 6; j++) yk = (j == sub) ? wv[j] : yk;
        const double dlt = -yk;
        const double x = seq_cur[pose_seq[p]][(size_t)pose_local[p] * 6 + sub];
        seq_cand[pose_seq[p]][(size_t)pose_local[p] * 6 + sub] = x + dlt;
        m = fma(-0.5 * w[21 + sub] * dlt, dlt, m);
        st2 = dlt * dlt;
        x2 = x * x;
    }
    sh[0][threadIdx.x] = m; sh[1][threadIdx.x] = st2; sh[2][threadIdx.x] = x2;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
#pragma unroll
            for (int q = 0; q < 3; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) partial[(size_t)blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
    if (!ticket) return;
    // fused finalize (no chain segments): the last block adds the rows up, as finalize_backsub_kernel would
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0;
    finalize_backsub_body(Ks, gridDim.x, partial, red);
}



// one block: a segment of the reduction buffer summed across the ranks over peer memory (vg_peer.cuh)
__global__ void __launch_bounds__(256) peer_exchange_kernel(double *buf, int count, PeerCtx pc)
{
    __shared__ double scratch[2048];
    peer_allreduce(buf, count, pc, scratch, 2048);
}

// the second half alone: the sum of an exchange an evaluation kernel posted earlier
__global__ void __launch_bounds__(256) peer_collect_kernel(double *buf, int count, PeerCtx pc, unsigned long long *done, int post)
{
    __shared__ double scratch[2048];
    if (post) peer_post(buf, count, pc);          // (an evaluation that left the posting to its successor)
    peer_collect(buf, count, pc, scratch, 2048);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *done = pc.epoch;
}

}  // namespace

// Tables of the shared-block reduction (fused into the evaluation kernel): offsets of each dataset's sums
// region and, per entry of the reduced system, its sources.
void build_finalize_tables(const DatasetDesc *h_desc, const int *grids, int n_ds, int Ks, std::vector<int> &offsets,
                           std::vector<FinOut> &outs, std::vector<FinSrc> &srcs, size_t *partial_doubles)
{
    offsets.assign(n_ds + 1, 0);
    size_t off = 0;
    for (int ds = 0; ds < n_ds; ds++) { offsets[ds] = (int)off; off += (size_t)grids[ds] * h_desc[ds].ne; }
    offsets[n_ds] = (int)off;
    *partial_doubles = off + 8;
    const int n_out = Ks * (Ks + 1) / 2 + Ks + 1;
    std::vector<std::vector<FinSrc>> per(n_out);
    auto a_index = [&](int i, int j) { if (i > j) { const int t = i; i = j; j = t; } return i * Ks - i * (i - 1) / 2 + (j - i); };
    for (int ds = 0; ds < n_ds; ds++) {
        const DatasetDesc &d = h_desc[ds];
        if (d.n_img <= 0) continue;
        int e = 0;
        for (int a = 0; a < d.W; a++)
            for (int b = a; b < d.W; b++, e++) {
                const int ka = d.kind[a], kb = d.kind[b];
                int o = -1;
                if (ka == COL_SHARED && kb == COL_SHARED) o = a_index(d.idx[a], d.idx[b]);
                else if (ka == COL_SHARED && kb == COL_RESID) o = Ks * (Ks + 1) / 2 + d.idx[a];
                else if (ka == COL_RESID && kb == COL_RESID) o = n_out - 1;
                if (o >= 0) per[o].push_back(FinSrc{offsets[ds], d.ne, grids[ds], e});
            }
    }
    outs.clear(); srcs.clear();
    int o = 0;
    for (int i = 0; i < Ks; i++)
        for (int j = i; j < Ks; j++, o++) {
            FinOut f{red_off_A(Ks) + i * Ks + j, i == j ? -1 : red_off_A(Ks) + j * Ks + i, 1.0, (int)srcs.size(), 0};
            for (auto &x : per[o]) srcs.push_back(x);
            f.src_end = (int)srcs.size();
            outs.push_back(f);
        }
    for (int i = 0; i < Ks; i++, o++) {
        FinOut f{red_off_g(Ks) + i, -1, 1.0, (int)srcs.size(), 0};
        for (auto &x : per[o]) srcs.push_back(x);
        f.src_end = (int)srcs.size();
        outs.push_back(f);
    }
    FinOut f{red_off_cost(Ks), -1, 0.5, (int)srcs.size(), 0};
    for (auto &x : per[o]) srcs.push_back(x);
    f.src_end = (int)srcs.size();
    outs.push_back(f);
}

int pose_factor_blocks(int n_pose) { return (n_pose + FACTOR_POSES - 1) / FACTOR_POSES; }
int pose_backsub_blocks(int n_pose) { return (n_pose + BACK_POSES - 1) / BACK_POSES; }

size_t pose_scratch(int n_pose, int Ks, int n_seg)
{
    const size_t nb_pose = (n_pose + FACTOR_POSES - 1) / FACTOR_POSES;
    const size_t npair = (size_t)Ks * (Ks + 1) / 2 + Ks;
    return nb_pose * 4 + nb_pose * npair + 4 * (size_t)n_seg + 16;
}

cudaError_t launch_pose_schur(const DatasetDesc *d_desc, int n_pose, int Ks,
                              const int *pose_start, const int *contrib_ds, const int *contrib_img,
                              double *scale, LmConsts lm, double *ws, double *partial, size_t partial_doubles,
                              double *red, int *fail_flag, int rank, int nranks, SolverLaunch sl,
                              const unsigned char *chain_mask, int n_seg, const SolveArgs *fused, unsigned int *ticket)
{
    // partial: [max |g| of each block (nb_pose) | of each chain segment (n_seg, written by chain_factor) | gram rows]
    const int nb_pose = (n_pose + FACTOR_POSES - 1) / FACTOR_POSES;
    if (pose_scratch(n_pose, Ks, n_seg) > partial_doubles) return cudaErrorInvalidValue;
    double *p_gmax = partial;
    double *p_gram = partial + nb_pose + n_seg;
    FusedTail ft;
    memset(&ft, 0, sizeof ft);
    size_t smem = 0;
    if (fused && ticket && n_pose > 0 && nranks == 1) {
        // one rank: the last block of pose_factor combines the rows and solves the reduced system itself
        smem = sizeof(double) * (2 * (size_t)Ks * Ks + 9 * (size_t)Ks + 1);
        if (smem > 40 * 1024) return cudaErrorInvalidValue;
        ft.ticket = ticket; ft.n_gmax = nb_pose + n_seg; ft.red = red; ft.solve = *fused;
    }
    // the block's Z, z rows in shared memory for its Schur terms (when they fit the default 48 KB with the rest)
    const size_t zbytes = sizeof(double) * FACTOR_POSES * (6 * (size_t)Ks + 6);
    const int z_smem = zbytes <= 40 * 1024 ? 1 : 0;
    if (z_smem && zbytes > smem) smem = zbytes;
    if (n_pose > 0) {
        pose_factor_kernel<<<nb_pose, FACTOR_THREADS, smem, sl.stream>>>(d_desc, n_pose, Ks, pose_start, contrib_ds,
                                                                        contrib_img, scale, lm, ws, p_gmax, p_gram, fail_flag,
                                                                        chain_mask, ft, z_smem);
        if (sl.launches) count_launch(sl.launches);
    }
    if (ft.ticket) return cudaGetLastError();
    finalize_gram_kernel<<<1, FIN_THREADS, 0, sl.stream>>>(Ks, n_pose > 0 ? nb_pose : 0, p_gram, n_pose > 0 ? nb_pose + n_seg : 0, p_gmax,
                                                   red, fail_flag, rank, nranks);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

cudaError_t launch_pose_backsub(int n_pose, int Ks, const double *delta_a,
                                const double *const *seq_cur, double *const *seq_cand,
                                const int *pose_seq, const int *pose_local,
                                const double *ws, double *partial, size_t partial_doubles, double *red,
                                SolverLaunch sl, const unsigned char *chain_mask, double *chain_w, unsigned int *ticket)
{
    const int nb = (n_pose + BACK_POSES - 1) / BACK_POSES;
    if ((size_t)nb * 3 > partial_doubles) return cudaErrorInvalidValue;
    if (n_pose > 0) {
        pose_backsub_kernel<<<nb, POSE_THREADS, 0, sl.stream>>>(n_pose, Ks, delta_a, seq_cur, seq_cand, pose_seq,
                                                                pose_local, ws, partial, chain_mask, chain_w, ticket, red);
        if (sl.launches) count_launch(sl.launches);
    }
    return cudaGetLastError();
}

// sums the rows the back-substitution kernels left: one per pose_backsub block, then one per chain segment
cudaError_t launch_finalize_backsub(int Ks, int n_rows, const double *partial, double *red, SolverLaunch sl)
{
    finalize_backsub_kernel<<<1, 96, 0, sl.stream>>>(Ks, n_rows, partial, red);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

cudaError_t launch_reduced_solve(int Ks, const SolveArgs &sa, LmConsts lm, SolverLaunch sl)
{
    const size_t smem = sizeof(double) * (2 * (size_t)Ks * Ks + 9 * (size_t)Ks + 1);
    if (smem > 40 * 1024) return cudaErrorInvalidValue;       // Ks <= 48
    reduced_solve_kernel<<<1, SOLVE_THREADS, smem, sl.stream>>>(Ks, sa, lm);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

cudaError_t launch_peer_exchange(double *buf, int count, const PeerCtx &pc, SolverLaunch sl)
{
    if (count > PEER_SLOT_DOUBLES) return cudaErrorInvalidValue;
    peer_exchange_kernel<<<1, 256, 0, sl.stream>>>(buf, count, pc);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

cudaError_t launch_peer_collect(double *buf, int count, const PeerCtx &pc, unsigned long long *done, bool post, SolverLaunch sl)
{
    if (count > PEER_SLOT_DOUBLES) return cudaErrorInvalidValue;
    peer_collect_kernel<<<1, 256, 0, sl.stream>>>(buf, count, pc, done, post ? 1 : 0);
    if (sl.launches) count_launch(sl.launches);
    return cudaGetLastError();
}

}  // namespace vg
