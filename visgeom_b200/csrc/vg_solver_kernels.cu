// vg_solver_kernels.cu -- see vg_solver_kernels.cuh.  All reductions are two-stage
// (per-block partials in a fixed layout, then one block summing them in index order),
// so results are bit-reproducible run to run for a fixed problem.
#include "vg_solver_kernels.cuh"

namespace vg {

namespace {

__host__ __device__ inline int pk(int a, int b, int W) { return a * W - a * (a - 1) / 2 + (b - a); }
__device__ __forceinline__ int pks(int a, int b, int W) { return a <= b ? pk(a, b, W) : pk(b, a, W); }

constexpr int ACC_THREADS = 256;
constexpr int POSE_THREADS = 128;
constexpr int GRAM_POSES = 64;       // poses per block in gram_reduce
constexpr int GRAM_THREADS = 224;    // >= Ks(Ks+1)/2 + Ks for Ks <= 18; larger Ks loops

// ---- A, g_a, cost: fold the per-CTA block sums of the evaluation kernel -------------------
// one block; datasets are folded in sequentially so that two datasets sharing a
// camera add into the same entries in a fixed order
__global__ void __launch_bounds__(ACC_THREADS)
finalize_shared_kernel(const DatasetDesc *desc_all, int n_ds, int Ks, const double *partial,
                       const int *partial_off, const int *n_blocks, double *red)
{
    double *A = red + red_off_A(Ks), *g = red + red_off_g(Ks), *cost = red + red_off_cost(Ks);
    for (int i = threadIdx.x; i < red_off_model(Ks); i += blockDim.x) red[i] = 0.0;   // A, g_a, cost
    __syncthreads();
    for (int ds = 0; ds < n_ds; ds++) {
        const DatasetDesc &d = desc_all[ds];
        const double *pp = partial + partial_off[ds];
        for (int e = threadIdx.x; e < d.ne; e += blockDim.x) {
            // packed index -> (a,b)
            int a = 0, rem = e;
            while (rem >= d.W - a) { rem -= d.W - a; a++; }
            const int b = a + rem;
            const int ka = d.kind[a], kb = d.kind[b];
            if (ka == COL_CONST || kb == COL_CONST || ka == COL_POSE || kb == COL_POSE) continue;
            double s = 0.0;
            for (int blk = 0; blk < n_blocks[ds]; blk++) s += pp[(size_t)blk * d.ne + e];
            if (ka == COL_SHARED && kb == COL_SHARED) {
                A[d.idx[a] * Ks + d.idx[b]] += s;
                if (a != b) A[d.idx[b] * Ks + d.idx[a]] += s;
            } else if (ka == COL_SHARED && kb == COL_RESID) {
                g[d.idx[a]] += s;
            } else if (ka == COL_RESID && kb == COL_RESID) {
                cost[0] += 0.5 * s;
            }
        }
        __syncthreads();
    }
}

// ---- per-pose factorisation ------------------------------------------------------------
// lower-triangular packed index (i >= j)
__device__ __forceinline__ constexpr int lt(int i, int j) { return i * (i + 1) / 2 + j; }

__device__ __forceinline__ void forward_subst(const double (&Lm)[21], double (&x)[6])
{
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < i; k++) s = fma(-Lm[lt(i, k)], x[k], s);
        x[i] = s / Lm[lt(i, i)];
    }
}

__device__ __forceinline__ void backward_subst(const double (&Lm)[21], double (&x)[6])
{
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = x[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s = fma(-Lm[lt(k, i)], x[k], s);
        x[i] = s / Lm[lt(i, i)];
    }
}

__global__ void __launch_bounds__(POSE_THREADS)
pose_factor_kernel(const DatasetDesc *desc_all, int n_pose, int Ks,
                   const int *pose_start, const int *contrib_ds, const int *contrib_img,
                   double *scale, LmConsts lm, double *ws, double *partial_gmax, int *fail_flag)
{
    __shared__ double sh_max[POSE_THREADS];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double gmax = 0.0;
    if (p < n_pose) {
        double C[21], b[6];
#pragma unroll
        for (int i = 0; i < 21; i++) C[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = 0.0;
        const int c0 = pose_start[p], c1 = pose_start[p + 1];
        for (int c = c0; c < c1; c++) {
            const DatasetDesc &d = desc_all[contrib_ds[c]];
            const double *H = d.H + (size_t)contrib_img[c] * d.ne;
            const int pc = d.pose_col, W = d.W;
#pragma unroll
            for (int i = 0; i < 6; i++) {
#pragma unroll
                for (int j = 0; j <= i; j++) C[lt(i, j)] += H[pk(pc + j, pc + i, W)];
                b[i] += H[pk(pc + i, W - 1, W)];
            }
        }
        double *w = ws + (size_t)p * pose_ws_stride(Ks);
        double Lm[21], lam[6];
        bool empty = true;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double ckk = C[lt(k, k)];
            if (ckk != 0.0) empty = false;
            double sc;
            if (lm.init_scale) {
                sc = lm.jacobi_scaling ? 1.0 / (1.0 + sqrt(ckk)) : 1.0;
                scale[(size_t)p * 6 + k] = sc;
            } else {
                sc = scale[(size_t)p * 6 + k];
            }
            const double s2 = sc * sc;
            lam[k] = fmin(fmax(s2 * ckk, lm.min_diag), lm.max_diag) / (lm.radius * s2);
            gmax = fmax(gmax, fabs(b[k]));
        }
#pragma unroll
        for (int i = 0; i < 21; i++) Lm[i] = C[i];
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
            // Cholesky, in place
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double s = Lm[lt(j, j)];
#pragma unroll
                for (int k = 0; k < j; k++) s = fma(-Lm[lt(j, k)], Lm[lt(j, k)], s);
                if (!(s > 0.0)) { ok = false; s = 1.0; }
                s = sqrt(s);
                Lm[lt(j, j)] = s;
                const double inv = 1.0 / s;
#pragma unroll
                for (int i = j + 1; i < 6; i++) {
                    double t = Lm[lt(i, j)];
#pragma unroll
                    for (int k = 0; k < j; k++) t = fma(-Lm[lt(i, k)], Lm[lt(j, k)], t);
                    Lm[lt(i, j)] = t * inv;
                }
            }
        }
        if (!ok) atomicExch(fail_flag, 1);
        forward_subst(Lm, b);   // z = L^-1 b
#pragma unroll
        for (int i = 0; i < 21; i++) w[i] = Lm[i];
#pragma unroll
        for (int k = 0; k < 6; k++) { w[21 + k] = lam[k]; w[27 + k] = b[k]; }
        double *Z = w + 33;     // 6 x Ks, row-major
        for (int i = 0; i < 6 * Ks; i++) Z[i] = 0.0;
        for (int c = c0; c < c1; c++) {
            const DatasetDesc &d = desc_all[contrib_ds[c]];
            const double *H = d.H + (size_t)contrib_img[c] * d.ne;
            const int pc = d.pose_col, W = d.W;
            for (int q = 0; q < d.n_sl; q++) {
                const int col = d.sl_col[q], sidx = d.sl_idx[q];
                double e[6];
#pragma unroll
                for (int i = 0; i < 6; i++) e[i] = H[pks(col, pc + i, W)];
                forward_subst(Lm, e);
#pragma unroll
                for (int i = 0; i < 6; i++) Z[i * Ks + sidx] += e[i];
            }
        }
    }
    sh_max[threadIdx.x] = gmax;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh_max[threadIdx.x] = fmax(sh_max[threadIdx.x], sh_max[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial_gmax[blockIdx.x] = sh_max[0];
}

// ---- Schur complement terms: S_red = sum Z^T Z, v_red = sum Z^T z -----------------------
__global__ void __launch_bounds__(GRAM_THREADS)
gram_kernel(int n_pose, int Ks, const double *ws, double *partial)
{
    const int npair = Ks * (Ks + 1) / 2 + Ks;
    const int p0 = blockIdx.x * GRAM_POSES, p1 = min(n_pose, p0 + GRAM_POSES);
    const int stride = pose_ws_stride(Ks);
    for (int t = threadIdx.x; t < npair; t += blockDim.x) {
        int a, b;   // b == Ks -> the z column
        if (t < Ks * (Ks + 1) / 2) {
            a = 0; int rem = t;
            while (rem >= Ks - a) { rem -= Ks - a; a++; }
            b = a + rem;
        } else {
            a = t - Ks * (Ks + 1) / 2; b = Ks;
        }
        double s = 0.0;
        for (int p = p0; p < p1; p++) {
            const double *w = ws + (size_t)p * stride;
            const double *Z = w + 33, *z = w + 27;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const double za = Z[k * Ks + a];
                const double zb = (b < Ks) ? Z[k * Ks + b] : z[k];
                s = fma(za, zb, s);
            }
        }
        partial[(size_t)blockIdx.x * npair + t] = s;
    }
}

__global__ void __launch_bounds__(256)
finalize_gram_kernel(int Ks, int n_blocks, const double *partial, int n_gmax, const double *partial_gmax,
                     double *red, int *fail_flag, int rank, int nranks)
{
    const int npair = Ks * (Ks + 1) / 2 + Ks;
    double *S = red + red_off_S(Ks), *v = red + red_off_v(Ks);
    for (int t = threadIdx.x; t < npair; t += blockDim.x) {
        double s = 0.0;
        for (int blk = 0; blk < n_blocks; blk++) s += partial[(size_t)blk * npair + t];
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
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int i = 0; i < n_gmax; i++) m = fmax(m, partial_gmax[i]);
        for (int r = 0; r < nranks; r++) red[red_off_gmax(Ks) + r] = (r == rank) ? m : 0.0;
        red[red_off_fail(Ks)] = (double)(*fail_flag);
        *fail_flag = 0;
    }
}

// ---- back substitution -------------------------------------------------------------------
__global__ void __launch_bounds__(POSE_THREADS)
pose_backsub_kernel(int n_pose, int Ks, const double *delta_a, const double *const *seq_cur,
                    double *const *seq_cand, const int *pose_seq, const int *pose_local,
                    const double *ws, double *partial)
{
    __shared__ double sh[3][POSE_THREADS];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0, st2 = 0.0, x2 = 0.0;
    if (p < n_pose) {
        const double *w = ws + (size_t)p * pose_ws_stride(Ks);
        double Lm[21], wv[6];
#pragma unroll
        for (int i = 0; i < 21; i++) Lm[i] = w[i];
        const double *Z = w + 33;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            double s = w[27 + k];
            for (int a = 0; a < Ks; a++) s = fma(Z[k * Ks + a], delta_a[a], s);
            wv[k] = s;
            m = fma(-0.5 * s, s, m);
        }
        backward_subst(Lm, wv);   // wv = L^-T w ; delta = -wv
        const double *cur = seq_cur[pose_seq[p]] + (size_t)pose_local[p] * 6;
        double *cand = seq_cand[pose_seq[p]] + (size_t)pose_local[p] * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double dlt = -wv[k];
            const double x = cur[k];
            cand[k] = x + dlt;
            m = fma(-0.5 * w[21 + k] * dlt, dlt, m);
            st2 = fma(dlt, dlt, st2);
            x2 = fma(x, x, x2);
        }
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
}

__global__ void finalize_backsub_kernel(int Ks, int n_blocks, const double *partial, double *red)
{
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int blk = 0; blk < n_blocks; blk++) s += partial[(size_t)blk * 3 + threadIdx.x];
        red[red_off_model(Ks) + threadIdx.x] = s;
    }
}

}  // namespace

// The fused evaluation kernel leaves, per dataset, one row of ne block sums per persistent
// CTA (EvalArgs::cta_partial).  h_tab = [offset(ds) | rows(ds)] in doubles / rows; d_tab is
// the same table on the device.
void shared_partial_table(const DatasetDesc *h_desc, const int *grids, int n_ds, int *h_tab)
{
    size_t off = 0;
    for (int ds = 0; ds < n_ds; ds++) {
        h_tab[ds] = (int)off;
        h_tab[n_ds + ds] = h_desc[ds].n_img > 0 ? grids[ds] : 0;
        off += (size_t)grids[ds] * h_desc[ds].ne;
    }
}

size_t shared_partial_doubles(const DatasetDesc *h_desc, const int *grids, int n_ds)
{
    size_t off = 0;
    for (int ds = 0; ds < n_ds; ds++) off += (size_t)grids[ds] * h_desc[ds].ne;
    return off + 8;
}

cudaError_t launch_finalize_shared(const DatasetDesc *d_desc, int n_ds, int Ks, const double *partial,
                                   const int *d_tab, double *red, SolverLaunch sl)
{
    finalize_shared_kernel<<<1, ACC_THREADS, 0, sl.stream>>>(d_desc, n_ds, Ks, partial, d_tab, d_tab + n_ds, red);
    if (sl.launches) (*sl.launches)++;
    return cudaGetLastError();
}

size_t pose_scratch(int n_pose, int Ks)
{
    const size_t nb_pose = (n_pose + POSE_THREADS - 1) / POSE_THREADS;
    const size_t nb_gram = (n_pose + GRAM_POSES - 1) / GRAM_POSES;
    const size_t npair = (size_t)Ks * (Ks + 1) / 2 + Ks;
    return nb_pose * 4 + nb_gram * npair + 16;
}

cudaError_t launch_pose_schur(const DatasetDesc *d_desc, int n_pose, int Ks,
                              const int *pose_start, const int *contrib_ds, const int *contrib_img,
                              double *scale, LmConsts lm, double *ws, double *partial, size_t partial_doubles,
                              double *red, int *fail_flag, int rank, int nranks, SolverLaunch sl)
{
    const int nb_pose = (n_pose + POSE_THREADS - 1) / POSE_THREADS;
    const int nb_gram = (n_pose + GRAM_POSES - 1) / GRAM_POSES;
    if (pose_scratch(n_pose, Ks) > partial_doubles) return cudaErrorInvalidValue;
    double *p_gmax = partial;
    double *p_gram = partial + nb_pose;
    if (n_pose > 0) {
        pose_factor_kernel<<<nb_pose, POSE_THREADS, 0, sl.stream>>>(d_desc, n_pose, Ks, pose_start, contrib_ds,
                                                                   contrib_img, scale, lm, ws, p_gmax, fail_flag);
        gram_kernel<<<nb_gram, GRAM_THREADS, 0, sl.stream>>>(n_pose, Ks, ws, p_gram);
        if (sl.launches) (*sl.launches) += 2;
    }
    finalize_gram_kernel<<<1, 256, 0, sl.stream>>>(Ks, n_pose > 0 ? nb_gram : 0, p_gram, n_pose > 0 ? nb_pose : 0, p_gmax,
                                                   red, fail_flag, rank, nranks);
    if (sl.launches) (*sl.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_pose_backsub(int n_pose, int Ks, const double *delta_a,
                                const double *const *seq_cur, double *const *seq_cand,
                                const int *pose_seq, const int *pose_local,
                                const double *ws, double *partial, size_t partial_doubles, double *red,
                                SolverLaunch sl)
{
    const int nb = (n_pose + POSE_THREADS - 1) / POSE_THREADS;
    if ((size_t)nb * 3 > partial_doubles) return cudaErrorInvalidValue;
    if (n_pose > 0) {
        pose_backsub_kernel<<<nb, POSE_THREADS, 0, sl.stream>>>(n_pose, Ks, delta_a, seq_cur, seq_cand, pose_seq,
                                                                pose_local, ws, partial);
        if (sl.launches) (*sl.launches)++;
    }
    finalize_backsub_kernel<<<1, 32, 0, sl.stream>>>(Ks, nb, partial, red);
    if (sl.launches) (*sl.launches)++;
    return cudaGetLastError();
}

}  // namespace vg
