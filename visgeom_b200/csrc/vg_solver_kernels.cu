// vg_solver_kernels.cu -- see vg_solver_kernels.cuh.  All reductions are two-stage
// (per-block partials in a fixed layout, then one block summing them in index order),
// so results are bit-reproducible run to run for a fixed problem.
#include "vg_solver_kernels.cuh"

#include <vector>

namespace vg {

namespace {

__host__ __device__ inline int pk(int a, int b, int W) { return a * W - a * (a - 1) / 2 + (b - a); }
__device__ __forceinline__ int pks(int a, int b, int W) { return a <= b ? pk(a, b, W) : pk(b, a, W); }

constexpr int ACC_THREADS = 256;
constexpr int POSE_THREADS = 128;     // pose_backsub
constexpr int FACTOR_THREADS = 64;    // pose_factor: one pose per thread, >= one block per SM at 10 000 poses

// ---- per-pose factorisation ------------------------------------------------------------
// lower-triangular packed index (i >= j)
__device__ __forceinline__ constexpr int lt(int i, int j) { return i * (i + 1) / 2 + j; }

// x <- L^-1 x; invd = reciprocals of L's diagonal (one division per pivot instead of one per solve)
__device__ __forceinline__ void forward_subst(const double (&Lm)[21], const double (&invd)[6], double (&x)[6])
{
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < i; k++) s = fma(-Lm[lt(i, k)], x[k], s);
        x[i] = s * invd[i];
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

// ---- Schur complement terms: S_red = sum Z^T Z, v_red = sum Z^T z -----------------------
// Tail of pose_factor: the block's poses (one per thread) are folded into one row of partial sums.
// thread = (entry t of the upper triangle of Z^T Z plus the Z^T z column, slice q of the block's poses); the
// slices of an entry are added in slice order afterwards, so the sum order is fixed.  sh: blockDim.x doubles.
__device__ __forceinline__ void block_gram(int n_pose, int Ks, const double *ws, double *partial, double *sh)
{
    const int npair = Ks * (Ks + 1) / 2 + Ks;
    const int p0 = blockIdx.x * blockDim.x, p1 = min(n_pose, p0 + (int)blockDim.x);
    const int stride = pose_ws_stride(Ks);
    const int Q = max(1, (int)blockDim.x / npair);
    const int per = ((int)blockDim.x + Q - 1) / Q;
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

__global__ void __launch_bounds__(FACTOR_THREADS)
pose_factor_kernel(const DatasetDesc *desc_all, int n_pose, int Ks,
                   const int *pose_start, const int *contrib_ds, const int *contrib_img,
                   double *scale, LmConsts lm, double *ws, double *partial_gmax, double *partial_gram, int *fail_flag,
                   const unsigned char *chain_mask)
{
    __shared__ double sh_max[FACTOR_THREADS];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double gmax = 0.0;
    // chain_mask: elements the chain kernels (vg_priors.cu) have already factorised into ws
    if (p < n_pose && !(chain_mask && chain_mask[p])) {
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
        double Lm[21], lam[6], invd[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
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
                invd[j] = inv;
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
        forward_subst(Lm, invd, b);   // z = L^-1 b
#pragma unroll
        for (int i = 0; i < 21; i++) w[i] = Lm[i];
#pragma unroll
        for (int k = 0; k < 6; k++) { w[21 + k] = lam[k]; w[27 + k] = b[k]; }
        double *Z = w + 33;     // 6 x Ks, row-major
        const bool single = (c1 - c0 == 1);      // one image observes this pose: every Z entry is written once
        if (!single)
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
                forward_subst(Lm, invd, e);
                if (single) {
#pragma unroll
                    for (int i = 0; i < 6; i++) Z[i * Ks + sidx] = e[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 6; i++) Z[i * Ks + sidx] += e[i];
                }
            }
            if (single)      // shared columns this dataset does not touch
                for (int sidx = 0; sidx < Ks; sidx++) {
                    bool touched = false;
                    for (int q = 0; q < d.n_sl; q++) touched = touched || d.sl_idx[q] == sidx;
                    if (!touched)
                        for (int i = 0; i < 6; i++) Z[i * Ks + sidx] = 0.0;
                }
        }
    }
    sh_max[threadIdx.x] = gmax;
    __syncthreads();          // also: this block's Z, z are written (read back below by other threads of the block)
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh_max[threadIdx.x] = fmax(sh_max[threadIdx.x], sh_max[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial_gmax[blockIdx.x] = sh_max[0];
    block_gram(n_pose, Ks, ws, partial_gram, sh_max);
}

__global__ void __launch_bounds__(256)
finalize_gram_kernel(int Ks, int n_blocks, const double *partial, int n_gmax, const double *partial_gmax,
                     double *red, int *fail_flag, int rank, int nranks)
{
    // thread = (entry t, slice q of the blocks); slices combined in slice order (fixed sum order)
    __shared__ double sh[256];
    const int npair = Ks * (Ks + 1) / 2 + Ks;
    double *S = red + red_off_S(Ks), *v = red + red_off_v(Ks);
    const int Q = max(1, (int)blockDim.x / npair);
    const int per = (n_blocks + Q - 1) / Q;
    for (int t0 = 0; t0 < npair; t0 += blockDim.x) {
        const int q = Q > 1 ? threadIdx.x / npair : 0;
        const int t = Q > 1 ? threadIdx.x - q * npair : t0 + threadIdx.x;
        double s = 0.0;
        if (t < npair && q < Q) {
            const int b0 = q * per, b1 = min(n_blocks, b0 + per);
            for (int blk = b0; blk < b1; blk++) s += partial[(size_t)blk * npair + t];
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
    if (threadIdx.x == 32) {
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
                    const double *ws, double *partial, const unsigned char *chain_mask, double *chain_w)
{
    __shared__ double sh[3][POSE_THREADS];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0, st2 = 0.0, x2 = 0.0;
    const bool chained = p < n_pose && chain_mask && chain_mask[p];
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
        if (chained) {
            // coupled elements: the triangular solve runs along the segment (chain_backsub_kernel, vg_priors.cu)
#pragma unroll
            for (int k = 0; k < 6; k++) chain_w[(size_t)p * 6 + k] = wv[k];
        } else {
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

int pose_factor_blocks(int n_pose) { return (n_pose + FACTOR_THREADS - 1) / FACTOR_THREADS; }
int pose_backsub_blocks(int n_pose) { return (n_pose + POSE_THREADS - 1) / POSE_THREADS; }

size_t pose_scratch(int n_pose, int Ks, int n_seg)
{
    const size_t nb_pose = (n_pose + FACTOR_THREADS - 1) / FACTOR_THREADS;
    const size_t npair = (size_t)Ks * (Ks + 1) / 2 + Ks;
    return nb_pose * 4 + nb_pose * npair + 4 * (size_t)n_seg + 16;
}

cudaError_t launch_pose_schur(const DatasetDesc *d_desc, int n_pose, int Ks,
                              const int *pose_start, const int *contrib_ds, const int *contrib_img,
                              double *scale, LmConsts lm, double *ws, double *partial, size_t partial_doubles,
                              double *red, int *fail_flag, int rank, int nranks, SolverLaunch sl,
                              const unsigned char *chain_mask, int n_seg)
{
    // partial: [max |g| of each block (nb_pose) | of each chain segment (n_seg, written by chain_factor) | gram rows]
    const int nb_pose = (n_pose + FACTOR_THREADS - 1) / FACTOR_THREADS;
    if (pose_scratch(n_pose, Ks, n_seg) > partial_doubles) return cudaErrorInvalidValue;
    double *p_gmax = partial;
    double *p_gram = partial + nb_pose + n_seg;
    if (n_pose > 0) {
        pose_factor_kernel<<<nb_pose, FACTOR_THREADS, 0, sl.stream>>>(d_desc, n_pose, Ks, pose_start, contrib_ds,
                                                                     contrib_img, scale, lm, ws, p_gmax, p_gram, fail_flag,
                                                                     chain_mask);
        if (sl.launches) (*sl.launches)++;
    }
    finalize_gram_kernel<<<1, 256, 0, sl.stream>>>(Ks, n_pose > 0 ? nb_pose : 0, p_gram, n_pose > 0 ? nb_pose + n_seg : 0, p_gmax,
                                                   red, fail_flag, rank, nranks);
    if (sl.launches) (*sl.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_pose_backsub(int n_pose, int Ks, const double *delta_a,
                                const double *const *seq_cur, double *const *seq_cand,
                                const int *pose_seq, const int *pose_local,
                                const double *ws, double *partial, size_t partial_doubles, double *red,
                                SolverLaunch sl, const unsigned char *chain_mask, double *chain_w)
{
    const int nb = (n_pose + POSE_THREADS - 1) / POSE_THREADS;
    if ((size_t)nb * 3 > partial_doubles) return cudaErrorInvalidValue;
    if (n_pose > 0) {
        pose_backsub_kernel<<<nb, POSE_THREADS, 0, sl.stream>>>(n_pose, Ks, delta_a, seq_cur, seq_cand, pose_seq,
                                                                pose_local, ws, partial, chain_mask, chain_w);
        if (sl.launches) (*sl.launches)++;
    }
    return cudaGetLastError();
}

// sums the rows the back-substitution kernels left: one per pose_backsub block, then one per chain segment
cudaError_t launch_finalize_backsub(int Ks, int n_rows, const double *partial, double *red, SolverLaunch sl)
{
    finalize_backsub_kernel<<<1, 32, 0, sl.stream>>>(Ks, n_rows, partial, red);
    if (sl.launches) (*sl.launches)++;
    return cudaGetLastError();
}

}  // namespace vg
