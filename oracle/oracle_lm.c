/*
 * oracle_lm.c -- CPU ORACLE (test infrastructure, NOT product code).  See oracle_lm.h.
 */
#include "oracle_lm.h"
#include "visgeom_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_K 16
#define MAX_CHAIN 5

typedef struct {
    int model, K, constant, shared_off;
    double params[MAX_K], lo[MAX_K], hi[MAX_K];
} cam_t;

typedef struct {
    int is_global, constant, n;
    int shared_off;   /* global & free */
    int pose_off;     /* sequence & free */
    double *values;
    unsigned char *fixed;   /* per element: SetParameterBlockConstant on one element ("anchor", unified_calibration.cpp:803-806) */
} tr_t;

/* TransformationPrior block on element `index` of a transform (unified_calibration.cpp:808-829) */
typedef struct { int tr, index; vgo_transformation_prior f; } tp_t;
/* OdometryPrior block between elements i and i+1 of a sequence transform (unified_calibration.cpp:793-802) */
typedef struct { int tr, i; vgo_odometry_prior f; } op_t;

typedef struct {
    int cam, P, n_img, L, D, ne;
    int tr[MAX_CHAIN], status[MAX_CHAIN];
    double *board, *obs;
    int *seq_index;
    double *H;        /* n_img x ne */
    double loss_a;    /* SoftLOneLoss(a) on every block; 0: NULL loss */
} ds_t;

struct vgo_problem {
    int n_cam, n_tr, n_ds, n_tp, n_op, n_fixed;
    cam_t *cams;
    tr_t *trs;
    ds_t *dss;
    tp_t *tps;
    op_t *ops;
};

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void vgo_solve_options_default(vgo_solve_options *o)
{
    o->max_num_iterations = 1000;
    o->function_tolerance = 1e-15;
    o->gradient_tolerance = 1e-15;
    o->parameter_tolerance = 1e-15;
    o->initial_radius = 1e4;
    o->max_radius = 1e16;
    o->min_radius = 1e-32;
    o->min_relative_decrease = 1e-3;
    o->min_lm_diagonal = 1e-6;
    o->max_lm_diagonal = 1e32;
    o->jacobi_scaling = 1;
    o->max_consecutive_invalid = 5;
    o->verbose = 0;
    o->threads = 1;
}

vgo_problem *vgo_problem_create(void)
{
    return (vgo_problem *)calloc(1, sizeof(vgo_problem));
}

void vgo_problem_destroy(vgo_problem *p)
{
    if (!p) return;
    for (int i = 0; i < p->n_tr; i++) { free(p->trs[i].values); free(p->trs[i].fixed); }
    free(p->tps); free(p->ops);
    for (int i = 0; i < p->n_ds; i++) {
        free(p->dss[i].board); free(p->dss[i].obs); free(p->dss[i].seq_index); free(p->dss[i].H);
    }
    free(p->cams); free(p->trs); free(p->dss); free(p);
}

int vgo_problem_add_camera(vgo_problem *p, int model, const double *value, int constant)
{
    int K = vgo_num_params(model);
    if (K < 0) return -1;
    p->cams = (cam_t *)realloc(p->cams, sizeof(cam_t) * (size_t)(p->n_cam + 1));
    cam_t *c = &p->cams[p->n_cam];
    memset(c, 0, sizeof *c);
    c->model = model; c->K = K; c->constant = constant; c->shared_off = -1;
    for (int i = 0; i < K; i++) {
        c->params[i] = value[i];
        c->lo[i] = vgo_lower_bound(model, i);   /* unified_calibration.cpp:621-626 */
        c->hi[i] = vgo_upper_bound(model, i);
    }
    return p->n_cam++;
}

int vgo_problem_set_bounds(vgo_problem *p, int cam, int idx, double lo, double hi)
{
    if (cam < 0 || cam >= p->n_cam || idx < 0 || idx >= p->cams[cam].K) return -1;
    p->cams[cam].lo[idx] = lo; p->cams[cam].hi[idx] = hi;
    return 0;
}

int vgo_problem_add_transform(vgo_problem *p, int is_global, int constant, int n, const double *values)
{
    if (n < 1 || (is_global && n != 1)) return -1;
    p->trs = (tr_t *)realloc(p->trs, sizeof(tr_t) * (size_t)(p->n_tr + 1));
    tr_t *t = &p->trs[p->n_tr];
    t->is_global = is_global; t->constant = constant; t->n = n;
    t->shared_off = t->pose_off = -1;
    t->values = (double *)malloc(sizeof(double) * 6 * (size_t)n);
    memcpy(t->values, values, sizeof(double) * 6 * (size_t)n);
    t->fixed = (unsigned char *)calloc((size_t)n, 1);
    return p->n_tr++;
}

int vgo_problem_add_transformation_prior(vgo_problem *p, int tr, int index, const double *stiffness, const double *xi_prior)
{
    if (tr < 0 || tr >= p->n_tr || index < 0 || index >= p->trs[tr].n) return -1;
    p->tps = (tp_t *)realloc(p->tps, sizeof(tp_t) * (size_t)(p->n_tp + 1));
    tp_t *t = &p->tps[p->n_tp];
    t->tr = tr; t->index = index;
    /* the prior value is the transform's value when the block is created (calib_cost_functions.h:85-86) */
    vgo_transformation_prior_init(&t->f, stiffness, xi_prior ? xi_prior : p->trs[tr].values + 6 * (size_t)index);
    return p->n_tp++;
}

int vgo_problem_add_odometry(vgo_problem *p, int tr, double errV, double errW, double lambda, int n, const double *odom)
{
    if (tr < 0 || tr >= p->n_tr || p->trs[tr].is_global || n != p->trs[tr].n) return -1;
    p->ops = (op_t *)realloc(p->ops, sizeof(op_t) * (size_t)(p->n_op + (n > 1 ? n - 1 : 0) + 1));
    for (int i = 0; i + 1 < n; i++) {
        op_t *e = &p->ops[p->n_op++];
        e->tr = tr; e->i = i;
        vgo_odometry_prior_init(&e->f, errV, errW, lambda, odom + 6 * (size_t)i, odom + 6 * (size_t)(i + 1));
    }
    return 0;
}

int vgo_problem_set_loss(vgo_problem *p, int dataset, double a)
{
    if (dataset < 0 || dataset >= p->n_ds || !(a >= 0.0)) return -1;
    p->dss[dataset].loss_a = a;
    return 0;
}

int vgo_problem_set_pose_constant(vgo_problem *p, int tr, int index, int constant)
{
    if (tr < 0 || tr >= p->n_tr || index < 0 || index >= p->trs[tr].n) return -1;
    p->n_fixed += (constant ? 1 : 0) - (p->trs[tr].fixed[index] ? 1 : 0);
    p->trs[tr].fixed[index] = constant ? 1 : 0;
    return 0;
}

int vgo_problem_add_dataset(vgo_problem *p, int cam, int P, const double *board,
                            int n_img, const double *obs, const int *seq_index,
                            int chain_len, const int *transform_ids, const int *status)
{
    if (cam < 0 || cam >= p->n_cam || chain_len < 1 || chain_len > MAX_CHAIN || P < 1 || n_img < 0)
        return -1;
    int nseq = 0, seq_tr = -1;
    for (int e = 0; e < chain_len; e++) {
        if (transform_ids[e] < 0 || transform_ids[e] >= p->n_tr) return -1;
        if (!p->trs[transform_ids[e]].is_global) { nseq++; seq_tr = transform_ids[e]; }
    }
    if (nseq != 1) return -2;          /* unified_calibration.cpp:223-228 */
    for (int i = 0; i < n_img; i++) {
        int s = seq_index ? seq_index[i] : i;
        if (s < 0 || s >= p->trs[seq_tr].n) return -3;
    }
    p->dss = (ds_t *)realloc(p->dss, sizeof(ds_t) * (size_t)(p->n_ds + 1));
    ds_t *d = &p->dss[p->n_ds];
    memset(d, 0, sizeof *d);
    d->cam = cam; d->P = P; d->n_img = n_img; d->L = chain_len;
    d->D = p->cams[cam].K + 6 * chain_len;
    d->ne = (d->D + 1) * (d->D + 2) / 2;
    for (int e = 0; e < chain_len; e++) { d->tr[e] = transform_ids[e]; d->status[e] = status[e]; }
    d->board = (double *)malloc(sizeof(double) * 3 * (size_t)P);
    memcpy(d->board, board, sizeof(double) * 3 * (size_t)P);
    d->obs = (double *)malloc(sizeof(double) * 2 * (size_t)P * (size_t)(n_img ? n_img : 1));
    memcpy(d->obs, obs, sizeof(double) * 2 * (size_t)P * (size_t)n_img);
    d->seq_index = (int *)malloc(sizeof(int) * (size_t)(n_img ? n_img : 1));
    for (int i = 0; i < n_img; i++) d->seq_index[i] = seq_index ? seq_index[i] : i;
    d->H = (double *)malloc(sizeof(double) * (size_t)d->ne * (size_t)(n_img ? n_img : 1));
    return p->n_ds++;
}

int vgo_problem_get_camera(const vgo_problem *p, int cam, double *out)
{
    if (cam < 0 || cam >= p->n_cam) return -1;
    memcpy(out, p->cams[cam].params, sizeof(double) * (size_t)p->cams[cam].K);
    return 0;
}

int vgo_problem_set_camera(vgo_problem *p, int cam, const double *value)
{
    if (cam < 0 || cam >= p->n_cam) return -1;
    memcpy(p->cams[cam].params, value, sizeof(double) * (size_t)p->cams[cam].K);
    return 0;
}

int vgo_problem_get_transform(const vgo_problem *p, int tr, double *out)
{
    if (tr < 0 || tr >= p->n_tr) return -1;
    memcpy(out, p->trs[tr].values, sizeof(double) * 6 * (size_t)p->trs[tr].n);
    return 0;
}

int vgo_problem_set_transform(vgo_problem *p, int tr, const double *values)
{
    if (tr < 0 || tr >= p->n_tr) return -1;
    memcpy(p->trs[tr].values, values, sizeof(double) * 6 * (size_t)p->trs[tr].n);
    return 0;
}

/* --- evaluation of one dataset at the problem's current parameters --- */
static void ds_gather_xi(const vgo_problem *p, const ds_t *d, double **seq_tmp,
                         const double *xi[MAX_CHAIN], int is_global[MAX_CHAIN])
{
    *seq_tmp = NULL;
    for (int e = 0; e < d->L; e++) {
        const tr_t *t = &p->trs[d->tr[e]];
        is_global[e] = t->is_global;
        if (t->is_global) xi[e] = t->values;
        else {
            /* gather the per-image poses through seq_index */
            double *g = (double *)malloc(sizeof(double) * 6 * (size_t)(d->n_img ? d->n_img : 1));
            for (int i = 0; i < d->n_img; i++)
                memcpy(g + 6 * (size_t)i, t->values + 6 * (size_t)d->seq_index[i], 6 * sizeof(double));
            xi[e] = g;
            *seq_tmp = g;
        }
    }
}

/* Ceres' SoftLOneLoss::Evaluate (loss_function.cc: sum = 1 + s c, tmp = sqrt(sum), rho = 2 b (tmp - 1),
 * rho' = 1 / tmp, rho'' = -c rho' / (2 sum) < 0) followed by its Corrector (corrector.cc): with rho'' <= 0 the
 * residuals and Jacobians of the block are scaled by sqrt(rho') and the block's cost is rho / 2.  On the packed
 * [J r]^T [J r] block: every entry times rho', the last one (r^T r) replaced by rho. */
static void apply_loss(ds_t *d)
{
    if (!(d->loss_a > 0.0)) return;
    const double b = d->loss_a * d->loss_a;
    for (int i = 0; i < d->n_img; i++) {
        double *H = d->H + (size_t)i * d->ne;
        const double tmp = sqrt(1.0 + H[d->ne - 1] / b);
        const double rho1 = 1.0 / tmp;
        for (int e = 0; e < d->ne - 1; e++) H[e] *= rho1;
        H[d->ne - 1] = 2.0 * b * (tmp - 1.0);
    }
}

static void ds_eval(const vgo_problem *p, ds_t *d, double *r, int with_H, int threads)
{
    const double *xi[MAX_CHAIN];
    int is_global[MAX_CHAIN];
    double *tmp;
    ds_gather_xi(p, d, &tmp, xi, is_global);
    vgo_evaluate_batch(p->cams[d->cam].model, p->cams[d->cam].params, d->n_img, d->P,
                       d->board, d->obs, d->L, d->status, is_global, xi,
                       r, NULL, NULL, with_H ? d->H : NULL, threads);
    if (with_H) apply_loss(d);
    free(tmp);
}

int vgo_problem_residuals(vgo_problem *p, int dataset, double *r)
{
    if (dataset < 0 || dataset >= p->n_ds) return -1;
    ds_eval(p, &p->dss[dataset], r, 0, 1);
    return 0;
}

/* --- LM working set --- */
typedef struct {
    int Ks, n_pose;
    double *A, *ga;            /* Ks x Ks, Ks */
    double *C, *E, *b;         /* n_pose x 36, n_pose x Ks x 6, n_pose x 6 */
    double *O;                 /* n_op x 36: block (pose of element i+1, pose of element i) of J^T J */
    double cost;
} normal_eq;

static void layout(vgo_problem *p, int *Ks, int *n_pose)
{
    int off = 0, po = 0;
    for (int i = 0; i < p->n_cam; i++) {
        cam_t *c = &p->cams[i];
        if (c->constant) c->shared_off = -1; else { c->shared_off = off; off += c->K; }
    }
    for (int i = 0; i < p->n_tr; i++) {
        tr_t *t = &p->trs[i];
        t->shared_off = t->pose_off = -1;
        if (t->constant) continue;
        if (t->is_global) { t->shared_off = off; off += 6; }
        else { t->pose_off = po; po += t->n; }
    }
    *Ks = off; *n_pose = po;
}

static double evaluate_all(vgo_problem *p, int threads)
{
    double cost = 0;
    for (int k = 0; k < p->n_ds; k++) {
        ds_t *d = &p->dss[k];
        ds_eval(p, d, NULL, 1, threads);
        for (int i = 0; i < d->n_img; i++) cost += 0.5 * d->H[(size_t)i * d->ne + d->ne - 1];
    }
    for (int k = 0; k < p->n_tp; k++) {
        const tp_t *t = &p->tps[k];
        double r[6];
        vgo_transformation_prior_eval(&t->f, p->trs[t->tr].values + 6 * (size_t)t->index, r, NULL);
        for (int i = 0; i < 6; i++) cost += 0.5 * r[i] * r[i];
    }
    for (int k = 0; k < p->n_op; k++) {
        const op_t *e = &p->ops[k];
        const double *x = p->trs[e->tr].values + 6 * (size_t)e->i;
        double r[6];
        vgo_odometry_prior_eval(&e->f, x, x + 6, r, NULL, NULL);
        for (int i = 0; i < 6; i++) cost += 0.5 * r[i] * r[i];
    }
    return cost;
}

int vgo_problem_evaluate(vgo_problem *p, int threads, double *cost)
{
    *cost = evaluate_all(p, threads);
    return 0;
}

static void assemble(const vgo_problem *p, normal_eq *n)
{
    const int Ks = n->Ks;
    memset(n->A, 0, sizeof(double) * (size_t)Ks * Ks);
    memset(n->ga, 0, sizeof(double) * (size_t)Ks);
    memset(n->C, 0, sizeof(double) * 36 * (size_t)n->n_pose);
    memset(n->E, 0, sizeof(double) * 6 * (size_t)Ks * (size_t)n->n_pose);
    memset(n->b, 0, sizeof(double) * 6 * (size_t)n->n_pose);
    n->cost = 0;
    for (int k = 0; k < p->n_ds; k++) {
        const ds_t *d = &p->dss[k];
        const cam_t *c = &p->cams[d->cam];
        const int W = d->D + 1;
        /* kind: 0 const, 1 shared(idx), 2 pose(dim), 3 residual */
        int kind[64], idx[64], pose_base = -1;
        for (int a = 0; a < c->K; a++) {
            kind[a] = c->shared_off >= 0 ? 1 : 0; idx[a] = c->shared_off + a;
        }
        for (int e = 0; e < d->L; e++) {
            const tr_t *t = &p->trs[d->tr[e]];
            for (int q = 0; q < 6; q++) {
                int a = c->K + 6 * e + q;
                if (t->is_global) { kind[a] = t->shared_off >= 0 ? 1 : 0; idx[a] = t->shared_off + q; }
                else { kind[a] = t->pose_off >= 0 ? 2 : 0; idx[a] = q; pose_base = t->pose_off; }
            }
        }
        kind[d->D] = 3; idx[d->D] = 0;
        for (int i = 0; i < d->n_img; i++) {
            const double *H = d->H + (size_t)i * d->ne;
            const int pz = pose_base >= 0 ? pose_base + d->seq_index[i] : -1;
            int e = 0;
            for (int a = 0; a < W; a++) {
                for (int bb = a; bb < W; bb++, e++) {
                    const double h = H[e];
                    const int ka = kind[a], kb = kind[bb];
                    if (ka == 0 || kb == 0) continue;
                    if (ka == 1 && kb == 1) {
                        n->A[idx[a] * Ks + idx[bb]] += h;
                        if (a != bb) n->A[idx[bb] * Ks + idx[a]] += h;
                    } else if (ka == 1 && kb == 2) {
                        n->E[((size_t)pz * Ks + idx[a]) * 6 + idx[bb]] += h;
                    } else if (ka == 2 && kb == 1) {
                        n->E[((size_t)pz * Ks + idx[bb]) * 6 + idx[a]] += h;
                    } else if (ka == 2 && kb == 2) {
                        n->C[(size_t)pz * 36 + idx[a] * 6 + idx[bb]] += h;
                        if (a != bb) n->C[(size_t)pz * 36 + idx[bb] * 6 + idx[a]] += h;
                    } else if (ka == 1 && kb == 3) {
                        n->ga[idx[a]] += h;
                    } else if (ka == 2 && kb == 3) {
                        n->b[(size_t)pz * 6 + idx[a]] += h;
                    } else if (ka == 3 && kb == 3) {
                        n->cost += 0.5 * h;
                    }
                }
            }
        }
    }
    /* the 6-residual blocks: J^T J, J^T r and 1/2 r^T r of every prior (what Ceres forms from the functor's output) */
    for (int k = 0; k < p->n_tp; k++) {
        const tp_t *t = &p->tps[k];
        const tr_t *tr = &p->trs[t->tr];
        double r[6], J[36];
        vgo_transformation_prior_eval(&t->f, tr->values + 6 * (size_t)t->index, r, J);
        for (int i = 0; i < 6; i++) n->cost += 0.5 * r[i] * r[i];
        for (int a = 0; a < 6; a++) {
            double g = 0;
            for (int i = 0; i < 6; i++) g += J[6 * i + a] * r[i];
            for (int b2 = 0; b2 < 6; b2++) {
                double h = 0;
                for (int i = 0; i < 6; i++) h += J[6 * i + a] * J[6 * i + b2];
                if (tr->shared_off >= 0) n->A[(tr->shared_off + a) * Ks + tr->shared_off + b2] += h;
                else if (tr->pose_off >= 0) n->C[(size_t)(tr->pose_off + t->index) * 36 + 6 * a + b2] += h;
            }
            if (tr->shared_off >= 0) n->ga[tr->shared_off + a] += g;
            else if (tr->pose_off >= 0) n->b[(size_t)(tr->pose_off + t->index) * 6 + a] += g;
        }
    }
    for (int k = 0; k < p->n_op; k++) {
        const op_t *e = &p->ops[k];
        const tr_t *tr = &p->trs[e->tr];
        const double *x = tr->values + 6 * (size_t)e->i;
        double r[6], J1[36], J2[36];
        vgo_odometry_prior_eval(&e->f, x, x + 6, r, J1, J2);
        for (int i = 0; i < 6; i++) n->cost += 0.5 * r[i] * r[i];
        memset(n->O + 36 * (size_t)k, 0, 36 * sizeof(double));
        if (tr->pose_off < 0) continue;
        const size_t q1 = (size_t)(tr->pose_off + e->i), q2 = q1 + 1;
        for (int a = 0; a < 6; a++) {
            double g1 = 0, g2 = 0;
            for (int i = 0; i < 6; i++) { g1 += J1[6 * i + a] * r[i]; g2 += J2[6 * i + a] * r[i]; }
            n->b[q1 * 6 + a] += g1; n->b[q2 * 6 + a] += g2;
            for (int b2 = 0; b2 < 6; b2++) {
                double h11 = 0, h22 = 0, h21 = 0;
                for (int i = 0; i < 6; i++) {
                    h11 += J1[6 * i + a] * J1[6 * i + b2];
                    h22 += J2[6 * i + a] * J2[6 * i + b2];
                    h21 += J2[6 * i + a] * J1[6 * i + b2];
                }
                n->C[q1 * 36 + 6 * a + b2] += h11;
                n->C[q2 * 36 + 6 * a + b2] += h22;
                n->O[36 * (size_t)k + 6 * a + b2] = h21;
            }
        }
    }
    /* constant elements of a free sequence: their columns leave the problem */
    for (int i = 0; i < p->n_tr; i++) {
        const tr_t *t = &p->trs[i];
        if (t->pose_off < 0) continue;
        for (int q = 0; q < t->n; q++) {
            if (!t->fixed[q]) continue;
            const size_t z = (size_t)(t->pose_off + q);
            memset(n->C + z * 36, 0, 36 * sizeof(double));
            memset(n->E + z * Ks * 6, 0, sizeof(double) * 6 * (size_t)Ks);
            memset(n->b + z * 6, 0, 6 * sizeof(double));
        }
    }
    for (int k = 0; k < p->n_op; k++) {
        const op_t *e = &p->ops[k];
        const tr_t *tr = &p->trs[e->tr];
        if (tr->fixed[e->i] || tr->fixed[e->i + 1]) memset(n->O + 36 * (size_t)k, 0, 36 * sizeof(double));
    }
}

static int chol(double *M, int n);
static void chol_solve(const double *Lm, int n, double *x);
static double clampd(double v, double lo, double hi);

/* LM step of a problem whose pose blocks are coupled (odometry): the full (Ks + 6 n_pose) system is formed and
 * factorised densely -- the same normal equations, no structure exploited (an independent route to the step the
 * CUDA engine computes with a block-tridiagonal elimination).  Returns 1 and da, dp, model_change on success. */
static int dense_step(const vgo_problem *p, const normal_eq *n, const double *scale_a, const double *scale_p,
                      double radius, const vgo_solve_options *o, double *da, double *dp, double *model_change)
{
    const int Ks = n->Ks, NP = n->n_pose;
    const size_t N = (size_t)Ks + 6 * (size_t)NP;
    double *M = (double *)calloc(N * N + 1, sizeof(double));
    double *Hm = (double *)calloc(N * N + 1, sizeof(double));
    double *g = (double *)calloc(N + 1, sizeof(double));
    double *d = (double *)calloc(N + 1, sizeof(double));
    for (int i = 0; i < Ks; i++) {
        for (int j = 0; j < Ks; j++) Hm[(size_t)i * N + j] = n->A[i * Ks + j];
        g[i] = n->ga[i];
    }
    for (int q = 0; q < NP; q++) {
        const size_t r0 = (size_t)Ks + 6 * (size_t)q;
        for (int a = 0; a < 6; a++) {
            for (int b2 = 0; b2 < 6; b2++) Hm[(r0 + a) * N + r0 + b2] = n->C[(size_t)q * 36 + 6 * a + b2];
            for (int s2 = 0; s2 < Ks; s2++) {
                const double e = n->E[((size_t)q * Ks + s2) * 6 + a];
                Hm[(r0 + a) * N + s2] = e; Hm[(size_t)s2 * N + r0 + a] = e;
            }
            g[r0 + a] = n->b[(size_t)q * 6 + a];
        }
    }
    for (int k = 0; k < p->n_op; k++) {
        const op_t *e = &p->ops[k];
        const tr_t *tr = &p->trs[e->tr];
        if (tr->pose_off < 0) continue;
        const size_t r1 = (size_t)Ks + 6 * (size_t)(tr->pose_off + e->i), r2 = r1 + 6;
        for (int a = 0; a < 6; a++)
            for (int b2 = 0; b2 < 6; b2++) {
                const double h = n->O[36 * (size_t)k + 6 * a + b2];
                Hm[(r2 + a) * N + r1 + b2] = h; Hm[(r1 + b2) * N + r2 + a] = h;
            }
    }
    memcpy(M, Hm, sizeof(double) * N * N);
    for (size_t j = 0; j < N; j++) {
        const double sc = j < (size_t)Ks ? scale_a[j] : scale_p[j - (size_t)Ks];
        const double s2 = sc * sc, hjj = Hm[j * N + j];
        if (j >= (size_t)Ks) {
            /* a pose nothing observes / a constant element: identity row, zero step */
            const size_t q = (j - (size_t)Ks) / 6;
            int empty = 1;
            for (int k = 0; k < 6; k++) if (n->C[q * 36 + 7 * k] != 0.0) empty = 0;
            if (empty) { M[j * N + j] = 1.0; continue; }
        }
        M[j * N + j] += clampd(s2 * hjj, o->min_lm_diagonal, o->max_lm_diagonal) / (radius * s2);
    }
    int ok = chol(M, (int)N) == 0;
    if (ok) {
        for (size_t j = 0; j < N; j++) d[j] = -g[j];
        chol_solve(M, (int)N, d);
        double gd = 0, dHd = 0;
        for (size_t i = 0; i < N; i++) {
            gd += g[i] * d[i];
            double t = 0;
            for (size_t j = 0; j < N; j++) t += Hm[i * N + j] * d[j];
            dHd += d[i] * t;
        }
        *model_change = -gd - 0.5 * dHd;
        memcpy(da, d, sizeof(double) * (size_t)Ks);
        memcpy(dp, d + Ks, sizeof(double) * 6 * (size_t)NP);
    }
    free(M); free(Hm); free(g); free(d);
    return ok;
}

/* in-place Cholesky of an n x n SPD matrix (row-major, lower); returns 0 on success */
static int chol(double *M, int n)
{
    for (int j = 0; j < n; j++) {
        double s = M[j * n + j];
        for (int k = 0; k < j; k++) s -= M[j * n + k] * M[j * n + k];
        if (!(s > 0.0)) return -1;
        s = sqrt(s);
        M[j * n + j] = s;
        for (int i = j + 1; i < n; i++) {
            double t = M[i * n + j];
            for (int k = 0; k < j; k++) t -= M[i * n + k] * M[j * n + k];
            M[i * n + j] = t / s;
        }
    }
    return 0;
}

static void chol_solve(const double *Lm, int n, double *x)
{
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= Lm[i * n + k] * x[k];
        x[i] = s / Lm[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= Lm[k * n + i] * x[k];
        x[i] = s / Lm[i * n + i];
    }
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct { double *cam; double **tr; } snapshot;

static void save_params(const vgo_problem *p, snapshot *s)
{
    s->cam = (double *)malloc(sizeof(double) * MAX_K * (size_t)(p->n_cam ? p->n_cam : 1));
    s->tr = (double **)malloc(sizeof(double *) * (size_t)(p->n_tr ? p->n_tr : 1));
    for (int i = 0; i < p->n_cam; i++) memcpy(s->cam + MAX_K * i, p->cams[i].params, sizeof(double) * MAX_K);
    for (int i = 0; i < p->n_tr; i++) {
        s->tr[i] = (double *)malloc(sizeof(double) * 6 * (size_t)p->trs[i].n);
        memcpy(s->tr[i], p->trs[i].values, sizeof(double) * 6 * (size_t)p->trs[i].n);
    }
}

static void restore_params(vgo_problem *p, const snapshot *s)
{
    for (int i = 0; i < p->n_cam; i++) memcpy(p->cams[i].params, s->cam + MAX_K * i, sizeof(double) * MAX_K);
    for (int i = 0; i < p->n_tr; i++)
        memcpy(p->trs[i].values, s->tr[i], sizeof(double) * 6 * (size_t)p->trs[i].n);
}

static void free_snapshot(vgo_problem *p, snapshot *s)
{
    for (int i = 0; i < p->n_tr; i++) free(s->tr[i]);
    free(s->tr); free(s->cam);
}

int vgo_problem_solve(vgo_problem *p, const vgo_solve_options *o, vgo_solve_summary *sum)
{
    const double t_start = now_s();
    double t_eval = 0;
    int n_eval = 0;
    normal_eq n;
    layout(p, &n.Ks, &n.n_pose);
    const int Ks = n.Ks, NP = n.n_pose;
    n.A = (double *)calloc((size_t)(Ks * Ks + 1), sizeof(double));
    n.ga = (double *)calloc((size_t)(Ks + 1), sizeof(double));
    n.C = (double *)calloc(36 * (size_t)(NP + 1), sizeof(double));
    n.E = (double *)calloc(6 * (size_t)(Ks + 1) * (size_t)(NP + 1), sizeof(double));
    n.b = (double *)calloc(6 * (size_t)(NP + 1), sizeof(double));
    n.O = (double *)calloc(36 * (size_t)(p->n_op + 1), sizeof(double));
    const int coupled = p->n_op > 0 || p->n_fixed > 0;     /* pose blocks no longer independent / all free */
    double *scale_a = (double *)malloc(sizeof(double) * (size_t)(Ks + 1));
    double *scale_p = (double *)malloc(sizeof(double) * 6 * (size_t)(NP + 1));
    double *Lp = (double *)malloc(sizeof(double) * 36 * (size_t)(NP + 1));   /* chol factors */
    double *S = (double *)malloc(sizeof(double) * (size_t)(Ks * Ks + 1));
    double *rhs = (double *)malloc(sizeof(double) * (size_t)(Ks + 1));
    double *da = (double *)malloc(sizeof(double) * (size_t)(Ks + 1));
    double *dp = (double *)malloc(sizeof(double) * 6 * (size_t)(NP + 1));
    double *Y = (double *)malloc(sizeof(double) * 6 * (size_t)(Ks + 1));
    memset(sum, 0, sizeof *sum);

    double t0 = now_s();
    double cost = evaluate_all(p, o->threads);
    t_eval += now_s() - t0; n_eval++;
    assemble(p, &n);
    cost = n.cost;
    sum->initial_cost = cost;

    /* Jacobi scaling, fixed at iteration 0: s_j = 1/(1+sqrt(sum J_ij^2)) */
    for (int j = 0; j < Ks; j++)
        scale_a[j] = o->jacobi_scaling ? 1.0 / (1.0 + sqrt(n.A[j * Ks + j])) : 1.0;
    for (int q = 0; q < NP; q++)
        for (int k = 0; k < 6; k++)
            scale_p[6 * q + k] = o->jacobi_scaling ? 1.0 / (1.0 + sqrt(n.C[(size_t)q * 36 + 7 * k])) : 1.0;

    double radius = o->initial_radius, decrease_factor = 2.0;
    int invalid_run = 0;
    sum->termination = 3;
    int iter = 0;

    /* gradient check at the start (Ceres IterationZero) */
    for (;;) {
        /* max-norm of the projected gradient */
        double gmax = 0;
        for (int i = 0; i < p->n_cam; i++) {
            const cam_t *c = &p->cams[i];
            if (c->shared_off < 0) continue;
            for (int k = 0; k < c->K; k++) {
                double x = c->params[k], g = n.ga[c->shared_off + k];
                double pg = x - clampd(x - g, c->lo[k], c->hi[k]);
                if (fabs(pg) > gmax) gmax = fabs(pg);
            }
        }
        for (int i = 0; i < p->n_tr; i++) {
            const tr_t *t = &p->trs[i];
            if (t->shared_off >= 0)
                for (int k = 0; k < 6; k++) if (fabs(n.ga[t->shared_off + k]) > gmax) gmax = fabs(n.ga[t->shared_off + k]);
        }
        for (size_t q = 0; q < (size_t)NP * 6; q++) if (fabs(n.b[q]) > gmax) gmax = fabs(n.b[q]);
        if (gmax <= o->gradient_tolerance) { sum->termination = 1; break; }
        if (iter >= o->max_num_iterations) { sum->termination = 3; break; }
        if (radius < o->min_radius) { sum->termination = 4; break; }
        iter++;

        /* ---- LM step at the current radius ---- */
        int ok = 1;
        double model_change = 0, step2 = 0, x2 = 0;
        if (coupled) {
            ok = dense_step(p, &n, scale_a, scale_p, radius, o, da, dp, &model_change);
            if (ok && !(model_change > 0.0)) ok = 0;
        } else {
        memcpy(S, n.A, sizeof(double) * (size_t)Ks * Ks);
        for (int j = 0; j < Ks; j++) {
            double s2 = scale_a[j] * scale_a[j];
            S[j * Ks + j] += clampd(s2 * n.A[j * Ks + j], o->min_lm_diagonal, o->max_lm_diagonal) / (radius * s2);
            rhs[j] = -n.ga[j];
        }
        for (int q = 0; q < NP && ok; q++) {
            double *Lq = Lp + (size_t)q * 36;
            const double *Cq = n.C + (size_t)q * 36;
            const double *Eq = n.E + (size_t)q * Ks * 6;
            const double *bq = n.b + (size_t)q * 6;
            int empty = 1;
            for (int k = 0; k < 6; k++) if (Cq[7 * k] != 0.0) empty = 0;
            if (empty) { memset(Lq, 0, 36 * sizeof(double)); continue; }
            memcpy(Lq, Cq, 36 * sizeof(double));
            for (int k = 0; k < 6; k++) {
                double s2 = scale_p[6 * q + k] * scale_p[6 * q + k];
                Lq[7 * k] += clampd(s2 * Cq[7 * k], o->min_lm_diagonal, o->max_lm_diagonal) / (radius * s2);
            }
            if (chol(Lq, 6)) { ok = 0; break; }
            /* Y = E C~^-1 (row s of E solved against C~), S -= Y E^T, rhs += Y b */
            for (int s = 0; s < Ks; s++) {
                double *y = Y + 6 * s;
                memcpy(y, Eq + 6 * s, 6 * sizeof(double));
                chol_solve(Lq, 6, y);
            }
            for (int s = 0; s < Ks; s++) {
                const double *y = Y + 6 * s;
                for (int s2 = 0; s2 < Ks; s2++) {
                    const double *e2 = Eq + 6 * s2;
                    S[s * Ks + s2] -= y[0] * e2[0] + y[1] * e2[1] + y[2] * e2[2] + y[3] * e2[3] + y[4] * e2[4] + y[5] * e2[5];
                }
                rhs[s] += y[0] * bq[0] + y[1] * bq[1] + y[2] * bq[2] + y[3] * bq[3] + y[4] * bq[4] + y[5] * bq[5];
            }
        }
        if (ok && Ks > 0) {
            if (chol(S, Ks)) ok = 0;
            else { memcpy(da, rhs, sizeof(double) * (size_t)Ks); chol_solve(S, Ks, da); }
        }
        if (ok) {
            /* back substitution + model cost change = -g^T d - 1/2 d^T H d */
            double gd = 0, dHd = 0;
            for (int s = 0; s < Ks; s++) {
                gd += n.ga[s] * da[s];
                double t = 0;
                for (int s2 = 0; s2 < Ks; s2++) t += n.A[s * Ks + s2] * da[s2];
                dHd += da[s] * t;
            }
            for (int q = 0; q < NP; q++) {
                const double *Lq = Lp + (size_t)q * 36;
                const double *Cq = n.C + (size_t)q * 36;
                const double *Eq = n.E + (size_t)q * Ks * 6;
                const double *bq = n.b + (size_t)q * 6;
                double *d = dp + 6 * (size_t)q;
                if (Lq[0] == 0.0) { memset(d, 0, 6 * sizeof(double)); continue; }
                double Etd[6];
                for (int k = 0; k < 6; k++) {
                    double t = 0;
                    for (int s = 0; s < Ks; s++) t += Eq[6 * s + k] * da[s];
                    Etd[k] = t;
                    d[k] = -(bq[k] + t);
                }
                chol_solve(Lq, 6, d);
                for (int k = 0; k < 6; k++) {
                    gd += bq[k] * d[k];
                    double t = 0;
                    for (int k2 = 0; k2 < 6; k2++) t += Cq[6 * k + k2] * d[k2];
                    dHd += d[k] * t + 2.0 * d[k] * Etd[k];
                }
            }
            model_change = -gd - 0.5 * dHd;
            if (!(model_change > 0.0)) ok = 0;
        }
        }
        if (!ok) {
            /* invalid step (Ceres: StepIsInvalid) */
            invalid_run++;
            sum->num_unsuccessful++;
            if (invalid_run >= o->max_consecutive_invalid) { sum->termination = 5; break; }
            radius /= decrease_factor; decrease_factor *= 2.0;
            continue;
        }
        invalid_run = 0;

        /* candidate = Pi(x + delta) */
        snapshot snap;
        save_params(p, &snap);
        for (int i = 0; i < p->n_cam; i++) {
            cam_t *c = &p->cams[i];
            if (c->shared_off < 0) continue;
            for (int k = 0; k < c->K; k++) {
                double x = c->params[k];
                double xn = clampd(x + da[c->shared_off + k], c->lo[k], c->hi[k]);
                x2 += x * x; step2 += (xn - x) * (xn - x);
                c->params[k] = xn;
            }
        }
        for (int i = 0; i < p->n_tr; i++) {
            tr_t *t = &p->trs[i];
            if (t->shared_off >= 0) {
                for (int k = 0; k < 6; k++) {
                    double x = t->values[k], dd = da[t->shared_off + k];
                    x2 += x * x; step2 += dd * dd;
                    t->values[k] = x + dd;
                }
            } else if (t->pose_off >= 0) {
                for (int q = 0; q < t->n; q++)
                    for (int k = 0; k < 6 && !t->fixed[q]; k++) {
                        double x = t->values[6 * q + k], dd = dp[6 * (size_t)(t->pose_off + q) + k];
                        x2 += x * x; step2 += dd * dd;
                        t->values[6 * q + k] = x + dd;
                    }
            }
        }
        if (sqrt(step2) <= o->parameter_tolerance * (sqrt(x2) + o->parameter_tolerance)) {
            restore_params(p, &snap);
            free_snapshot(p, &snap);
            sum->termination = 2;
            break;
        }
        t0 = now_s();
        double new_cost = evaluate_all(p, o->threads);
        t_eval += now_s() - t0; n_eval++;
        const double rho = (cost - new_cost) / model_change;
        if (o->verbose)
            printf("%4d  cost %.12e  new %.12e  rho %.3e  radius %.3e  |step| %.3e\n",
                   iter, cost, new_cost, rho, radius, sqrt(step2));
        if (rho > o->min_relative_decrease) {
            const double cost_change = cost - new_cost;
            const double old_cost = cost;
            assemble(p, &n);
            cost = n.cost;
            sum->num_successful++;
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3.0));
            if (radius > o->max_radius) radius = o->max_radius;
            decrease_factor = 2.0;
            free_snapshot(p, &snap);
            if (fabs(cost_change) <= o->function_tolerance * old_cost) { sum->termination = 0; break; }
        } else {
            restore_params(p, &snap);
            free_snapshot(p, &snap);
            sum->num_unsuccessful++;
            radius /= decrease_factor; decrease_factor *= 2.0;
        }
    }
    sum->iterations = iter;
    sum->final_cost = cost;
    sum->seconds_total = now_s() - t_start;
    sum->seconds_evaluate = t_eval;
    sum->num_evaluations = n_eval;
    free(n.A); free(n.ga); free(n.C); free(n.E); free(n.b); free(n.O);
    free(scale_a); free(scale_p); free(Lp); free(S); free(rhs); free(da); free(dp); free(Y);
    return 0;
}
