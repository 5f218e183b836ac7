/*
 * ddw_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Float64 CPU restatement of the convex programs that the reference's
 * NumericalMarkowitz layer hands to cvxpylayers -> diffcp -> SCS
 * (third-party, unpinned in the reference's setup.py:9, absent from
 * /root/reference and not installable here).  The reference states the
 * *problem*; its solution is unique whenever Q is positive definite, so
 * the oracle solves the problem exactly instead of imitating SCS's ADMM
 * (whose own answers carry ~1e-4 noise, docs/source/layers.rst:387-391).
 *
 *   problem     deepdow/layers/allocate.py:217-238
 *                 maximize rets@w - sum_squares(covmat_sqrt@w) - alpha*||w||^2
 *                 s.t. sum(w) == 1, 0 <= w <= max_weight
 *   pre-proc    deepdow/layers/allocate.py:267-273
 *                 gamma_sqrt_ = gamma_sqrt.repeat((1, N*N)).view(B, N, N)   (a TILING:
 *                 gamma_sqrt_[b,j,k] = gamma_sqrt[(b*N*N + j*N + k) % B])
 *                 A = gamma_sqrt_ * covmat_sqrt ; alpha_abs = |alpha|
 *   backward    implicit differentiation of the KKT system on the active
 *               set (what diffcp computes at a non-degenerate solution).
 *
 * Algorithm here (deliberately different from the CUDA kernel's primary
 * path): Mehrotra predictor-corrector interior point in float64 down to a
 * ~1e-13 gap, active-set identification, then an exact reduced-KKT polish
 * with primal-dual corrections.  Every solve returns its own KKT residual,
 * which is a certificate of optimality independent of how w was found.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_INFEASIBLE 1
#define ORC_INEXACT 2

/* ---- small dense helpers (row-major, lower triangle) ------------------- */

static int chol_lower(double *M, int n, int ld)
{
    for (int k = 0; k < n; ++k) {
        double d = M[k * ld + k];
        for (int j = 0; j < k; ++j) d -= M[k * ld + j] * M[k * ld + j];
        if (!(d > 0.0)) return k + 1;
        d = sqrt(d);
        M[k * ld + k] = d;
        for (int i = k + 1; i < n; ++i) {
            double s = M[i * ld + k];
            for (int j = 0; j < k; ++j) s -= M[i * ld + j] * M[k * ld + j];
            M[i * ld + k] = s / d;
        }
    }
    return 0;
}

static void chol_solve(const double *M, int n, int ld, double *x)
{
    for (int i = 0; i < n; ++i) {
        double s = x[i];
        for (int j = 0; j < i; ++j) s -= M[i * ld + j] * x[j];
        x[i] = s / M[i * ld + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int j = i + 1; j < n; ++j) s -= M[j * ld + i] * x[j];
        x[i] = s / M[i * ld + i];
    }
}

static void symv(const double *Q, int n, const double *x, double *y)
{
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += Q[i * n + j] * x[j];
        y[i] = s;
    }
}

/* A = tile(gamma) * covmat_sqrt for problem b  (allocate.py:268-273) */
static void build_A(const double *C, const double *gamma, long B, long N, long b, double *A)
{
    long base = b * N * N;
    for (long e = 0; e < N * N; ++e) A[e] = gamma[(base + e) % B] * C[base + e];
}

/* Q = A^T A + |alpha| I */
static void build_Q(const double *A, long N, double alpha_abs, double *Q)
{
    for (long i = 0; i < N; ++i)
        for (long j = 0; j <= i; ++j) {
            double s = 0.0;
            for (long k = 0; k < N; ++k) s += A[k * N + i] * A[k * N + j];
            if (i == j) s += alpha_abs;
            Q[i * N + j] = s;
            Q[j * N + i] = s;
        }
}

/* ---- reduced KKT solve on a given active set ---------------------------- */
/* st[i]: -1 at 0, +1 at u, 0 free.  Returns 0 ok / 1 singular.  nu_out is
 * the multiplier of sum(w)=1 (undefined when there is no free variable). */
static int reduced_solve(const double *Q, const double *r, double u, int n, const signed char *st,
                         double *w, double *nu_out, int *nfree_out, double *M, double *y1, double *y2,
                         int *idx)
{
    int nf = 0, nu_cnt = 0;
    for (int i = 0; i < n; ++i) {
        if (st[i] == 0) idx[nf++] = i;
        else if (st[i] > 0) nu_cnt++;
        w[i] = st[i] > 0 ? u : 0.0;
    }
    *nfree_out = nf;
    if (nf == 0) return 0;
    double c = 1.0 - u * (double)nu_cnt;
    for (int a = 0; a < nf; ++a) {
        int i = idx[a];
        double s = 0.0;
        for (int j = 0; j < n; ++j)
            if (st[j] > 0) s += Q[i * n + j];
        y1[a] = r[i] - 2.0 * u * s;
        y2[a] = 1.0;
        for (int bq = 0; bq <= a; ++bq) M[a * nf + bq] = Q[i * n + idx[bq]];
    }
    if (chol_lower(M, nf, nf)) return 1;
    chol_solve(M, nf, nf, y1);
    chol_solve(M, nf, nf, y2);
    double s1 = 0.0, s2 = 0.0;
    for (int a = 0; a < nf; ++a) { s1 += y1[a]; s2 += y2[a]; }
    double nu = (s1 - 2.0 * c) / s2;
    for (int a = 0; a < nf; ++a) w[idx[a]] = 0.5 * (y1[a] - nu * y2[a]);
    *nu_out = nu;
    return 0;
}

/* KKT residual of (w, st) in the units of the gradient; also returns nu */
static double kkt_residual(const double *Q, const double *r, double u, int n, const double *w,
                           const signed char *st, double *h, double *nu_io, int have_nu)
{
    symv(Q, n, w, h);
    double sumw = 0.0;
    for (int i = 0; i < n; ++i) { h[i] = 2.0 * h[i] - r[i]; sumw += w[i]; }
    double nu;
    if (have_nu) nu = *nu_io;
    else {
        /* no free variable: any nu in [max_L(-h), min_U(-h)] certifies optimality */
        double lo = -INFINITY, hi = INFINITY;
        for (int i = 0; i < n; ++i) {
            if (st[i] < 0 && -h[i] > lo) lo = -h[i];
            if (st[i] > 0 && -h[i] < hi) hi = -h[i];
        }
        if (lo == -INFINITY && hi == INFINITY) nu = 0.0;
        else if (lo == -INFINITY) nu = hi;
        else if (hi == INFINITY) nu = lo;
        else nu = 0.5 * (lo + hi);
        *nu_io = nu;
    }
    double res = fabs(sumw - 1.0);
    for (int i = 0; i < n; ++i) {
        double g = h[i] + nu, v;
        if (st[i] == 0) v = fabs(g);
        else if (st[i] < 0) v = g < 0 ? -g : 0.0;
        else v = g > 0 ? g : 0.0;
        if (v > res) res = v;
        if (-w[i] > res) res = -w[i];
        if (w[i] - u > res) res = w[i] - u;
    }
    return res;
}

/* exact active-set polish starting from a guess; returns 1 when a KKT point was reached */
static int polish(const double *Q, const double *r, double u, int n, signed char *st, double *w,
                  double *nu_out, double *wk /* n*n + 4n */, int *idx)
{
    double *M = wk, *y1 = wk + n * n, *y2 = y1 + n, *h = y2 + n;
    const double tol = 1e-13;
    for (int it = 0; it < 60; ++it) {
        int nf; double nu = 0.0;
        if (reduced_solve(Q, r, u, n, st, w, &nu, &nf, M, y1, y2, idx)) return 0;
        symv(Q, n, w, h);
        for (int i = 0; i < n; ++i) h[i] = 2.0 * h[i] - r[i];
        int changed = 0;
        if (nf > 0) {
            for (int i = 0; i < n; ++i) {
                double g = h[i] + nu;
                if (st[i] == 0) {
                    if (w[i] < -tol) { st[i] = -1; changed = 1; }
                    else if (w[i] > u + tol) { st[i] = 1; changed = 1; }
                } else if (st[i] < 0) { if (g < -tol) { st[i] = 0; changed = 1; } }
                else { if (g > tol) { st[i] = 0; changed = 1; } }
            }
        } else {
            int nu_cnt = 0, il = -1, iu = -1;
            for (int i = 0; i < n; ++i) {
                if (st[i] > 0) { nu_cnt++; if (iu < 0 || h[i] > h[iu]) iu = i; }
                else if (il < 0 || h[i] < h[il]) il = i;
            }
            double c = 1.0 - u * nu_cnt;
            if (c > 1e-12 && il >= 0) { st[il] = 0; changed = 1; }
            else if (c < -1e-12 && iu >= 0) { st[iu] = 0; changed = 1; }
            else if (il >= 0 && iu >= 0 && -h[il] > -h[iu] + tol) { st[il] = 0; st[iu] = 0; changed = 1; }
        }
        if (!changed) {
            /* tie rule shared with the CUDA kernel: a free weight that lands exactly on a bound is
             * reported as active, so the state is a function of the returned weights */
            for (int i = 0; i < n; ++i) {
                w[i] = w[i] < 0 ? 0.0 : (w[i] > u ? u : w[i]);
                if (st[i] == 0) st[i] = w[i] <= 0.0 ? -1 : (w[i] >= u ? 1 : 0);
            }
            *nu_out = nu;
            return 1 + (nf > 0);
        }
    }
    return 0;
}

/* ---- Mehrotra predictor-corrector IPM ----------------------------------- */
static double step_len(int n, int useU, const double *w, const double *t, const double *lam,
                       const double *mu, const double *dw, const double *dlam, const double *dmu)
{
    double a = 1.0;
    for (int i = 0; i < n; ++i) {
        if (dw[i] < 0 && -w[i] / dw[i] < a) a = -w[i] / dw[i];
        if (dlam[i] < 0 && -lam[i] / dlam[i] < a) a = -lam[i] / dlam[i];
        if (useU) {
            if (dw[i] > 0 && t[i] / dw[i] < a) a = t[i] / dw[i];
            if (dmu[i] < 0 && -mu[i] / dmu[i] < a) a = -mu[i] / dmu[i];
        }
    }
    return a;
}

static void ipm(const double *Q, const double *r, double u, int n, double *w, double *lam, double *mu,
                double *nu_out, double *wk /* n*n + 10n */)
{
    double *H = wk, *h = wk + n * n, *y2 = h + n, *dw = y2 + n, *dlam = dw + n, *dmu = dlam + n,
           *dw2 = dmu + n, *dl2 = dw2 + n, *dm2 = dl2 + n, *t = dm2 + n, *rhs = t + n;
    int useU = u < 1.0;
    int ncomp = n * (useU ? 2 : 1);
    double nu = 0.0;
    for (int i = 0; i < n; ++i) {
        w[i] = 1.0 / n;
        t[i] = u - w[i];
        lam[i] = 1.0 / w[i];
        mu[i] = useU ? 1.0 / t[i] : 0.0;
    }
    for (int it = 0; it < 80; ++it) {
        symv(Q, n, w, h);
        double rp = -1.0, gap = 0.0, rdmax = 0.0;
        for (int i = 0; i < n; ++i) {
            h[i] = 2.0 * h[i] - r[i] + nu;
            rp += w[i];
            gap += lam[i] * w[i] + (useU ? mu[i] * t[i] : 0.0);
            double rd = fabs(h[i] - lam[i] + mu[i]);
            if (rd > rdmax) rdmax = rd;
        }
        gap /= ncomp;
        if (gap < 1e-14 && rdmax < 1e-11 && fabs(rp) < 1e-13) break;
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j <= i; ++j) H[i * n + j] = 2.0 * Q[i * n + j];
            H[i * n + i] += lam[i] / w[i] + (useU ? mu[i] / t[i] : 0.0);
        }
        if (chol_lower(H, n, n)) break;
        double s2 = 0.0;
        for (int i = 0; i < n; ++i) y2[i] = 1.0;
        chol_solve(H, n, n, y2);
        for (int i = 0; i < n; ++i) s2 += y2[i];
        /* predictor */
        double s1 = 0.0;
        for (int i = 0; i < n; ++i) dw[i] = -h[i];
        chol_solve(H, n, n, dw);
        for (int i = 0; i < n; ++i) s1 += dw[i];
        double dnu = (s1 + rp) / s2;
        for (int i = 0; i < n; ++i) {
            dw[i] -= dnu * y2[i];
            dlam[i] = -lam[i] - lam[i] / w[i] * dw[i];
            dmu[i] = useU ? -mu[i] + mu[i] / t[i] * dw[i] : 0.0;
        }
        double a = step_len(n, useU, w, t, lam, mu, dw, dlam, dmu);
        double gap_aff = 0.0;
        for (int i = 0; i < n; ++i)
            gap_aff += (lam[i] + a * dlam[i]) * (w[i] + a * dw[i]) +
                       (useU ? (mu[i] + a * dmu[i]) * (t[i] - a * dw[i]) : 0.0);
        gap_aff /= ncomp;
        double sigma = gap_aff / gap;
        sigma = sigma * sigma * sigma;
        /* corrector */
        for (int i = 0; i < n; ++i) {
            double cl = (sigma * gap - dw[i] * dlam[i]) / w[i];
            double cu = useU ? (sigma * gap + dw[i] * dmu[i]) / t[i] : 0.0;
            rhs[i] = cl - cu;
            dw2[i] = -h[i] + cl - cu;
        }
        chol_solve(H, n, n, dw2);
        s1 = 0.0;
        for (int i = 0; i < n; ++i) s1 += dw2[i];
        double dnu2 = (s1 + rp) / s2;
        for (int i = 0; i < n; ++i) {
            dw2[i] -= dnu2 * y2[i];
            double cl = (sigma * gap - dw[i] * dlam[i]) / w[i];
            double cu = useU ? (sigma * gap + dw[i] * dmu[i]) / t[i] : 0.0;
            dl2[i] = -lam[i] + cl - lam[i] / w[i] * dw2[i];
            dm2[i] = useU ? -mu[i] + cu + mu[i] / t[i] * dw2[i] : 0.0;
        }
        a = 0.995 * step_len(n, useU, w, t, lam, mu, dw2, dl2, dm2);
        if (a > 1.0) a = 1.0;
        for (int i = 0; i < n; ++i) {
            w[i] += a * dw2[i];
            lam[i] += a * dl2[i];
            if (useU) mu[i] += a * dm2[i];
            t[i] = u - w[i];
        }
        nu += a * dnu2;
    }
    *nu_out = nu;
}

/* one problem: Q (n*n), r -> w, st, kkt.  wk: n*n + 16n doubles, idx: n ints */
static int solve_one(const double *Qin, const double *rin, double u, int n, double *w, signed char *st,
                     double *kkt_out, double *wk, int *idx)
{
    double nu_feas = (double)n * u;
    if (nu_feas < 1.0 - 1e-12) {
        for (int i = 0; i < n; ++i) { w[i] = NAN; st[i] = 0; }
        *kkt_out = INFINITY;
        return ORC_INFEASIBLE;
    }
    double *Q = wk, *r = Q + n * n, *lam = r + n, *mu = lam + n, *wi = mu + n, *rest = wi + n;
    /* scale the objective so gradients are O(1); the minimiser does not change */
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        if (fabs(rin[i]) > s) s = fabs(rin[i]);
        if (2.0 * Qin[i * n + i] > s) s = 2.0 * Qin[i * n + i];
    }
    if (!(s > 0.0)) s = 1.0;
    for (int e = 0; e < n * n; ++e) Q[e] = Qin[e] / s;
    for (int i = 0; i < n; ++i) r[i] = rin[i] / s;
    if (nu_feas <= 1.0 + 1e-12) { /* the box leaves a single feasible point */
        for (int i = 0; i < n; ++i) { w[i] = u; st[i] = 1; }
        double nu = 0.0;
        *kkt_out = kkt_residual(Q, r, u, n, w, st, rest, &nu, 0);
        return ORC_OK;
    }
    double nu = 0.0;
    /* rest needs n*n + 11n for ipm: allocate inside wk tail */
    ipm(Q, r, u, n, wi, lam, mu, &nu, rest);
    for (int i = 0; i < n; ++i) {
        st[i] = 0;
        if (wi[i] < lam[i]) st[i] = -1;
        if (u < 1.0 && (u - wi[i]) < mu[i]) st[i] = 1;
    }
    int ok = polish(Q, r, u, n, st, w, &nu, rest, idx);
    if (ok) {
        *kkt_out = kkt_residual(Q, r, u, n, w, st, rest, &nu, ok == 2);
        return ORC_OK;
    }
    /* polish cycled (degenerate): fall back to the interior-point iterate */
    for (int i = 0; i < n; ++i) {
        w[i] = wi[i] < 0 ? 0.0 : (wi[i] > u ? u : wi[i]);
        st[i] = (w[i] < 1e-10) ? -1 : ((u - w[i] < 1e-10) ? 1 : 0);
    }
    *kkt_out = kkt_residual(Q, r, u, n, w, st, rest, &nu, 1);
    return ORC_INEXACT;
}

/* ---- public entry points ------------------------------------------------ */

/* forward of NumericalMarkowitz.forward (allocate.py:240-273), float64.
 * rets (B,N), covmat_sqrt (B,N,N), gamma_sqrt (B,), alpha (B,), all contiguous.
 * Out: w (B,N), state (B,N) in {-1,0,+1}, kkt (B,) scaled KKT residual, status (B,). */
int orc_markowitz_fwd(const double *rets, const double *covmat_sqrt, const double *gamma_sqrt,
                      const double *alpha, double max_weight, long B, long N, double *w,
                      signed char *state, double *kkt, int *status, int nthreads)
{
    int n = (int)N;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    int worst = 0;
#pragma omp parallel
    {
        double *A = (double *)malloc(sizeof(double) * (size_t)n * n);
        double *Q = (double *)malloc(sizeof(double) * (size_t)n * n);
        double *wk = (double *)malloc(sizeof(double) * ((size_t)3 * n * n + 40 * (size_t)n + 64));
        int *idx = (int *)malloc(sizeof(int) * (size_t)n);
#pragma omp for schedule(dynamic, 4)
        for (long b = 0; b < B; ++b) {
            build_A(covmat_sqrt, gamma_sqrt, B, N, b, A);
            build_Q(A, N, fabs(alpha[b]), Q);
            int rc = solve_one(Q, rets + b * N, max_weight, n, w + b * N, state + b * N, kkt + b, wk, idx);
            status[b] = rc;
            if (rc) {
#pragma omp critical
                if (rc > worst) worst = rc;
            }
        }
        free(A); free(Q); free(wk); free(idx);
    }
    return worst;
}

/* backward: implicit differentiation of the KKT system on the active set.
 *   F = {st == 0};  [[2 Q_FF, 1],[1^T, 0]] [d_F; eta] = [g_F; 0],  d = 0 off F
 *   dL/drets = d ; dL/dQ = -2 d w^T ; dL/dA = -2 [(A d) w^T + (A w) d^T]
 *   dL/dcovmat_sqrt = gamma_ * dL/dA ; dL/dgamma_ = covmat_sqrt * dL/dA (scatter through
 *   the tiling) ; dL/dalpha = sign(alpha) * (-2 d.w)
 */
int orc_markowitz_bwd(const double *grad_w, const double *w, const signed char *state,
                      const double *covmat_sqrt, const double *gamma_sqrt, const double *alpha,
                      double max_weight, long B, long N, double *d_rets, double *d_covmat_sqrt,
                      double *d_gamma, double *d_alpha, int nthreads)
{
    int n = (int)N;
    (void)max_weight;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double *vecs = (double *)malloc(sizeof(double) * (size_t)B * 2 * n);
#pragma omp parallel
    {
        double *A = (double *)malloc(sizeof(double) * (size_t)n * n);
        double *M = (double *)malloc(sizeof(double) * (size_t)n * n);
        double *y1 = (double *)malloc(sizeof(double) * 4 * (size_t)n);
        int *idx = (int *)malloc(sizeof(int) * (size_t)n);
#pragma omp for schedule(dynamic, 4)
        for (long b = 0; b < B; ++b) {
            const double *g = grad_w + b * N, *wb = w + b * N;
            const signed char *st = state + b * N;
            double *y2 = y1 + n, *p = vecs + b * 2 * n, *q = p + n, *d = d_rets + b * N;
            build_A(covmat_sqrt, gamma_sqrt, B, N, b, A);
            int nf = 0;
            for (int i = 0; i < n; ++i) { d[i] = 0.0; if (st[i] == 0) idx[nf++] = i; }
            if (nf > 0) {
                double aa = fabs(alpha[b]);
                for (int a = 0; a < nf; ++a) {
                    for (int c = 0; c <= a; ++c) {
                        double s = 0.0;
                        for (int k = 0; k < n; ++k) s += A[k * n + idx[a]] * A[k * n + idx[c]];
                        if (a == c) s += aa;
                        M[a * nf + c] = s;
                    }
                    y1[a] = g[idx[a]];
                    y2[a] = 1.0;
                }
                if (chol_lower(M, nf, nf) == 0) {
                    chol_solve(M, nf, nf, y1);
                    chol_solve(M, nf, nf, y2);
                    double s1 = 0.0, s2 = 0.0;
                    for (int a = 0; a < nf; ++a) { s1 += y1[a]; s2 += y2[a]; }
                    double eta = s1 / s2;
                    for (int a = 0; a < nf; ++a) d[idx[a]] = 0.5 * (y1[a] - eta * y2[a]);
                }
            }
            double dot = 0.0;
            for (int j = 0; j < n; ++j) {
                double sp = 0.0, sq = 0.0;
                for (int k = 0; k < n; ++k) { sp += A[j * n + k] * d[k]; sq += A[j * n + k] * wb[k]; }
                p[j] = sp; q[j] = sq;
                dot += d[j] * wb[j];
            }
            double sg = alpha[b] > 0 ? 1.0 : (alpha[b] < 0 ? -1.0 : 0.0);
            d_alpha[b] = sg * (-2.0 * dot);
            long base = b * N * N;
            for (int j = 0; j < n; ++j)
                for (int k = 0; k < n; ++k) {
                    long e = base + (long)j * n + k;
                    double dA = -2.0 * (p[j] * wb[k] + q[j] * d[k]);
                    d_covmat_sqrt[e] = gamma_sqrt[e % B] * dA;
                }
        }
        free(A); free(M); free(y1); free(idx);
    }
    /* d_gamma[m] = sum over flat == m (mod B) of covmat_sqrt[flat] * dA[flat] */
#pragma omp parallel for schedule(static)
    for (long m = 0; m < B; ++m) {
        double acc = 0.0;
        for (long flat = m; flat < B * N * N; flat += B) {
            long b = flat / (N * N), rem = flat - b * N * N;
            long j = rem / N, k = rem - j * N;
            const double *p = vecs + b * 2 * n, *q = p + n;
            double dA = -2.0 * (p[j] * w[b * N + k] + q[j] * d_rets[b * N + k]);
            acc += covmat_sqrt[flat] * dA;
        }
        d_gamma[m] = acc;
    }
    free(vecs);
    return 0;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
