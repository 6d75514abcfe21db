/*
 * pmg_oracle_smoothers.c -- CPU checkers for the smoothers and the Krylov wrapper beyond weighted Jacobi
 * (SURVEY.md 8f-3): plain C99, part of oracle/liboracle.so.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h): never linked into, loaded by, or a fallback for libpmg.so.
 *
 * PARITY STATUS, per function (tests/test_oracle_smoothers.py):
 *   orc_gs               PINNED   = GaussSeidelSmoother::smooth (Smoother.hpp:119-168) bit for bit (ref_gs)
 *   orc_cycle_s, GS_LEX  PINNED   = the reference's MultigridSolver with a GaussSeidelSmoother injected (ref_cycle_s)
 *   orc_jacobi_weights   PINNED   = one reference-pinned weighted-Jacobi sweep per weight (ref_jacobi, num_iter = 0)
 *   orc_cg               PINNED   = ConjugateGradientSmoother::smooth (Smoother.hpp:170-256) bit for bit (ref_cg)
 *   orc_rbgs             unpinned : the reference has no red-black ordering.  Same update expression as orc_gs, other
 *                                   visiting order; checked against a hand-computed 5 x 5 case, against "a red half
 *                                   sweep is a Jacobi sweep restricted to the red points", and on a one-row grid against orc_gs.
 *   orc_pcg              unpinned : preconditioned CG is not in the reference.  With precond = 0 it must agree with the
 *                                   pinned orc_cg to rounding (same Krylov iterates, recurrence instead of recomputed r).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static double *zeros_s(long l)
{
    double *p = (double *)calloc((size_t)l, sizeof(double));
    if (!p) abort();
    return p;
}

/* Smoother.hpp:134-145: x = 0.25 * (W + E + S + N + h*h*f), in place, lexicographic; `num_iter` sweeps (loop `<`);
 * after every sweep the full residual norm, pushed to residuals, break if < eps (:147-159) */
int orc_gs(double *x, const double *f, int width, int height, double h, int num_iter, double eps, double *residuals)
{
    long l = (long)width * height;
    int done = 0;
    for (int iter = 0; iter < num_iter; ++iter) {
        for (int y = 1; y < height - 1; ++y)
            for (int xp = 1; xp < width - 1; ++xp) {
                long i = (long)y * width + xp;
                x[i] = 0.25 * (x[i - 1] + x[i + 1] + x[i - width] + x[i + width] + h * h * f[i]);
            }
        ++done;
        double *r = zeros_s(l);
        orc_residual(r, x, f, width, height, h);
        double rn = orc_norm(r, l);
        free(r);
        if (residuals) residuals[iter] = rn;
        if (rn < eps) break;
    }
    return done;
}

/* the same update in red-black order: all points with (x + y) even ("red"), then all with (x + y) odd ("black") */
int orc_rbgs(double *x, const double *f, int width, int height, double h, int sweeps)
{
    for (int s = 0; s < sweeps; ++s)
        for (int colour = 0; colour < 2; ++colour)
            for (int y = 1; y < height - 1; ++y)
                for (int xp = 1; xp < width - 1; ++xp) {
                    if (((xp + y) & 1) != colour) continue;
                    long i = (long)y * width + xp;
                    x[i] = 0.25 * (x[i - 1] + x[i + 1] + x[i - width] + x[i + width] + h * h * f[i]);
                }
    return sweeps;
}

/* weighted Jacobi with one weight per sweep (Chebyshev-Jacobi); each sweep is orc_jacobi(num_iter = 0) */
int orc_jacobi_weights(double *x, const double *f, int width, int height, double h, const double *w, int sweeps)
{
    for (int s = 0; s < sweeps; ++s) orc_jacobi(x, f, width, height, h, w[s], 0, 0.0, NULL);
    return sweeps;
}

/* Chebyshev weights for the interval [lo, hi] of eigenvalues of D^-1 A: w_k = 1 / (d - c cos(pi (2k+1) / (2 n))) */
void orc_chebyshev_weights(double lo, double hi, int n, double *w)
{
    const double d = 0.5 * (hi + lo), c = 0.5 * (hi - lo);
    for (int k = 0; k < n; ++k) w[k] = 1.0 / (d - c * cos(M_PI * (2 * k + 1) / (2.0 * n)));
}

static void smooth_s(int smoother, double *x, const double *f, int n, double h, double omega, int sweeps)
{
    if (sweeps <= 0) return;
    if (smoother == ORC_SMOOTHER_JACOBI) {
        orc_jacobi(x, f, n, n, h, omega, sweeps - 1, 0.0, NULL);
    } else if (smoother == ORC_SMOOTHER_RBGS) {
        orc_rbgs(x, f, n, n, h, sweeps);
    } else if (smoother == ORC_SMOOTHER_GS_LEX) {
        orc_gs(x, f, n, n, h, sweeps, 0.0, NULL);
    } else {
        double w[64];
        if (sweeps > 64) abort();
        orc_chebyshev_weights(ORC_CHEB_LO, ORC_CHEB_HI, sweeps, w);
        orc_jacobi_weights(x, f, n, n, h, w, sweeps);
    }
}

/* MultiGrid.hpp:57-136 with the smoother injected through the Smoother* slot; sweep counts are TRUE counts */
static void mu_cycle_s(int smoother, double omega, int alpha, int nu1, int nu2, int coarse_sweeps, int prolong_mode,
                       double *phi, const double *f, int n, double h, int w_form)
{
    if (n <= 5) {
        smooth_s(smoother, phi, f, n, h, omega, coarse_sweeps);
        return;
    }
    smooth_s(smoother, phi, f, n, h, omega, nu1);
    long l = (long)n * n;
    double *res = zeros_s(l);
    orc_residual(res, phi, f, n, n, h);
    int nc = (n - 1) / 2 + 1;
    long lc = (long)nc * nc;
    double *res_c = zeros_s(lc);
    orc_restrict_fw(res, res_c, n, nc);
    double *e_c = zeros_s(lc);
    int reps = w_form ? alpha : 1;
    for (int k = 0; k < reps; ++k)
        mu_cycle_s(smoother, omega, alpha, nu1, nu2, coarse_sweeps, prolong_mode, e_c, res_c, nc, 2 * h, w_form);
    orc_prolong_add(phi, e_c, n, nc, prolong_mode);
    smooth_s(smoother, phi, f, n, h, omega, nu2);
    free(res);
    free(res_c);
    free(e_c);
}

int orc_cycle_s(double *phi, const double *f, int n, double h, int kind, int smoother, double omega, int alpha, int nu1,
                int nu2, int coarse_sweeps, int prolong_mode)
{
    if (kind != ORC_CYCLE_V && kind != ORC_CYCLE_W) return -1;
    if (smoother < 0 || smoother > 3) return -1;
    mu_cycle_s(smoother, omega, alpha, nu1, nu2, coarse_sweeps, prolong_mode, phi, f, n, h, kind == ORC_CYCLE_W);
    return 0;
}

/* DynamicGridUtils.hpp:71-82 apply_laplacian: Ap = (4p - W - E - S - N) / h^2 on the interior, ring untouched (zero) */
static void apply_a(const double *p, double *ap, int width, int height, double h)
{
    for (int y = 1; y < height - 1; ++y)
        for (int xp = 1; xp < width - 1; ++xp) {
            long i = (long)y * width + xp;
            ap[i] = (4 * p[i] - p[i - 1] - p[i + 1] - p[i - width] - p[i + width]) / (h * h);
        }
}

static double dot_s(const double *a, const double *b, long l)
{
    double acc = 0.0;
    for (long i = 0; i < l; ++i) acc += a[i] * b[i];
    return acc;
}

/* ConjugateGradientSmoother::smooth (Smoother.hpp:170-256), statement for statement: x is ZEROED first (:186), the
 * residual norm pushed is that of the RECURRENCE residual, after which r is recomputed from x (:216-217) */
int orc_cg(double *x, const double *f, int width, int height, double h, int num_iter, double eps, double *residuals)
{
    long l = (long)width * height;
    double *r = zeros_s(l), *p = zeros_s(l), *ap = zeros_s(l);
    int n_res = 0;
    memset(x, 0, (size_t)l * sizeof(double));
    orc_residual(r, x, f, width, height, h);
    double nr = orc_norm(r, l);
    orc_residual(r, x, f, width, height, h);
    if (residuals) residuals[n_res] = nr;
    ++n_res;
    memcpy(p, r, (size_t)l * sizeof(double));
    for (int iter = 0; iter < num_iter; ++iter) {
        double rtr = dot_s(r, r, l);
        apply_a(p, ap, width, height, h);
        double pap = dot_s(p, ap, l);
        double alpha = rtr / pap;
        for (long i = 0; i < l; ++i) x[i] += alpha * p[i];
        for (long i = 0; i < l; ++i) r[i] -= alpha * ap[i];
        nr = orc_norm(r, l);
        orc_residual(r, x, f, width, height, h);
        if (residuals) residuals[n_res] = nr;
        ++n_res;
        if (nr < eps) break;
        double rtr_new = dot_s(r, r, l);
        double beta = rtr_new / rtr;
        for (long i = 0; i < l; ++i) p[i] = r[i] + beta * p[i];
    }
    free(r);
    free(p);
    free(ap);
    return n_res;
}

/* Preconditioned CG on A x = f from the given x (ring = Dirichlet data), M = one multigrid cycle from a zero start
 * (precond = 1) or the identity (precond = 0).  hist[0] = ||r0||, hist[k] = ||r_k|| (recurrence residual); stops when
 * hist[k] < rel_tol * hist[0].  The order of every operation is what libpmg's pmg_pcg reproduces. */
int orc_pcg(double *x, const double *f, int n, double h, int precond, int smoother, double omega, int nu1, int nu2,
            int coarse_sweeps, int prolong_mode, double rel_tol, int max_iter, double *hist)
{
    long l = (long)n * n;
    double *r = zeros_s(l), *z = zeros_s(l), *p = zeros_s(l), *ap = zeros_s(l);
    orc_residual(r, x, f, n, n, h);
    hist[0] = sqrt(dot_s(r, r, l));
    int k = 0;
    double rz = 0.0;
    while (k < max_iter && !(hist[k] < rel_tol * hist[0]) && hist[0] > 0.0) {
        if (precond) {
            memset(z, 0, (size_t)l * sizeof(double));
            mu_cycle_s(smoother, omega, 1, nu1, nu2, coarse_sweeps, prolong_mode, z, r, n, h, 0);
        } else {
            memcpy(z, r, (size_t)l * sizeof(double));
        }
        double rz_new = dot_s(r, z, l);
        if (k == 0) {
            memcpy(p, z, (size_t)l * sizeof(double));
        } else {
            double beta = rz_new / rz;
            for (long i = 0; i < l; ++i) p[i] = z[i] + beta * p[i];
        }
        rz = rz_new;
        apply_a(p, ap, n, n, h);
        double alpha = rz / dot_s(p, ap, l);
        for (long i = 0; i < l; ++i) x[i] += alpha * p[i];
        for (long i = 0; i < l; ++i) r[i] -= alpha * ap[i];
        ++k;
        hist[k] = sqrt(dot_s(r, r, l));
    }
    free(r);
    free(z);
    free(p);
    free(ap);
    return k;
}
