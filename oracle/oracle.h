/*
 * oracle.h -- C ABI shared by the two CPU checkers:
 *
 *   oracle/pmg_oracle.c     plain-C restatement of the reference CPU multigrid path
 *                           (symbols  orc_*)
 *   oracle/ref_driver.cpp   thin extern "C" driver over the UNMODIFIED reference headers
 *                           compiled in place from /root/reference  (symbols  ref_*)
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load these
 * libraries, and only as the checker / the CPU baseline.  The product (libpmg.so) never links,
 * loads or calls them and has no CPU fallback.
 *
 * Data layout (reference convention, Smoother.hpp:65, DynamicGridUtils.hpp:65): row-major
 * N x N doubles INCLUDING the Dirichlet boundary ring, idx = y*width + x, N = 2^k+1,
 * h = 1/(N-1).
 *
 * Every function exists twice with identical signatures: PREFIX = orc_ (restatement) and
 * PREFIX = ref_ (real reference).  tests/test_oracle_vs_ref.py pins orc_* against ref_* bit
 * for bit, and both against the golden fixtures in tests/golden/.
 */
#ifndef PMG_ORACLE_H
#define PMG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CYCLE_V = 0, ORC_CYCLE_W = 1, ORC_CYCLE_F = 2 };
enum { ORC_PROLONG_REFERENCE = 0, ORC_PROLONG_FULL = 1 };

#define ORC_DECLARE(P)                                                                            \
    /* num_iter+1 weighted-Jacobi sweeps (Smoother.hpp:59 loop is `<=`); residuals (nullable)  */ \
    /* receives the smoother's own per-sweep ||r||; returns the number of sweeps performed     */ \
    int P##jacobi(double *x, const double *f, int width, int height, double h, double omega,      \
                  int num_iter, double eps, double *residuals);                                   \
    void P##residual(double *r, const double *x, const double *f, int width, int height,          \
                     double h);                                                                   \
    double P##norm(const double *v, long l);                                                      \
    void P##restrict_fw(const double *fine, double *coarse, int nf, int nc);                      \
    /* fine += P*coarse ; returns 0, or -1 if the mode is not available in this library        */ \
    int P##prolong_add(double *fine, const double *coarse, int nf, int nc, int mode);             \
    void P##rhs(double *f, int width, int height, double h);                                      \
    void P##exact(double *u, double h, int width, int height);                                    \
    /* one multigrid cycle in place on phi (F = the runner's wrapper, MultiGridTestRunner.hpp: */ \
    /* 192-205: restrict phi to n_coarse, FMG up with the analytic RHS, copy back)             */ \
    int P##cycle(double *phi, const double *f, int n, double h, int kind, double omega,           \
                 double eps, int alpha, int v1, int v2, int prolong_mode);                        \
    /* phi updated in place; hist[0] = ||f - A phi0||, hist[k] = ||r|| after cycle k; stops     */ \
    /* when hist[k] < rel_tol*hist[0] or k == max_cycles; returns number of cycles done        */ \
    int P##solve(double *phi, const double *f, int n, int kind, double omega, double eps,         \
                 int alpha, int v1, int v2, int prolong_mode, double rel_tol, int max_cycles,     \
                 double *hist);                                                                   \
    /* NOT in the reference ("general-RHS FMG"): one full-multigrid pass for an   */ \
    /* arbitrary f and Dirichlet ring -- the SPECIFICATION libpmg's PMG_CYCLE_FMG is tested      */ \
    /* against; the ref_ library returns -1                                                     */ \
    int P##fmg_general(double *phi, const double *f, int n, double h, double omega, int v1,       \
                       int v2, int prolong_mode);

ORC_DECLARE(orc_)
ORC_DECLARE(ref_)

/* ---- smoothers beyond weighted Jacobi and the Krylov wrapper: pmg_oracle_smoothers.c ------------- */
enum { ORC_SMOOTHER_JACOBI = 0, ORC_SMOOTHER_RBGS = 1, ORC_SMOOTHER_GS_LEX = 2, ORC_SMOOTHER_CHEBYSHEV = 3 };
/* Chebyshev-Jacobi smooths the upper part of the spectrum of D^-1 A (eigenvalues in (0, 2)): [1/2, 2] in 2-D */
#define ORC_CHEB_LO 0.5
#define ORC_CHEB_HI 2.0
#define ORC_DECLARE_SMOOTHERS(P)                                                                                      \
    /* GaussSeidelSmoother::smooth (Smoother.hpp:119-168): num_iter lexicographic sweeps (its loop is `<`), eps    */ \
    /* exit; residuals (nullable) gets the per-sweep ||r||; returns the sweeps done                                */ \
    int P##gs(double *x, const double *f, int width, int height, double h, int num_iter, double eps,                  \
              double *residuals);                                                                                     \
    /* one V- or W-cycle with the smoother injected through the reference's Smoother* slot; TRUE sweep counts;     */ \
    /* -1 where the library has no such smoother                                                                   */ \
    int P##cycle_s(double *phi, const double *f, int n, double h, int kind, int smoother, double omega, int alpha,     \
                   int nu1, int nu2, int coarse_sweeps, int prolong_mode);                                            \
    /* ConjugateGradientSmoother::smooth (Smoother.hpp:170-256): x is zeroed, num_iter CG steps, eps exit;          */ \
    /* residuals gets ||r|| before the first and after every step; returns how many were written                   */ \
    int P##cg(double *x, const double *f, int width, int height, double h, int num_iter, double eps,                  \
              double *residuals);
ORC_DECLARE_SMOOTHERS(orc_)
ORC_DECLARE_SMOOTHERS(ref_)
/* not in the reference (the ref_ library has no counterpart) */
int orc_rbgs(double *x, const double *f, int width, int height, double h, int sweeps);
int orc_jacobi_weights(double *x, const double *f, int width, int height, double h, const double *w, int sweeps);
void orc_chebyshev_weights(double lo, double hi, int n, double *w);
int orc_pcg(double *x, const double *f, int n, double h, int precond, int smoother, double omega, int nu1, int nu2,
            int coarse_sweeps, int prolong_mode, double rel_tol, int max_iter, double *hist);

#ifdef __cplusplus
}
#endif
#endif
