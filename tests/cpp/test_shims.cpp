// test_shims.cpp -- a reference-style driver written against include/pmg.hpp only.
// It follows MultigridTestRunner::run_cycle (2_part_MG/MultiGridTestRunner.hpp:127-256): phi = 0, manufactured
// RHS, JacobiSmoother(eps) injected into MultigridSolver(&smoother, alpha, N), a fixed number of cycles through a
// member-function pointer, then "Final Relative L2 Error"; and ParallelTestRunner's operator protocol
// (3_part_parallel/ParallelTestRunner.cu:231-468) through class Parallel on device memory.
// Output: one JSON object per line, checked by tests/test_cpp_shims.py against the goldens.
#include <cmath>
#include <cstdio>
#include <vector>

#include "pmg.hpp"

using namespace pmg::compat;

static void rhs(std::vector<double> &f, std::vector<double> &u, int n)
{
    double h = 1.0 / (n - 1);
    double factor = (M_PI * M_PI) * 2.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double x = i * h, y = j * h;
            f[(size_t)j * n + i] = factor * std::sin(M_PI * x) * std::sin(M_PI * y);
            u[(size_t)j * n + i] = std::sin(M_PI * x) * std::sin(M_PI * y);
        }
}

static double norm(const std::vector<double> &v)
{
    double s = 0;
    for (double x : v) s += x * x;
    return std::sqrt(s);
}

using CycleFunction = void (MultigridSolver::*)(double *, const double *, int, double);  // MultiGridTestRunner.hpp:125

static void run_cycle(const char *name, int n, CycleFunction fn, bool fcycle)
{
    std::vector<double> phi((size_t)n * n, 0.0), f(phi.size()), u(phi.size()), err(phi.size());
    rhs(f, u, n);
    JacobiSmoother smoother(1e-7);           // 2_part_MG/main.cpp:12
    MultigridSolver mg(&smoother, 3, n);     // alpha = 3 (:15)
    double h = 1.0 / (n - 1);
    if (fcycle)
        mg.f_cycle_from_fine(phi.data(), f.data(), n);
    else
        (mg.*fn)(phi.data(), f.data(), n, h);
    for (size_t i = 0; i < phi.size(); ++i) err[i] = phi[i] - u[i];
    std::printf("{\"test\": \"mg_cpu_exec\", \"cycle\": \"%s\", \"n\": %d, \"rel_l2_error\": %.17g}\n", name, n,
                norm(err) / norm(u));
}

// The F-cycle block of MultigridTestRunner::run_cycle, statement for statement (MultiGridTestRunner.hpp:138-146,
// 192-205): coarse right-hand side and zero iterate on the N_coarse grid, compute_coarsest_grid, f_cycle through the
// member-function pointer, final_solution copied back.
static void run_f_cycle_like_the_runner(int n)
{
    std::vector<double> phi((size_t)n * n, 0.0), phi_tmp(phi.size()), f(phi.size()), u(phi.size()), err(phi.size());
    rhs(f, u, n);
    JacobiSmoother smoother(1e-7);
    MultigridSolver mg(&smoother, 3, n);
    int n_coarse = mg.N_coarse;
    int l_coarse = n_coarse * n_coarse;
    double h_coarse = 1.0 / (n_coarse - 1);
    std::vector<double> f_coarse(l_coarse), u_coarse(l_coarse);
    rhs(f_coarse, u_coarse, n_coarse);
    CycleFunction cycle_func = &MultigridSolver::f_cycle;
    {
        double *phi_coarse = new double[l_coarse];
        std::fill(phi_coarse, phi_coarse + l_coarse, 0.0);
        std::copy(phi.begin(), phi.end(), phi_tmp.begin());
        double *before = phi_coarse;
        mg.compute_coarsest_grid(phi_tmp.data(), phi_coarse, n, mg.N_coarse);
        (mg.*cycle_func)(phi_coarse, f_coarse.data(), n_coarse, h_coarse);
        std::copy(mg.final_solution, mg.final_solution + phi.size(), phi.begin());
        delete[] phi_coarse;
        delete[] before;  // (the reference leaks this one)
    }
    for (size_t i = 0; i < phi.size(); ++i) err[i] = phi[i] - u[i];
    std::printf("{\"test\": \"mg_cpu_exec\", \"cycle\": \"F_runner_shape\", \"n\": %d, \"rel_l2_error\": %.17g}\n", n,
                norm(err) / norm(u));
    // per-call knobs: a changed public field must take effect on the next call (not a stale cached hierarchy)
    std::vector<double> a((size_t)n * n, 0.0), b(a.size(), 0.0);
    mg.v_cycle(a.data(), f.data(), n, 1.0 / (n - 1));
    mg.prolong_mode = PMG_PROLONG_FULL;
    mg.v_cycle(b.data(), f.data(), n, 1.0 / (n - 1));
    bool differ = false;
    for (size_t i = 0; i < a.size() && !differ; ++i) differ = a[i] != b[i];
    std::printf("{\"test\": \"knobs\", \"n\": %d, \"prolong_mode_change_seen\": %s}\n", n, differ ? "true" : "false");
}

static void history(int n)
{
    std::vector<double> phi((size_t)n * n, 0.0), f(phi.size()), u(phi.size());
    rhs(f, u, n);
    pmg_config c;
    pmg_config_default(&c, n);
    c.omega = 2.0 / 3.0;
    pmg::Solver s(c);
    s.set_rhs(f.data());
    s.set_guess(phi.data());
    std::vector<double> hist = s.solve(PMG_CYCLE_V, 1e-8, 100);
    std::printf("{\"test\": \"history\", \"n\": %d, \"cycles\": %d, \"hist\": [", n, (int)hist.size() - 1);
    for (size_t i = 0; i < hist.size(); ++i) std::printf("%s%.17g", i ? ", " : "", hist[i]);
    std::printf("]}\n");
}

static void parallel_ops(int n)
{
    size_t l = (size_t)n * n, bytes = l * sizeof(double);
    std::vector<double> x(l, 0.0), f(l), u(l), r(l, 0.0);
    rhs(f, u, n);
    void *dx, *df, *dr, *dc, *dp;
    int nc = (n - 1) / 2 + 1;
    pmg::check(pmg_device_alloc(&dx, bytes));
    pmg::check(pmg_device_alloc(&df, bytes));
    pmg::check(pmg_device_alloc(&dr, bytes));
    pmg::check(pmg_device_alloc(&dc, (size_t)nc * nc * sizeof(double)));
    pmg::check(pmg_device_alloc(&dp, bytes));
    pmg::check(pmg_memcpy(dx, x.data(), bytes, 1, 0));
    pmg::check(pmg_memcpy(df, f.data(), bytes, 1, 0));
    pmg::check(pmg_memcpy(dr, r.data(), bytes, 1, 0));
    pmg::check(pmg_memcpy(dp, r.data(), bytes, 1, 0));
    std::vector<double> zc((size_t)nc * nc, 0.0);
    pmg::check(pmg_memcpy(dc, zc.data(), zc.size() * sizeof(double), 1, 0));
    double h = 1.0 / (n - 1);
    Parallel::ComputeJacobi((double *)dx, (double *)df, n, n, h, 1);  // v = 1 -> 2 sweeps
    Parallel::ComputeResidual((double *)dr, (double *)dx, (double *)df, n, n, h);
    Parallel::ComputeRestriction((double *)dr, (double *)dc, n, nc);
    Parallel::ComputeProlungator((double *)dc, (double *)dp, nc, n);
    std::vector<double> p(l), rc((size_t)nc * nc);
    pmg::check(pmg_memcpy(x.data(), dx, bytes, 0, 1));
    pmg::check(pmg_memcpy(r.data(), dr, bytes, 0, 1));
    pmg::check(pmg_memcpy(rc.data(), dc, rc.size() * sizeof(double), 0, 1));
    pmg::check(pmg_memcpy(p.data(), dp, bytes, 0, 1));
    int m = n / 2, mc = nc / 2;
    std::printf("{\"test\": \"parallel_ops\", \"n\": %d, \"x_mid\": %.17g, \"x_11\": %.17g, \"r_mid\": %.17g, \"r_11\": %.17g, "
                "\"rc_mid\": %.17g, \"rc_11\": %.17g, \"p_11\": %.17g, \"p_22\": %.17g, \"p_23\": %.17g, \"p_33\": %.17g}\n",
                n, x[(size_t)m * n + m], x[(size_t)n + 1], r[(size_t)m * n + m], r[(size_t)n + 1],
                rc[(size_t)mc * nc + mc], rc[(size_t)nc + 1], p[(size_t)n + 1], p[(size_t)2 * n + 2],
                p[(size_t)2 * n + 3], p[(size_t)3 * n + 3]);
    // ParallelMultiGridSolver::v_cycle on device memory (ParallelTestRunner.cu:152-186 protocol, 3 cycles)
    std::vector<double> z(l, 0.0);
    pmg::check(pmg_memcpy(dx, z.data(), bytes, 1, 0));
    ParallelMultiGridSolver pm(3);
    for (int it = 0; it < 3; ++it) pm.v_cycle((double *)dx, (double *)df, n, h);
    pmg::check(pmg_memcpy(x.data(), dx, bytes, 0, 1));
    std::vector<double> err(l);
    for (size_t i = 0; i < l; ++i) err[i] = x[i] - u[i];
    std::printf("{\"test\": \"gpu_exec_v3\", \"n\": %d, \"rel_l2_error\": %.17g}\n", n, norm(err) / norm(u));
    pmg_device_free(dx); pmg_device_free(df); pmg_device_free(dr); pmg_device_free(dc); pmg_device_free(dp);
}

int main()
{
    try {
        for (int n : {129, 257}) {
            run_cycle("V", n, &MultigridSolver::v_cycle, false);
            run_cycle("W", n, &MultigridSolver::w_cycle, false);
            run_cycle("F", n, nullptr, true);
            run_f_cycle_like_the_runner(n);
        }
        history(257);
        for (int n : {9, 33, 257}) parallel_ops(n);
        // Smoother interface: residuals vector, num_iter + 1 sweeps
        {
            int n = 33;
            std::vector<double> x((size_t)n * n, 0.0), f(x.size()), u(x.size());
            rhs(f, u, n);
            JacobiSmoother sm(0.0);
            std::vector<double> res;
            sm.smooth(x.data(), f.data(), n, n, 1.0 / (n - 1), 1, nullptr, &res);
            std::printf("{\"test\": \"smoother\", \"n\": %d, \"sweeps\": %d, \"res0\": %.17g, \"res1\": %.17g, \"x_mid\": %.17g}\n",
                        n, (int)res.size(), res[0], res[1], x[(size_t)(n / 2) * n + n / 2]);
        }
        // the other smoothers of Smoother.hpp behind the same interface: injected into MultigridSolver and called directly
        {
            int n = 65;
            std::vector<double> f((size_t)n * n), u(f.size());
            rhs(f, u, n);
            GaussSeidelSmoother gs(0.0);
            RedBlackGaussSeidelSmoother rb;
            ChebyshevJacobiSmoother ch;
            Smoother *all[3] = {&gs, &rb, &ch};
            const char *names[3] = {"gs_lex", "rbgs", "chebyshev"};
            for (int i = 0; i < 3; ++i) {
                std::vector<double> phi((size_t)n * n, 0.0), err(phi.size());
                MultigridSolver mg(all[i], 2, n);
                mg.v_cycle(phi.data(), f.data(), n, 1.0 / (n - 1));
                for (size_t q = 0; q < phi.size(); ++q) err[q] = phi[q] - u[q];
                std::printf("{\"test\": \"injected\", \"smoother\": \"%s\", \"n\": %d, \"rel_l2_error\": %.17g, \"phi_mid\": %.17g}\n",
                            names[i], n, norm(err) / norm(u), phi[(size_t)(n / 2) * n + n / 2]);
            }
            std::vector<double> x((size_t)n * n, 0.0), res;
            gs.smooth(x.data(), f.data(), n, n, 1.0 / (n - 1), 3, nullptr, &res);
            std::printf("{\"test\": \"gs_smooth\", \"n\": %d, \"sweeps\": %d, \"res_last\": %.17g, \"x_mid\": %.17g}\n", n,
                        (int)res.size(), res.back(), x[(size_t)(n / 2) * n + n / 2]);
            ConjugateGradientSmoother cg(0.0);
            std::vector<double> y((size_t)n * n, 7.0), cres;  // the reference zeroes x first
            cg.smooth(y.data(), f.data(), n, n, 1.0 / (n - 1), 10, nullptr, &cres);
            std::printf("{\"test\": \"cg_smooth\", \"n\": %d, \"entries\": %d, \"res0\": %.17g, \"res_last\": %.17g}\n", n,
                        (int)cres.size(), cres.front(), cres.back());
        }
    } catch (const pmg::Error &e) {
        std::printf("{\"test\": \"error\", \"status\": %d, \"what\": \"%s\"}\n", (int)e.status, e.what());
        return 3;
    }
    return 0;
}
