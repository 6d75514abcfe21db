// pmg_runner -- the reference's experiment runners as a real command-line tool on top of include/pmg.hpp.
//
// The reference has no argv: parameters are edited in 2_part_MG/main.cpp / 3_part_parallel/main.cu and the
// program is recompiled.  This tool runs the same PROTOCOLS on libpmg.so and writes the same OUTPUT_RESULT/*.txt
// formats, so the reference's python_plot/*.py scripts work on its output unchanged:
//
//   cycles     MultigridTestRunner::run_all_cycles            (MultiGridTestRunner.hpp:28-38)   stdout only
//   time_h     ...::run_all_cycles_time_h                     (:59-90)   timings_{v,w,f}_cycle.txt        "N seconds"
//   err_h      ...::run_all_cycles_err_h                      (:40-57)   h_errors_{v,w,f}_cycle<it>.txt   "N relerr"
//   err_norm   ...::run_all_cycles_err_norm_iteration         (:110-122) error_{v,w,f}_cycle<it>.txt      one value per line
//   gpu        ParallelTestRunner::run_all_cycles             (ParallelTestRunner.cu:127-141)
//                                                              timings_parallel_{v,w}_cycle.txt             "N seconds"
//   ops        ParallelTestRunner::plotTimeSequentialVsParallel (:98-125), device side only
//                                                              timings_{residual,jacobi,restriction,prolungator}_gpu.txt
//                                                                                                           "num_thread N seconds"
//   err_vector ...::run_all_cycles with flag_err_vector_iteration (MultiGridTestRunner.hpp:150-160, 214-222, 247-248): the error
//              FIELD phi - phi_exact before the first and after every cycle, written with
//              save_errors_vector_to_file_last_iteration_gpu (save_vector_err_file.hpp:67-89)
//                                                              ERR_VECTOR/iteration_last_gpu.txt (under ./OUTPUT_RESULT)
//   smoother   SolverRunner-style study of one smoother (Smoother::test hook, Smoother.hpp:50-57,100-115): --smoother
//              jacobi|gs|rbgs|chebyshev|cg, --iters sweeps; residual and error norms per sweep to
//              smoother_<name>_N<n>.txt "sweep residual relerr", error fields to ERR_VECTOR/iteration_<10 i>.txt (jacobi)
//   history    (new) residual history to a relative tolerance (V, W, FMG-then-V): history_<cycle>_N<n>.txt  "cycle norm"
//
// usage: pmg_runner <mode> [--n 33,65,...] [--iters K] [--alpha A] [--omega W] [--eps E] [--tol T]
//                          [--prolong reference|full] [--out DIR]
// defaults = 2_part_MG/main.cpp: n = 33..1025, iters = 1, alpha = 3, omega = 1, eps = 1e-7.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "pmg.hpp"

using namespace pmg::compat;

struct Options {
    std::string mode = "cycles", out = "OUTPUT_RESULT";
    std::vector<int> n_list = {33, 65, 129, 257, 513, 1025};
    int iters = 1, alpha = 3, prolong = PMG_PROLONG_REFERENCE;
    double omega = 1.0, eps = 1e-7, tol = 1e-8;
    std::string smoother = "jacobi";
};

static void manufactured(std::vector<double> &f, std::vector<double> &u, int n)
{
    // DynamicGridUtils::compute_rhs / compute_exact_solution with a = p = q = 1 (DynamicGridUtils.hpp:97-124)
    const double h = 1.0 / (n - 1), factor = (M_PI * M_PI) * 2.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double x = i * h, y = j * h;
            u[(size_t)j * n + i] = std::sin(M_PI * x) * std::sin(M_PI * y);
            f[(size_t)j * n + i] = factor * std::sin(M_PI * x) * std::sin(M_PI * y);
        }
}

static double norm(const std::vector<double> &v)
{
    double s = 0.0;
    for (double x : v) s += x * x;
    return std::sqrt(s);
}

static double rel_error(const std::vector<double> &phi, const std::vector<double> &u)
{
    std::vector<double> e(phi.size());
    for (size_t i = 0; i < phi.size(); ++i) e[i] = phi[i] - u[i];
    return norm(e) / norm(u);
}

struct CycleResult {
    double rel_err, seconds;
    std::vector<double> err_per_iter;  // relative L2 error before the first and after every cycle
};

// MultigridTestRunner::run_cycle (MultiGridTestRunner.hpp:127-264)
static CycleResult run_cycle(const Options &o, const std::string &name, int n)
{
    std::vector<double> phi((size_t)n * n, 0.0), f(phi.size()), u(phi.size());
    manufactured(f, u, n);
    WeightedJacobiSmoother smoother(o.eps, o.omega);
    MultigridSolver mg(&smoother, o.alpha, n);
    mg.prolong_mode = o.prolong;
    const double h = 1.0 / (n - 1);
    CycleResult r;
    r.err_per_iter.push_back(rel_error(phi, u));
    std::cout << "Running " << name << "...\n";
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int it = 0; it < o.iters; ++it) {
        if (name == "F-cycle")
            mg.f_cycle_from_fine(phi.data(), f.data(), n);
        else if (name == "W-cycle")
            mg.w_cycle(phi.data(), f.data(), n, h);
        else
            mg.v_cycle(phi.data(), f.data(), n, h);
        r.err_per_iter.push_back(rel_error(phi, u));
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    r.seconds = std::chrono::duration<double>(t1 - t0).count();
    r.rel_err = r.err_per_iter.back();
    std::cout << "  Final Relative L2 Error: " << r.rel_err << "\n";
    std::cout << "  Elapsed Time: " << r.seconds << " seconds\n";
    return r;
}

static void write_pairs(const std::string &path, const std::vector<std::pair<int, double>> &v)
{
    std::ofstream f(path);
    for (auto &p : v) f << p.first << " " << p.second << "\n";
}

static std::vector<int> parse_list(const char *s)
{
    std::vector<int> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::atoi(tok.c_str()));
    return v;
}

static double device_seconds(void (*fn)(void *), void *ctx)
{
    pmg::check(pmg_device_synchronize());
    auto t0 = std::chrono::high_resolution_clock::now();
    fn(ctx);
    pmg::check(pmg_device_synchronize());
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

struct OpCtx {
    double *x, *f, *r, *c;
    int n, nc;
    double h;
};

int main(int argc, char **argv)
{
    Options o;
    if (argc > 1 && argv[1][0] != '-') o.mode = argv[1];
    for (int i = 1; i + 1 < argc; ++i) {
        std::string a = argv[i];
        if (a == "--n") o.n_list = parse_list(argv[++i]);
        else if (a == "--iters") o.iters = std::atoi(argv[++i]);
        else if (a == "--alpha") o.alpha = std::atoi(argv[++i]);
        else if (a == "--omega") o.omega = std::atof(argv[++i]);
        else if (a == "--eps") o.eps = std::atof(argv[++i]);
        else if (a == "--tol") o.tol = std::atof(argv[++i]);
        else if (a == "--out") o.out = argv[++i];
        else if (a == "--smoother") o.smoother = argv[++i];
        else if (a == "--prolong") o.prolong = std::strcmp(argv[++i], "full") == 0 ? PMG_PROLONG_FULL : PMG_PROLONG_REFERENCE;
    }
    mkdir(o.out.c_str(), 0755);
    try {
        const char *names[3] = {"V-cycle", "W-cycle", "F-cycle"};
        if (o.mode == "cycles") {
            for (int n : o.n_list) {
                std::cout << "\n=== Multigrid Solution for N = " << n << " ===\n";
                for (auto nm : names) run_cycle(o, nm, n);
            }
        } else if (o.mode == "time_h" || o.mode == "err_h") {
            std::vector<std::pair<int, double>> t[3], e[3];
            for (int n : o.n_list) {
                std::cout << "\n=== Multigrid Solution for N = " << n << " ===\n";
                for (int k = 0; k < 3; ++k) {
                    CycleResult r = run_cycle(o, names[k], n);
                    t[k].push_back({n, r.seconds});
                    e[k].push_back({n, r.rel_err});
                }
            }
            const char *tag[3] = {"v", "w", "f"};
            for (int k = 0; k < 3; ++k) {
                if (o.mode == "time_h")
                    write_pairs(o.out + "/timings_" + tag[k] + "_cycle.txt", t[k]);
                else
                    write_pairs(o.out + "/h_errors_" + tag[k] + "_cycle" + std::to_string(o.iters) + ".txt", e[k]);
            }
        } else if (o.mode == "err_norm") {
            const char *tag[3] = {"v", "w", "f"};
            for (int n : o.n_list)
                for (int k = 0; k < 3; ++k) {
                    CycleResult r = run_cycle(o, names[k], n);
                    std::ofstream f(o.out + "/error_" + tag[k] + "_cycle" + std::to_string(o.iters) + ".txt");
                    for (double v : r.err_per_iter) f << v << "\n";
                }
        } else if (o.mode == "gpu") {
            // ParallelTestRunner::run_v_cycle / run_w_cycle (ParallelTestRunner.cu:152-228): device-resident phi, f
            std::vector<std::pair<int, double>> tv, tw;
            for (int n : o.n_list) {
                size_t bytes = (size_t)n * n * sizeof(double);
                std::vector<double> phi((size_t)n * n, 0.0), f(phi.size()), u(phi.size());
                manufactured(f, u, n);
                void *dx, *df;
                pmg::check(pmg_device_alloc(&dx, bytes));
                pmg::check(pmg_device_alloc(&df, bytes));
                for (int w = 0; w < 2; ++w) {
                    pmg::check(pmg_memcpy(dx, phi.data(), bytes, 1, 0));
                    pmg::check(pmg_memcpy(df, f.data(), bytes, 1, 0));
                    ParallelMultiGridSolver mg(o.alpha);
                    mg.omega = o.omega;
                    mg.prolong_mode = o.prolong;
                    mg.v_cycle((double *)dx, (double *)df, n, 1.0 / (n - 1));  // builds the hierarchy (not timed)
                    pmg::check(pmg_memcpy(dx, phi.data(), bytes, 1, 0));
                    auto t0 = std::chrono::high_resolution_clock::now();
                    for (int it = 0; it < o.iters; ++it) {
                        if (w)
                            mg.w_cycle((double *)dx, (double *)df, n, 1.0 / (n - 1));
                        else
                            mg.v_cycle((double *)dx, (double *)df, n, 1.0 / (n - 1));
                    }
                    pmg::check(pmg_device_synchronize());
                    double s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
                    std::vector<double> out((size_t)n * n);
                    pmg::check(pmg_memcpy(out.data(), dx, bytes, 0, 1));
                    std::cout << "N = " << n << (w ? " W" : " V") << "-cycle x" << o.iters << ": " << s
                              << " s, Final Relative L2 Error: " << rel_error(out, u) << "\n";
                    (w ? tw : tv).push_back({n, s});
                }
                pmg_device_free(dx);
                pmg_device_free(df);
            }
            write_pairs(o.out + "/timings_parallel_v_cycle.txt", tv);
            write_pairs(o.out + "/timings_parallel_w_cycle.txt", tw);
        } else if (o.mode == "ops") {
            // run_residual / run_jacobi (100 -> 101 sweeps) / run_restriction / run_prolungator, device side
            std::ofstream fr(o.out + "/timings_residual_gpu.txt"), fj(o.out + "/timings_jacobi_gpu.txt"),
                fs(o.out + "/timings_restriction_gpu.txt"), fp(o.out + "/timings_prolungator_gpu.txt");
            for (int n : o.n_list) {
                OpCtx c;
                c.n = n;
                c.nc = (n - 1) / 2 + 1;
                c.h = 1.0 / (n - 1);
                size_t bytes = (size_t)n * n * sizeof(double);
                std::vector<double> f((size_t)n * n), u(f.size()), z(f.size(), 0.0);
                manufactured(f, u, n);
                pmg::check(pmg_device_alloc((void **)&c.x, bytes));
                pmg::check(pmg_device_alloc((void **)&c.f, bytes));
                pmg::check(pmg_device_alloc((void **)&c.r, bytes));
                pmg::check(pmg_device_alloc((void **)&c.c, (size_t)c.nc * c.nc * sizeof(double)));
                pmg::check(pmg_memcpy(c.x, z.data(), bytes, 1, 0));
                pmg::check(pmg_memcpy(c.f, f.data(), bytes, 1, 0));
                pmg::check(pmg_memcpy(c.r, z.data(), bytes, 1, 0));
                pmg::check(pmg_memcpy(c.c, z.data(), (size_t)c.nc * c.nc * sizeof(double), 1, 0));
                Parallel::ComputeJacobi(c.x, c.f, n, n, c.h, 8);  // warm-up (also builds the blocked path scratch)
                pmg::check(pmg_memcpy(c.x, z.data(), bytes, 1, 0));
                double tj = device_seconds([](void *p) { OpCtx *q = (OpCtx *)p; Parallel::ComputeJacobi(q->x, q->f, q->n, q->n, q->h, 100); }, &c);
                double tr = device_seconds([](void *p) { OpCtx *q = (OpCtx *)p; Parallel::ComputeResidual(q->r, q->x, q->f, q->n, q->n, q->h); }, &c);
                double ts = device_seconds([](void *p) { OpCtx *q = (OpCtx *)p; Parallel::ComputeRestriction(q->r, q->c, q->n, q->nc); }, &c);
                double tp = device_seconds([](void *p) { OpCtx *q = (OpCtx *)p; Parallel::ComputeProlungator(q->c, q->r, q->nc, q->n); }, &c);
                fr << 32 << " " << n << " " << tr << "\n";
                fj << 32 << " " << n << " " << tj << "\n";
                fs << 32 << " " << n << " " << ts << "\n";
                fp << 32 << " " << n << " " << tp << "\n";
                std::cout << "N = " << n << ": jacobi x101 " << tj << " s (" << 24.0 * n * n * 101 / tj / 1e9
                          << " GB/s), residual " << tr << " s, restriction " << ts << " s, prolongation " << tp << " s\n";
                pmg_device_free(c.x);
                pmg_device_free(c.f);
                pmg_device_free(c.r);
                pmg_device_free(c.c);
            }
        } else if (o.mode == "err_vector") {
            for (int n : o.n_list)
                for (int k = 0; k < 3; ++k) {
                    std::vector<double> phi((size_t)n * n, 0.0), f(phi.size()), u(phi.size());
                    manufactured(f, u, n);
                    WeightedJacobiSmoother smoother(o.eps, o.omega);
                    MultigridSolver mg(&smoother, o.alpha, n);
                    mg.prolong_mode = o.prolong;
                    std::vector<std::vector<double>> err_vect_iteration;
                    auto snap = [&]() {
                        std::vector<double> e(phi.size());
                        for (size_t i = 0; i < phi.size(); ++i) e[i] = phi[i] - u[i];
                        err_vect_iteration.push_back(std::move(e));
                    };
                    snap();
                    for (int it = 0; it < o.iters; ++it) {
                        if (k == 2)
                            mg.f_cycle_from_fine(phi.data(), f.data(), n);
                        else if (k == 1)
                            mg.w_cycle(phi.data(), f.data(), n, 1.0 / (n - 1));
                        else
                            mg.v_cycle(phi.data(), f.data(), n, 1.0 / (n - 1));
                        snap();
                    }
                    save_errors_vector_to_file_last_iteration_gpu(err_vect_iteration);
                    std::cout << names[k] << " N = " << n << ": error field after " << o.iters
                              << " cycle(s) in ./OUTPUT_RESULT/ERR_VECTOR/iteration_last_gpu.txt, ||e||/||u|| = "
                              << rel_error(phi, u) << "\n";
                }
        } else if (o.mode == "smoother") {
            for (int n : o.n_list) {
                std::vector<double> x((size_t)n * n, 0.0), f(x.size()), u(x.size()), res, err;
                manufactured(f, u, n);
                const double h = 1.0 / (n - 1);
                WeightedJacobiSmoother jac(0.0, o.omega);
                GaussSeidelSmoother gs(0.0);
                RedBlackGaussSeidelSmoother rb;
                ChebyshevJacobiSmoother ch;
                ConjugateGradientSmoother cg(0.0);
                if (o.smoother == "jacobi") {
                    jac.switch_test_mode();  // error fields every 10 sweeps -> ./OUTPUT_RESULT/ERR_VECTOR/iteration_<10 i>.txt
                    jac.smooth(x.data(), f.data(), n, n, h, o.iters - 1, u.data(), &res, &err);
                } else if (o.smoother == "gs") {
                    gs.smooth(x.data(), f.data(), n, n, h, o.iters, u.data(), &res);
                } else if (o.smoother == "rbgs") {
                    rb.smooth(x.data(), f.data(), n, n, h, o.iters, u.data(), &res);
                } else if (o.smoother == "chebyshev") {
                    ch.smooth(x.data(), f.data(), n, n, h, o.iters - 1, u.data(), &res);
                } else if (o.smoother == "cg") {
                    cg.smooth(x.data(), f.data(), n, n, h, o.iters, u.data(), &res);
                } else {
                    std::fprintf(stderr, "unknown smoother %s\n", o.smoother.c_str());
                    return 2;
                }
                std::ofstream out(o.out + "/smoother_" + o.smoother + "_N" + std::to_string(n) + ".txt");
                out.precision(17);
                for (size_t i = 0; i < res.size(); ++i) {
                    out << i << " " << res[i];
                    if (i < err.size()) out << " " << err[i];
                    out << "\n";
                }
                std::cout << o.smoother << " N = " << n << ": " << res.size() << " residual norms, last " << res.back()
                          << ", final relative L2 error " << rel_error(x, u) << "\n";
            }
        } else if (o.mode == "history") {
            for (int n : o.n_list) {
                std::vector<double> f((size_t)n * n), u(f.size());
                manufactured(f, u, n);
                for (int k = 0; k < 3; ++k) {  // V, W, and one general full-multigrid pass followed by V-cycles
                    pmg_config c;
                    pmg_config_default(&c, n);
                    c.omega = o.omega;
                    c.gamma = o.alpha;
                    c.prolong_mode = o.prolong;
                    pmg::Solver s(c);
                    s.set_rhs(f.data());
                    s.zero_guess();
                    const pmg_cycle_kind kinds[3] = {PMG_CYCLE_V, PMG_CYCLE_W, PMG_CYCLE_FMG};
                    const char *tags[3] = {"v", "w", "fmg"};
                    std::vector<double> hist = s.solve(kinds[k], o.tol, 200);
                    std::ofstream out(o.out + "/history_" + tags[k] + "_N" + std::to_string(n) + ".txt");
                    out.precision(17);
                    for (size_t i = 0; i < hist.size(); ++i) out << i << " " << hist[i] << "\n";
                    std::cout << "N = " << n << " " << tags[k] << ": " << hist.size() - 1 << " cycles to "
                              << hist.back() / hist.front() << "\n";
                }
            }
        } else {
            std::fprintf(stderr, "unknown mode %s\n", o.mode.c_str());
            return 2;
        }
    } catch (const pmg::Error &e) {
        std::fprintf(stderr, "pmg_runner: %s\n", e.what());
        return 3;
    }
    return 0;
}
