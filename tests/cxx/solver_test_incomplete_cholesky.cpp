// test/solver_test_incomplete_cholesky.f90 restated against sigma.hpp: a random
// weighted graph Laplacian + I (nn = 128) in a csr_matrix; ldu(incomplete, level 0)
// as a stand-alone stationary solver (10 nn sweeps, bar 1e-14, :182-202) and as the
// preconditioner of cg(1e-16) (bar 1e-15, :213-226).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && (!std::strcmp(argv[1], "-v") || !std::strcmp(argv[1], "-V") || !std::strcmp(argv[1], "--verbose"));
    const int nn = 128;
    const dp p = std::log(1.0 * nn) / std::log(2.0) / nn;
    rng64 rnd(2718);

    ll_graph g;                                  // :60-72
    g.init(nn);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);
        for (int j = i + 1; j <= nn; j++)
            if (rnd.next() < p) {
                g.add_edge(i, j);
                g.add_edge(j, i);
            }
    }
    if (verbose) std::printf(" o Done generating random graph: %d edges, max degree %d\n", g.get_num_edges(), g.get_max_degree());

    csr_matrix A;                                // :88-107
    A.init(nn, nn);
    A.copy_graph(g);
    A.zero();
    for (int i = 1; i <= nn; i++) {
        A.add_value(i, i, 1.0);
        for (int32_t j : g.get_neighbors(i)) {
            const dp z = rnd.next();
            if (j > i) {
                A.set_value(i, j, -z);
                A.add_value(i, i, +z);
                A.set_value(j, i, -z);
                A.add_value(j, j, +z);
            }
        }
    }

    linear_solver *solver = cg(1.0e-16), *pc = ldu(true, 0);   // :115-119
    solver->setup(A);
    pc->setup(A);
    solver->set_max_iterations(50 * nn);          // safety net (not in the reference); must not trigger

    std::vector<dp> u(nn, 0.0), v(nn), f(nn, 0.0), r(nn, 0.0), q(nn, 0.0);
    for (dp &e : v) e = rnd.next();
    A.matvec(v.data(), q.data());                 // :140-144: smooth v with one application
    for (int i = 0; i < nn; i++) r[i] = v[i] - q[i];
    pc->solve(A, v.data(), r.data());
    A.matvec(v.data(), f.data());

    // incomplete Cholesky as a solver (:155-176)
    r = f;
    std::fill(q.begin(), q.end(), 0.0);
    for (int n = 1; n <= 10 * nn; n++) {
        pc->solve(A, q.data(), r.data());
        for (int i = 0; i < nn; i++) u[i] = u[i] + q[i];
        A.matvec(u.data(), q.data());
        for (int i = 0; i < nn; i++) r[i] = f[i] - q[i];
    }
    dp misfit = 0;
    for (int i = 0; i < nn; i++) misfit = std::fmax(misfit, std::fabs(u[i] - v[i]));
    if (misfit > 1.0e-14) {
        std::printf(" Incomplete Cholesky failed to produce sufficiently accurate solution after %d iterations.\n Error: %g\n", 10 * nn, misfit);
        return 1;
    }
    if (verbose) std::printf(" o Done solving with incomplete Cholesky.\n     Error: %g\n", misfit);

    // incomplete Cholesky as a preconditioner (:207-226)
    std::fill(u.begin(), u.end(), 0.0);
    solver->solve(A, u.data(), f.data(), pc);
    if (solver->capped()) { std::printf(" ILDU-preconditioned CG hit the safety cap\n"); return 1; }
    misfit = 0;
    for (int i = 0; i < nn; i++) misfit = std::fmax(misfit, std::fabs(u[i] - v[i]));
    if (misfit > 1.0e-15) {
        std::printf(" ILDU-preconditioned conjugate gradient method failed to produce sufficiently accurate solution.\n Error: %g\n", misfit);
        return 1;
    }
    if (verbose) std::printf(" o Done solving with ILDU-preconditioned CG: %d iterations.\n     Error: %g\n", solver->iterations, misfit);

    delete solver;
    delete pc;
    return 0;
}
