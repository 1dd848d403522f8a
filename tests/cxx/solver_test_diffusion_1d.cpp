// test/solver_test_diffusion_1d.f90 restated against sigma.hpp: CG on the
// ELLPACK matrix of -d^2/dx^2, nn = 127, exact solution x(1-x), bar 1e-14.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
using namespace sigma;

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && (!strcmp(argv[1], "-v") || !strcmp(argv[1], "-V") || !strcmp(argv[1], "--verbose"));
    const int nn = 127;
    const dp dx = 1.0 / (nn + 1);

    ll_graph g;
    g.init(nn, nn);
    for (int i = 1; i <= nn - 1; i++) {
        g.add_edge(i, i);
        g.add_edge(i, i + 1);
        g.add_edge(i + 1, i);
    }
    g.add_edge(nn, nn);

    ellpack_matrix A;                       // convert_graph_type(g, "ellpack") + A%set_graph(g)
    A.init(nn, nn);
    A.copy_graph(g);
    A.zero();
    for (int i = 1; i <= nn - 1; i++) {
        A.set_value(i, i, +2.0);
        A.set_value(i, i + 1, -1.0);
        A.set_value(i + 1, i, -1.0);
    }
    A.set_value(nn, nn, 2.0);
    if (verbose) std::printf(" Done creating matrix for 1d Laplace operator\n");

    std::vector<dp> u(nn, 0.0), v(nn), f(nn, 2.0 * dx * dx);
    for (int i = 1; i <= nn; i++) v[i - 1] = i * dx * (1.0 - i * dx);

    linear_solver *solver = cg(1.e-16);
    solver->setup(A);
    solver->set_max_iterations(20 * nn);    // safety net only (the reference loop has no cap)
    solver->solve(A, u.data(), f.data());

    dp misfit = 0.0;
    for (int i = 0; i < nn; i++) misfit = std::fmax(misfit, std::fabs(u[i] - v[i]));
    if (solver->capped() || misfit > 1.0e-14) {
        std::printf(" CG solver failed.\n Should have error < %g\n Error found: %g (iterations %d)\n", 1.0e-14, misfit, solver->iterations);
        return 1;
    }
    if (verbose) std::printf(" Error: %g  iterations: %d\n", misfit, solver->iterations);
    solver->destroy();
    delete solver;
    return 0;
}
