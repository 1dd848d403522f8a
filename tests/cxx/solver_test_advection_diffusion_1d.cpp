// test/solver_test_advection_diffusion_1d.f90 restated: BiCGSTAB(1e-12) on the
// nonsymmetric ELLPACK operator -d^2/dx^2 + c d/dx, nn = 1024, bar 1e-8.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
using namespace sigma;

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && !strcmp(argv[1], "-v");
    const int nn = 1024;
    const dp dx = 1.0 / (nn + 1), c = 0.5;

    ll_graph g;
    g.init(nn, nn);
    for (int i = 1; i <= nn - 1; i++) {
        g.add_edge(i, i);
        g.add_edge(i, i + 1);
        g.add_edge(i + 1, i);
    }
    g.add_edge(nn, nn);

    ellpack_matrix A;
    A.init(nn, nn);
    A.copy_graph(g);
    A.zero();
    for (int i = 1; i <= nn - 1; i++) {
        A.set_value(i, i, +2.0);
        A.set_value(i, i + 1, -1.0 + c * dx / 2);
        A.set_value(i + 1, i, -1.0 - c * dx / 2);
    }
    A.set_value(nn, nn, 2.0);

    std::vector<dp> u(nn, 0.0), v(nn), f(nn, 2.0 * dx * dx);
    for (int i = 1; i <= nn; i++) {
        const dp x = i * dx;
        v[i - 1] = 2.0 * (x - (std::exp(c * x) - 1) / (std::exp(c) - 1)) / c;
    }

    linear_solver *solver = bicgstab(1.0e-12);
    solver->setup(A);
    solver->set_max_iterations(20 * nn);
    solver->solve(A, u.data(), f.data());

    dp misfit = 0.0;
    for (int i = 0; i < nn; i++) misfit = std::fmax(misfit, std::fabs(u[i] - v[i]));
    if (solver->capped() || misfit > 1.0e-8) {
        std::printf(" BiCG-Stab solver failed.\n Should have error < %g\n Error found: %g\n", 1.0e-8, misfit);
        return 1;
    }
    if (verbose) std::printf(" Error: %g  iterations: %d\n", misfit, solver->iterations);
    delete solver;
    return 0;
}
