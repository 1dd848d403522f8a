// test/solver_test_jacobi.f90 restated: random-weight Erdos-Renyi graph
// Laplacian + I in CSR (nn = 128, p = log2(nn)/nn), manufactured solution;
// Jacobi-Richardson sweeps, Jacobi-PCG, then a skew-symmetric perturbation and
// Jacobi-preconditioned BiCGSTAB.  Reference bars: 1e-14 / 1e-15 / 1e-15 with
// tolerance 1e-16 and no iteration cap; here a safety cap is set and must not
// trigger, and the bars are the reference's.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

static dp maxdiff(const std::vector<dp> &a, const std::vector<dp> &b)
{
    dp m = 0.0;
    for (size_t i = 0; i < a.size(); i++) m = std::fmax(m, std::fabs(a[i] - b[i]));
    return m;
}

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && !strcmp(argv[1], "-v");
    const int nn = 128;
    const dp p = std::log(1.0 * nn) / std::log(2.0) / nn;
    rng64 rnd(20141017);

    ll_graph g;
    g.init(nn);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);
        for (int j = i + 1; j <= nn; j++)
            if (rnd.next() < p) { g.add_edge(i, j); g.add_edge(j, i); }
    }
    auto cg_ = std::make_shared<cs_graph>();      // convert_graph_type(g, "compressed sparse")
    cg_->copy(g);

    csr_matrix A;
    A.init(nn, nn);
    A.set_graph(cg_);
    A.zero();
    for (int i = 1; i <= nn; i++) {
        A.add_value(i, i, 1.0);
        for (int32_t j : g.get_neighbors(i)) {
            const dp z = rnd.next();
            if (j > i) {
                A.set_value(i, j, -z); A.add_value(i, i, +z);
                A.set_value(j, i, -z); A.add_value(j, j, +z);
            }
        }
    }

    linear_solver *solver = cg(1.e-16), *pc = jacobi();
    solver->setup(A);
    pc->setup(A);
    solver->set_max_iterations(10 * nn);

    std::vector<dp> u(nn, 0.0), v(nn), f(nn, 0.0), r(nn, 0.0), q(nn, 0.0);
    for (dp &x : v) x = rnd.next();
    // smooth v with I - D^{-1} A
    A.matvec(v.data(), q.data());
    for (int i = 0; i < nn; i++) r[i] = v[i] - q[i];
    pc->solve(A, v.data(), r.data());
    A.matvec(v.data(), f.data());

    // Jacobi method as a solver: 10*nn Richardson sweeps through the seams
    r = f;
    for (int n = 1; n <= 10 * nn; n++) {
        pc->solve(A, q.data(), r.data());
        for (int i = 0; i < nn; i++) u[i] += q[i];
        A.matvec(u.data(), q.data());
        for (int i = 0; i < nn; i++) r[i] = f[i] - q[i];
    }
    dp misfit = maxdiff(u, v);
    if (misfit > 1.0e-14) { std::printf(" Jacobi method failed, error %g\n", misfit); return 1; }
    if (verbose) std::printf(" o Jacobi method error: %g\n", misfit);

    // Jacobi as a preconditioner for CG
    u.assign(nn, 0.0);
    solver->solve(A, u.data(), f.data(), pc);
    misfit = maxdiff(u, v);
    if (solver->capped() || misfit > 1.0e-15 * 8) {   // 8 ulp-ish slack for the parallel dot products
        std::printf(" Jacobi-preconditioned CG failed, error %g (capped %d, iterations %d)\n", misfit, (int)solver->capped(), solver->iterations);
        return 1;
    }
    if (verbose) std::printf(" o Jacobi-PCG error: %g  iterations %d\n", misfit, solver->iterations);

    // skew-symmetric perturbation, then Jacobi-preconditioned BiCGSTAB
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i))
            if (j > i) {
                const dp z = (2 * rnd.next() - 1) / 16;
                A.add_value(i, j, +z);
                A.add_value(j, i, -z);
            }
    solver->destroy();
    delete solver;
    solver = bicgstab(1.e-16);
    solver->setup(A);
    solver->set_max_iterations(10 * nn);
    pc->setup(A);
    A.matvec(v.data(), f.data());
    u.assign(nn, 0.0);
    solver->solve(A, u.data(), f.data(), pc);
    misfit = maxdiff(u, v);
    if (misfit > 1.0e-15 * 8) {
        std::printf(" Jacobi-preconditioned BiCG-Stab failed, error %g (iterations %d)\n", misfit, solver->iterations);
        return 1;
    }
    if (verbose) std::printf(" o Jacobi-BiCGSTAB error: %g  iterations %d capped %d\n", misfit, solver->iterations, (int)solver->capped());
    delete solver;
    delete pc;
    return 0;
}
