// test/linear_operator_test_algebra.f90 restated against sigma.hpp: sums,
// products and adjoints of a csr_matrix A and a csc_matrix B on random graphs
// (nn = 64, p = log2(nn)/nn, entries 2q-1), with the reference's own bars:
// get_value of the sum 1e-14 (:157-165), (A+B)x 1e-14 (:190-196), (A*B)x 1e-14
// (:232-239), adjoint(A)x 1e-12 (:255-261), (adjoint(A)*A)x 1e-12 (:277-283).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

static dp max_abs_diff(const std::vector<dp> &a, const std::vector<dp> &b)
{
    dp r = 0;
    for (size_t i = 0; i < a.size(); i++) r = std::fmax(r, std::fabs(a[i] - b[i]));
    return r;
}
static dp max_abs(const std::vector<dp> &a)
{
    dp r = 0;
    for (dp v : a) r = std::fmax(r, std::fabs(v));
    return r;
}

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && (!std::strcmp(argv[1], "-v") || !std::strcmp(argv[1], "-V") || !std::strcmp(argv[1], "--verbose"));
    const int nn = 64;
    const dp p = std::log(1.0 * nn) / std::log(2.0) / nn;
    rng64 rnd(2024);

    // random graphs (:83-93)
    ll_graph g, h;
    g.init(nn);
    h.init(nn);
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn; j++) {
            if (rnd.next() < p) g.add_edge(i, j);
            if (rnd.next() < p) h.add_edge(j, i);
        }
    if (verbose) std::printf(" Done building random graphs.\n     Number of edges: %d %d\n", g.get_num_edges(), h.get_num_edges());

    // random matrices on those graphs (:109-143)
    csr_matrix A;
    csc_matrix B;
    A.init(nn, nn); A.copy_graph(g);
    B.init(nn, nn); B.copy_graph(h);
    A.add_reference();   // the caller's own reference: A and B live on the stack
    B.add_reference();
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i)) A.set_value(i, j, 2 * rnd.next() - 1);
    for (int i = 1; i <= nn; i++)
        for (int32_t j : h.get_neighbors(i)) B.set_value(i, j, 2 * rnd.next() - 1);

    std::vector<dp> w(nn), x(nn, 1.0), y(nn), z(nn);

    // ---- L = A + B (:151-196) ---------------------------------------------
    linear_operator *L = A + B;
    if (verbose) std::printf(" Testing operator sum\n");
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn; j++)
            if (std::fabs(L->get_value(i, j) - A.get_value(i, j) - B.get_value(i, j)) > 1.0e-14) {
                std::printf(" Getting entry %d %d from operator sum failed.\n", i, j);
                return 1;
            }
    std::fill(z.begin(), z.end(), 0.0);
    A.matvec_add(x.data(), z.data());
    B.matvec_add(x.data(), z.data());
    if (max_abs(z) == 0.0) { std::printf(" (A+B)*[1,...,1] = 0. Terminating.\n"); return 1; }
    std::fill(y.begin(), y.end(), 0.0);
    L->matvec_add(x.data(), y.data());
    if (max_abs(y) == 0.0) { std::printf(" L = A+B, y = L*[1,..,1] failed, got y = 0.\n"); return 1; }
    dp r = max_abs_diff(y, z);
    if (r > 1.0e-14) { std::printf(" L = A+B: ||y-z|| = %g\n", r); return 1; }
    L->destroy();
    delete L;

    // ---- L = A * B (:204-239) ---------------------------------------------
    L = A * B;
    if (verbose) std::printf(" Testing operator product\n");
    B.matvec(x.data(), w.data());
    A.matvec(w.data(), z.data());
    if (max_abs(z) == 0.0) { std::printf(" A*B*[1,..,1] = 0. Terminating.\n"); return 1; }
    std::fill(y.begin(), y.end(), 0.0);
    L->matvec(x.data(), y.data());
    if (max_abs(y) == 0.0) { std::printf(" L = A*B, y = L*[1,..,1] failed, got y = 0.\n"); return 1; }
    r = max_abs_diff(y, z);
    if (r > 1.0e-14) { std::printf(" L = A*B: ||y-z|| = %g\n", r); return 1; }
    L->destroy();
    delete L;

    // ---- L = adjoint(A) (:247-261) ----------------------------------------
    L = adjoint(A);
    if (verbose) std::printf(" Testing operator adjoint\n");
    L->matvec(x.data(), y.data());
    A.matvec_t(x.data(), z.data());
    r = max_abs_diff(y, z);
    if (r > 1.0e-12) { std::printf(" L = A*: || Lx - A*x || = %g\n", r); return 1; }
    if (L->get_value(3, 7) != A.get_value(7, 3)) { std::printf(" adjoint get_value failed\n"); return 1; }
    L->destroy();
    delete L;

    // ---- L = adjoint(A) * A (:269-283) ------------------------------------
    linear_operator *At = adjoint(A);
    L = *At * A;
    if (verbose) std::printf(" Testing adjoint and product\n");
    L->matvec(x.data(), y.data());
    A.matvec(x.data(), w.data());
    A.matvec_t(w.data(), z.data());
    r = max_abs_diff(y, z);
    if (r > 1.0e-12) { std::printf(" L = A*A: || Lx - A*Ax || = %g\n", r); return 1; }

    // host mutation of an operand reaches the expression (dirty-mirror protocol)
    A.scalar_multiply(2.0);
    L->matvec(x.data(), w.data());
    for (int i = 0; i < nn; i++)
        if (w[i] != 4.0 * y[i]) { std::printf(" stale operand mirror inside an expression\n"); return 1; }

    // a solver driven by an expression: (A^T A + B^T B + I) u = f with CG
    linear_operator *Bt = adjoint(B);
    linear_operator *BtB = *Bt * B;
    linear_operator *N = *L + *BtB;
    csr_matrix I;
    ll_graph gi;
    gi.init(nn);
    for (int i = 1; i <= nn; i++) gi.add_edge(i, i);
    I.init(nn, nn); I.copy_graph(gi);
    I.add_reference();
    for (int i = 1; i <= nn; i++) I.set_value(i, i, 1.0);
    linear_operator *S = *N + I;
    std::vector<dp> v(nn), f(nn), u(nn, 0.0);
    for (dp &e : v) e = rnd.next();
    S->matvec(v.data(), f.data());
    linear_solver *s = cg(1e-12);
    S->set_solver(s);
    S->solve(u.data(), f.data());
    r = max_abs_diff(u, v);
    if (verbose) std::printf(" CG on A^T A + B^T B + I: %d iterations, error %g\n", s->iterations, r);
    if (r > 1.0e-9) { std::printf(" CG on an operator expression failed: %g\n", r); return 1; }
    delete s;

    S->destroy();   // tears the whole expression tree down; A, B, I survive (own references)
    delete S;
    if (A.reference_count != 1 || B.reference_count != 1 || I.reference_count != 1) {
        std::printf(" reference counts after destroy: %d %d %d\n", A.reference_count, B.reference_count, I.reference_count);
        return 1;
    }
    return 0;
}
