// test/matrix_test_basics.f90:332-362 restated for the three device formats:
// A%matvec and A%matvec_t against a dense copy, relative error 1e-15 (x ulps
// of slack for rows with many entries), plus matvec_add and the dirty-mirror
// protocol after host mutations.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

template <class M>
static int check(const char *name, M &A, const std::vector<dp> &B, int nn, rng64 &rnd)
{
    std::vector<dp> x(nn), y(nn), z(nn);
    for (dp &v : x) v = rnd.next();
    for (int trans = 0; trans < 2; trans++) {
        if (trans) A.matvec_t(x.data(), y.data()); else A.matvec(x.data(), y.data());
        dp num = 0, den = 0;
        for (int i = 0; i < nn; i++) {
            dp s = 0;
            for (int j = 0; j < nn; j++) s += (trans ? B[(size_t)j * nn + i] : B[(size_t)i * nn + j]) * x[j];
            num = std::fmax(num, std::fabs(s - y[i]));
            den = std::fmax(den, std::fabs(s));
        }
        if (num / den > 1.0e-15 * 4) { std::printf(" %s matvec%s failed: %g\n", name, trans ? "_t" : "", num / den); return 1; }
    }
    // y = y + A x
    std::vector<dp> y0(nn, 0.5);
    A.matvec(x.data(), z.data());
    A.matvec_add(x.data(), y0.data());
    for (int i = 0; i < nn; i++)
        if (std::fabs(y0[i] - (0.5 + z[i])) > 1e-15 * (1 + std::fabs(z[i]))) { std::printf(" %s matvec_add failed\n", name); return 1; }
    // host mutation -> the mirror is refreshed
    A.scalar_multiply(2.0);
    A.matvec(x.data(), y.data());
    for (int i = 0; i < nn; i++)
        if (y[i] != 2.0 * z[i]) { std::printf(" %s stale mirror after scalar_multiply\n", name); return 1; }
    return 0;
}

// a batch of add_value calls on the device against the same calls issued one by one on the
// host (csr_matrix_add_value cs_matrices.f90:868-891 and friends): bit-identical values
template <class M>
static int check_assembly(const char *name, const ll_graph &g, int nn, rng64 rnd)
{
    M A, B;
    A.init(nn, nn); A.copy_graph(g);
    B.init(nn, nn); B.copy_graph(g);
    std::vector<int32_t> is, js;
    std::vector<dp> zs;
    for (int pass = 0; pass < 5; pass++)
        for (int i = 1; i <= nn; i++)
            for (int32_t j : g.get_neighbors(i)) {
                is.push_back(i);
                js.push_back(j);
                zs.push_back((2 * rnd.next() - 1) * std::pow(10.0, (int)(rnd.next() * 12) - 6));
            }
    for (size_t c = 0; c < is.size(); c++) B.add_value(is[c], js[c], zs[c]);
    A.add_values(is, js, zs);
    for (size_t k = 0; k < A.val.size(); k++)
        if (A.val[k] != B.val[k]) { std::printf(" %s add_values differs from the add_value loop at %zu\n", name, k); return 1; }
    std::vector<dp> x(nn, 1.0), y(nn), z(nn);
    A.matvec(x.data(), y.data());
    B.matvec(x.data(), z.data());
    for (int i = 0; i < nn; i++)
        if (y[i] != z[i]) { std::printf(" %s matvec after add_values failed\n", name); return 1; }
    return 0;
}

int main()
{
    const int nn = 64;
    rng64 rnd(5);
    ll_graph g;
    g.init(nn);
    std::vector<dp> B((size_t)nn * nn, 0.0);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);
        for (int j = 1; j <= nn; j++)
            if (j != i && rnd.next() < 0.1) g.add_edge(i, j);
    }
    csr_matrix A1;
    csc_matrix A2;
    ellpack_matrix A3;
    A1.init(nn, nn); A1.copy_graph(g);
    A2.init(nn, nn); A2.copy_graph(g);
    A3.init(nn, nn); A3.copy_graph(g);
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i)) {
            const dp z = 2 * rnd.next() - 1;
            B[(size_t)(i - 1) * nn + (j - 1)] = z;
            A1.set_value(i, j, z);
            A2.set_value(i, j, z);
            A3.set_value(i, j, z);
        }
    if (A1.get_value(1, 1) != B[0] || A2.get_value(1, 1) != B[0] || A3.get_value(1, 1) != B[0]) { std::printf(" get_value failed\n"); return 1; }
    int rc = check("csr", A1, B, nn, rnd) | check("csc", A2, B, nn, rnd) | check("ellpack", A3, B, nn, rnd);
    rc |= check_assembly<csr_matrix>("csr", g, nn, rnd) | check_assembly<csc_matrix>("csc", g, nn, rnd) |
          check_assembly<ellpack_matrix>("ellpack", g, nn, rnd);
    // set_solver / set_preconditioner / A%solve facade (linear_operator_interface.f90:213-280)
    csr_matrix S;
    S.init(nn, nn);
    ll_graph t;
    t.init(nn);
    for (int i = 1; i < nn; i++) { t.add_edge(i, i); t.add_edge(i, i + 1); t.add_edge(i + 1, i); }
    t.add_edge(nn, nn);
    S.copy_graph(t);
    for (int i = 1; i < nn; i++) { S.set_value(i, i, 2.5); S.set_value(i, i + 1, -1.0); S.set_value(i + 1, i, -1.0); }
    S.set_value(nn, nn, 2.5);
    linear_solver *s = cg(1e-13), *pc = jacobi();
    S.set_solver(s);
    S.set_preconditioner(pc);
    std::vector<dp> v(nn), f(nn), u(nn, 0.0);
    for (dp &e : v) e = rnd.next();
    S.matvec(v.data(), f.data());
    S.solve(u.data(), f.data());
    for (int i = 0; i < nn; i++)
        if (std::fabs(u[i] - v[i]) > 1e-11) { std::printf(" A%%solve facade failed\n"); rc = 1; break; }
    // non-square operators are refused at setup with the reference's message
    delete s;
    delete pc;
    return rc;
}
