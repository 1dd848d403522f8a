// test/eigensolver_test_lanczos.f90 restated: Erdos-Renyi graph Laplacian in
// CSR (nn = 128), nq = 11 Lanczos steps; three-term recurrence and
// orthogonality of the Lanczos vectors to 1e-14 (:130-170).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && !strcmp(argv[1], "-v");
    const int nn = 128;
    const dp p = std::log(1.0 * nn) / std::log(2.0) / nn;
    rng64 rnd(77);

    ll_graph g;
    g.init(nn);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);
        for (int j = i + 1; j <= nn; j++)
            if (rnd.next() < p) { g.add_edge(i, j); g.add_edge(j, i); }
    }
    csr_matrix A;
    A.init(nn, nn);
    A.copy_graph(g);
    A.zero();
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i)) {
            A.add_value(i, j, -1.0);
            A.add_value(i, i, +1.0);
        }

    const int nq = (int)std::sqrt(1.0 * nn);
    std::vector<dp> T(3 * nq), V((size_t)nn * nq), x(nn), y(nn);
    for (int i = 0; i < nn; i++) V[i] = 2 * rnd.next() - 1;     // random_number(Q(:,1)); Q = 2Q - 1
    lanczos(A, nq, T.data(), V.data(), /*use_q1=*/true);

    auto Tm = [&](int row, int col) { return T[3 * (size_t)(col - 1) + (row - 1)]; };
    auto Vc = [&](int col) { return V.data() + (size_t)(col - 1) * nn; };
    for (int i = 2; i <= nq - 1; i++) {
        A.matvec(Vc(i), x.data());
        dp num = 0, den = 0;
        for (int l = 0; l < nn; l++) {
            y[l] = Tm(2, i) * Vc(i)[l] + Tm(1, i - 1) * Vc(i - 1)[l] + Tm(3, i) * Vc(i + 1)[l];
            num += (y[l] - x[l]) * (y[l] - x[l]);
            den += x[l] * x[l];
        }
        const dp err = std::sqrt(num / den);
        if (err > 1.0e-14) { std::printf(" Computing Lanczos vector failed! three-term recurrence error %g\n", err); return 1; }
    }
    dp fro = 0;
    for (int a = 1; a <= nq; a++)
        for (int b = 1; b <= nq; b++) {
            dp s = 0;
            for (int l = 0; l < nn; l++) s += Vc(a)[l] * Vc(b)[l];
            if (a == b) s -= 1.0;
            fro += s * s;
        }
    const dp err = std::sqrt(fro) / nq;
    if (err > 1.0e-14) { std::printf(" Lanczos vectors are not orthogonal! %g\n", err); return 1; }
    if (verbose) std::printf(" o Lanczos identities hold, orthogonality %g\n", err);
    return 0;
}
