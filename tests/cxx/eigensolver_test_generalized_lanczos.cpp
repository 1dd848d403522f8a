// test/eigensolver_test_generalized_lanczos.f90 restated: stiffness A and mass
// B of a periodic 48 x 32 grid of right triangles, B%set_solver(cg(1.0d-15)),
// nq = max(nx, ny) generalized Lanczos steps; three-term recurrence and
// B-orthogonality to 1e-14.  (The reference program only prints on failure;
// here a failure is an exit code.)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

static const int nx = 48, ny = 32;
static int indx(int i, int j) { return ny * (j - 1) + i; }

template <class M>
static void add(M &A, const int *elem, const dp (*E)[3])   // A%add(elem, elem, E)
{
    for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) A.add_value(elem[k], elem[l], E[k][l]);
}

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && !strcmp(argv[1], "-v");
    const int nn = nx * ny;
    ll_graph g;
    g.init(nn);
    for (int i = 1; i <= ny; i++)
        for (int j = 1; j <= nx; j++) {
            const int k = indx(i, j);
            g.add_edge(k, k);
            const int ls[3] = {indx(i % ny + 1, j), indx(i, j % nx + 1), indx(i % ny + 1, j % nx + 1)};
            for (int l : ls) { g.add_edge(k, l); g.add_edge(l, k); }
        }

    csr_matrix A, B;
    A.init(nn, nn); B.init(nn, nn);
    A.copy_graph(g); B.copy_graph(g);
    A.zero(); B.zero();
    const dp area = 0.5;
    dp BE[3][3], AE[3][3] = {{+area, -area, 0.0}, {-area, 2 * area, -area}, {0.0, -area, +area}};
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) BE[a][b] = (a == b) ? area / 6.0 : area / 12.0;
    for (int i = 1; i <= ny; i++)
        for (int j = 1; j <= nx; j++) {
            int elem[3] = {indx(i, j), indx(i, j % nx + 1), indx(i % ny + 1, j % nx + 1)};
            add(A, elem, AE); add(B, elem, BE);
            elem[1] = indx(i % ny + 1, j);
            add(A, elem, AE); add(B, elem, BE);
        }

    const int nq = nx > ny ? nx : ny;
    std::vector<dp> T(3 * nq), V((size_t)nn * nq), U((size_t)nn * nq), w(nn), z(nn);
    linear_solver *bs = cg(1.0e-15);
    B.set_solver(bs);
    bs->set_max_iterations(5000);
    rng64 rnd(3);
    for (int i = 0; i < nn; i++) V[i] = 2 * rnd.next() - 1;
    generalized_lanczos(A, B, nq, T.data(), V.data(), /*use_q1=*/true);
    if (bs->capped()) { std::printf(" inner CG hit the safety cap\n"); return 1; }

    auto Tm = [&](int row, int col) { return T[3 * (size_t)(col - 1) + (row - 1)]; };
    auto Vc = [&](int col) { return V.data() + (size_t)(col - 1) * nn; };
    auto Uc = [&](int col) { return U.data() + (size_t)(col - 1) * nn; };
    for (int i = 1; i <= nq; i++) B.matvec(Vc(i), Uc(i));
    int rc = 0;
    for (int i = 2; i <= nq - 1; i++) {
        A.matvec(Vc(i), w.data());
        dp num = 0, den = 0;
        for (int l = 0; l < nn; l++) {
            z[l] = Tm(2, i) * Uc(i)[l] + Tm(1, i - 1) * Uc(i - 1)[l] + Tm(3, i) * Uc(i + 1)[l];
            num += (w[l] - z[l]) * (w[l] - z[l]);
            den += w[l] * w[l];
        }
        if (std::sqrt(num / den) > 1.0e-14) { std::printf(" Computing Lanczos vector failed! %g\n", std::sqrt(num / den)); rc = 1; }
    }
    dp fro = 0;
    for (int a = 1; a <= nq; a++)
        for (int b = 1; b <= nq; b++) {
            dp s = 0;
            for (int l = 0; l < nn; l++) s += Vc(a)[l] * Uc(b)[l];
            if (a == b) s -= 1.0;
            fro += s * s;
        }
    const dp err = std::sqrt(fro) / nq;
    if (err > 1.0e-14) { std::printf(" Lanczos vectors are not B-orthogonal! %g\n", err); rc = 1; }
    if (verbose) std::printf(" o generalized Lanczos: B-orthogonality %g, inner CG iterations %d\n", err, bs->iterations);
    delete bs;
    return rc;
}
