// test/matrix_test_composite.f90 restated against sigma.hpp: a 2 x 2 composite
// sparse_matrix (nn1 = 768, nn2 = 512) of two random weighted Laplacians (csr)
// coupled through ONE random graph h used as the csr (1,2) block and as the csc
// (2,1) block (:171-186); entries through A%get (:229-285) and A%matvec against
// the product written out from the graphs, RMS bar 1e-14 (:413-487).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

// erdos_renyi_graph (:560-590)
static void erdos_renyi_graph(ll_graph &g, int m, int n, dp p, bool symmetric, rng64 &rnd)
{
    g.init(m, n);
    for (int i = 1; i <= m; i++) {
        if (symmetric) g.add_edge(i, i);
        for (int j = i + 1; j <= n; j++)
            if (rnd.next() < p) {
                g.add_edge(i, j);
                if (symmetric) g.add_edge(j, i);
            }
    }
}

// erdos_renyi_matrix (:595-620)
static csr_matrix *erdos_renyi_matrix(const ll_graph &g)
{
    auto *A = new csr_matrix();
    A->init(g.n, g.m);
    A->copy_graph(g);
    for (int i = 1; i <= g.n; i++)
        for (int32_t j : g.get_neighbors(i)) {
            A->add_value(i, j, -1.0);
            A->add_value(i, i, +1.0);
        }
    return A;
}

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && (!std::strcmp(argv[1], "-v") || !std::strcmp(argv[1], "-V") || !std::strcmp(argv[1], "--verbose"));
    rng64 rnd(77);
    const int nn1 = 768, nn2 = 512, nn = nn1 + nn2;
    std::vector<dp> x(nn), y(nn, 0.0), z(nn, 0.0);
    for (dp &v : x) v = rnd.next();

    sparse_matrix A;
    A.set_dimensions(nn, nn);
    A.set_block_sizes({nn1, nn2}, {nn1, nn2});
    if (A.num_row_mats != 2 || A.num_col_mats != 2) { std::printf(" Setting number of blocks failed\n"); return 1; }

    ll_graph g1, g2, h;
    erdos_renyi_graph(g1, nn1, nn1, std::log(1.0 * nn1) / std::log(2.0) / nn1, true, rnd);
    A.set_submatrix(1, 1, *erdos_renyi_matrix(g1));
    erdos_renyi_graph(g2, nn2, nn2, std::log(1.0 * nn2) / std::log(2.0) / nn2, true, rnd);
    A.set_submatrix(2, 2, *erdos_renyi_matrix(g2));
    if (verbose) std::printf(" o Done generating the random weighted Laplacians: %d and %d edges\n", g1.get_num_edges(), g2.get_num_edges());

    // the coupling graph, converted to CS storage once and shared by both blocks
    erdos_renyi_graph(h, nn1, nn2, 6.0 / nn1, false, rnd);
    auto hcs = std::make_shared<cs_graph>();
    hcs->copy(h);
    auto *C12 = new csr_matrix();
    C12->init(nn1, nn2);
    C12->set_graph(hcs);
    A.set_submatrix(1, 2, *C12);
    auto *C21 = new csc_matrix();
    C21->init(nn2, nn1);
    C21->set_graph(hcs);
    A.set_submatrix(2, 1, *C21);
    if (verbose) std::printf(" o Done creating couplings via another random graph h (%d edges), references to h: %ld\n",
                             h.get_num_edges(), hcs.use_count());

    for (int i = 1; i <= nn1; i++)
        for (int32_t j : h.get_neighbors(i)) {
            A.set(1, 2, i, j, -1.0);
            A.set(2, 1, j, i, -1.0);
            A.add(1, 1, i, i, +1.0);
            A.add(2, 2, j, j, +1.0);
        }

    // entries (:229-285)
    std::vector<int> htdeg(nn2, 0);
    for (int i = 1; i <= nn1; i++)
        for (int32_t j : h.get_neighbors(i)) htdeg[(size_t)j - 1]++;
    for (int i = 1; i <= nn1; i++) {
        for (int j = 1; j <= nn1; j++) {
            dp correct = 0.0;
            if (j == i) correct = g1.get_degree(i) + h.get_degree(i) - 1.0;
            else if (g1.connected(i, j)) correct = -1.0;
            if (std::fabs(A.get(1, 1, i, j) - correct) > 1.0e-15) { std::printf(" entry (%d,%d) of sub-matrix (1,1) failed\n", i, j); return 1; }
        }
        for (int j = 1; j <= nn2; j++) {
            const dp correct = h.connected(i, j) ? -1.0 : 0.0;
            if (std::fabs(A.get(1, 2, i, j) - correct) > 1.0e-15) { std::printf(" entry (%d,%d) of sub-matrix (1,2) failed\n", i, j); return 1; }
            if (std::fabs(A.get(2, 1, j, i) - correct) > 1.0e-15) { std::printf(" entry (%d,%d) of sub-matrix (2,1) failed\n", j, i); return 1; }
        }
    }
    for (int i = 1; i <= nn2; i++)
        for (int j = 1; j <= nn2; j++) {
            dp correct = 0.0;
            if (j == i) correct = g2.get_degree(i) + htdeg[(size_t)i - 1] - 1.0;
            else if (g2.connected(i, j)) correct = -1.0;
            if (std::fabs(A.get(2, 2, i, j) - correct) > 1.0e-15) { std::printf(" entry (%d,%d) of sub-matrix (2,2) failed\n", i, j); return 1; }
        }
    // global indexing goes through the owning blocks (:465-485)
    if (A.get_value(nn1 + 3, nn1 + 3) != A.get(2, 2, 3, 3) || A.get_value(5, nn1 + 9) != A.get(1, 2, 5, 9)) {
        std::printf(" composite get_value failed\n");
        return 1;
    }
    if (verbose) std::printf(" o Done checking entries of A.\n");

    // matvec (:413-487)
    A.matvec(x.data(), y.data());
    for (int i = 1; i <= nn1; i++)
        for (int32_t j : g1.get_neighbors(i)) z[(size_t)i - 1] += x[(size_t)i - 1] - x[(size_t)j - 1];
    for (int i = 1; i <= nn2; i++)
        for (int32_t j : g2.get_neighbors(i)) z[(size_t)i + nn1 - 1] += x[(size_t)i + nn1 - 1] - x[(size_t)j + nn1 - 1];
    for (int i = 1; i <= nn1; i++)
        for (int32_t j : h.get_neighbors(i)) z[(size_t)i - 1] += x[(size_t)i - 1] - x[(size_t)j + nn1 - 1];
    for (int j = 1; j <= nn1; j++)
        for (int32_t i : h.get_neighbors(j)) z[(size_t)i + nn1 - 1] += x[(size_t)i + nn1 - 1] - x[(size_t)j - 1];
    dp num = 0, den = 0;
    for (int i = 0; i < nn; i++) { num += (y[i] - z[i]) * (y[i] - z[i]); den += x[i] * x[i]; }
    const dp mse = std::sqrt(num / den);
    if (mse > 1.0e-14) { std::printf(" Matrix-vector multiplication failed: %g\n", mse); return 1; }
    // symmetric by construction: the block-column loop of matvec_t gives the same vector
    A.matvec_t(x.data(), y.data());
    num = 0;
    for (int i = 0; i < nn; i++) num += (y[i] - z[i]) * (y[i] - z[i]);
    if (std::sqrt(num / den) > 1.0e-14) { std::printf(" Transposed matrix-vector multiplication failed\n"); return 1; }
    if (verbose) std::printf(" o Done checking matrix-vector multiplication: %g\n", mse);

    A.destroy();   // drops the blocks (their only reference was the composite's)
    return 0;
}
