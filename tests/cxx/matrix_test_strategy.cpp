// test/matrix_test_strategy.f90 restated against sigma.hpp: type(sparse_matrix) as the
// container of ONE storage strategy (:98-101 set_matrix_type("csr") -- here all three device
// formats in turn), an Erdos-Renyi graph Laplacian assembled with add_value (:109-117), entries
// against the graph (:127-153, exact) and A%matvec against the Laplacian written out from the
// neighbour lists (:225-254, relative error 1e-14).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

template <class M>
static int run(const ll_graph &g, int nn, const std::vector<dp> &x, const char *name, bool verbose, bool on_device)
{
    auto *L = new M();
    L->init(nn, nn);
    L->copy_graph(g);
    L->zero();
    sparse_matrix A;                                 // the strategy container: a 1 x 1 composite
    A.set_dimensions(nn, nn);
    A.set_block_sizes({nn}, {nn});
    A.set_submatrix(1, 1, *L);
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i)) {
            A.add(1, 1, i, j, -1.0);
            A.add(1, 1, i, i, +1.0);
        }
    for (int i = 1; i <= nn; i++) {
        const dp d = g.get_degree(i) - 1;
        if (A.get_value(i, i) != d) { std::printf(" %s: A(%d,%d) = %g, degree %g\n", name, i, i, A.get_value(i, i), d); return 1; }
        for (int j = i + 1; j <= nn; j++) {
            const dp z = A.get_value(i, j), want = g.connected(i, j) ? -1.0 : 0.0;
            if (z != want) { std::printf(" %s: Setting or getting matrix entry (%d,%d) failed: %g\n", name, i, j, z); return 1; }
        }
    }
    // rows and columns (:158-219): every stored neighbour carries -1, the vertex itself degree - 1
    std::vector<int32_t> nodes;
    std::vector<dp> slice;
    for (int dir = 0; dir < 2; dir++)
        for (int i = 1; i <= nn; i++) {
            if (dir == 0) L->get_row(nodes, slice, i); else L->get_column(nodes, slice, i);
            if ((int)nodes.size() != g.get_degree(i)) { std::printf(" %s: Getting matrix %s %d failed: %zu entries\n", name, dir ? "column" : "row", i, nodes.size()); return 1; }
            for (size_t k = 0; k < nodes.size(); k++) {
                const dp want = nodes[k] == i ? g.get_degree(i) - 1.0 : -1.0;
                if (slice[k] != want || !g.connected(i, nodes[k])) { std::printf(" %s: Getting matrix %s failed at (%d,%d)\n", name, dir ? "column" : "row", i, nodes[k]); return 1; }
            }
        }
    if (!on_device) {
        if (verbose) std::printf(" o %s: entries, rows and columns work (host only)\n", name);
        A.destroy();
        return 0;
    }
    std::vector<dp> y(nn), w(nn, 0.0);
    for (int i = 1; i <= nn; i++) {
        dp z = g.get_degree(i) * x[(size_t)i - 1];
        for (int32_t j : g.get_neighbors(i)) z -= x[(size_t)j - 1];
        y[(size_t)i - 1] = z;
    }
    A.matvec(x.data(), w.data());
    dp num = 0, den = 0;
    for (int i = 0; i < nn; i++) { num += (y[i] - w[i]) * (y[i] - w[i]); den += x[i] * x[i]; }
    const dp err = std::sqrt(num / den);
    if (err > 1.0e-14) { std::printf(" %s: Matrix-vector multiplication failed: %g\n", name, err); return 1; }
    if (verbose) std::printf(" o %s: entries and matrix-vector product work (%g)\n", name, err);
    A.destroy();
    return 0;
}

int main(int argc, char **argv)
{
    bool verbose = false, on_device = true;
    for (int a = 1; a < argc; a++) {
        if (!std::strcmp(argv[a], "-v") || !std::strcmp(argv[a], "-V") || !std::strcmp(argv[a], "--verbose")) verbose = true;
        if (!std::strcmp(argv[a], "--host-only")) on_device = false;
    }
    rng64 rnd(2718);
    const int nn = 256;
    const dp c = std::log(1.0 * nn) / std::log(2.0) / nn;
    ll_graph g;
    g.init(nn);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);
        for (int j = i + 1; j <= nn; j++)
            if (rnd.next() < c) { g.add_edge(i, j); g.add_edge(j, i); }
    }
    if (verbose) std::printf(" o Done generating Erdos-Renyi graph: %d vertices, %d edges\n", nn, g.get_num_edges());
    std::vector<dp> x(nn);
    for (dp &v : x) v = rnd.next();
    if (run<csr_matrix>(g, nn, x, "csr", verbose, on_device)) return 1;
    if (run<csc_matrix>(g, nn, x, "csc", verbose, on_device)) return 1;
    if (run<ellpack_matrix>(g, nn, x, "ellpack", verbose, on_device)) return 1;
    return 0;
}
