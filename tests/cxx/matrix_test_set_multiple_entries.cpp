// test/matrix_test_set_multiple_entries.f90 restated against sigma.hpp: the Laplacian of an
// Erdos-Renyi graph assembled from 2 x 2 element blocks B = [1 -1; -1 1] with
// A%add_multiple_values([i, j], [i, j], B) for every edge j > i (:92-114), for each device
// format; then A must have degree - 1 on the diagonal, -1 on every edge and nothing else
// (:120-152, exact comparisons).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

template <class M>
static int run(const ll_graph &g, int nn, const char *name, bool verbose)
{
    M A;
    A.init(nn, nn);
    A.copy_graph(g);
    A.zero();
    const std::vector<dp> B = {1.0, -1.0, -1.0, 1.0};
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i))
            if (j > i) A.add_multiple_values({i, j}, {i, j}, B);
    for (int i = 1; i <= nn; i++) {
        const dp d = g.get_degree(i) - 1;
        for (int j = 1; j <= nn; j++) {
            const dp z = A.get_value(i, j);
            if (j == i) {
                if (z != d) { std::printf(" %s: diagonal entry %d should equal its degree %g, found %g\n", name, i, d, z); return 1; }
            } else if (g.connected(i, j)) {
                if (z != -1.0) { std::printf(" %s: off-diagonal entry (%d,%d) should be -1, found %g\n", name, i, j, z); return 1; }
            } else if (z != 0.0) {
                std::printf(" %s: erroneously set entry (%d,%d) = %g\n", name, i, j, z);
                return 1;
            }
        }
    }
    if (verbose) std::printf(" o %s: setting multiple matrix entries works\n", name);
    return 0;
}

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && (!std::strcmp(argv[1], "-v") || !std::strcmp(argv[1], "-V") || !std::strcmp(argv[1], "--verbose"));
    rng64 rnd(31415);
    const int nn = 64;
    const dp p = std::log(1.0 * nn) / std::log(2.0) / nn;
    ll_graph g;
    g.init(nn);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);
        for (int j = i + 1; j <= nn; j++)
            if (rnd.next() < p) { g.add_edge(i, j); g.add_edge(j, i); }
    }
    if (verbose) std::printf(" o Done generating random graph: %d vertices, %d edges\n", nn, g.get_num_edges());
    if (run<csr_matrix>(g, nn, "csr", verbose)) return 1;
    if (run<csc_matrix>(g, nn, "csc", verbose)) return 1;
    if (run<ellpack_matrix>(g, nn, "ellpack", verbose)) return 1;
    return 0;
}
