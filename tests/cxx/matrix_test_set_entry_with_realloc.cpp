// test/matrix_test_set_entry_with_realloc.f90 restated against sigma.hpp: a ring graph with self
// edges (nn = 64, :52-60), its entries set (:73-79), then entries that were NOT pre-allocated in the
// graph set / added (:83-87), read back (:91-99), and a 2 x 2 block added across unallocated
// entries with add_multiple_values (:103-113).  Every call here is host-side index / value work
// (the reference's reallocation path is host code too), so this program also runs without a GPU;
// with one, the final matvec checks that the regrown matrix reaches the device intact.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
using namespace sigma;

template <class M>
static int run(const ll_graph &g, int nn, const char *name, bool verbose, bool on_device)
{
    M A;
    A.init(nn, nn);
    A.copy_graph(g);
    for (int i = 1; i <= nn; i++) {
        A.set_value(i, i, +2.0);
        const int j = i % nn + 1;
        A.set_value(i, j, -1.0);
        A.set_value(j, i, -1.0);
    }
    for (int i = 1; i <= nn; i++) {              // entries that have not been pre-allocated
        const int j = (i + 1) % nn + 1;
        A.set_value(i, j, -1.0);
        A.add_value(i, i, +1.0);
    }
    for (int i = 1; i <= nn; i++) {
        const int j = (i + 1) % nn + 1;
        if (A.get_value(i, j) != -1.0) { std::printf(" %s: Matrix entry (%d,%d) not set.\n", name, i, j); return 1; }
        if (A.get_value(i, i) != 3.0 || A.get_value(i, i % nn + 1) != -1.0 || A.get_value(i % nn + 1, i) != -1.0) {
            std::printf(" %s: an entry set before the reallocation was lost in row %d\n", name, i);
            return 1;
        }
    }
    A.add_multiple_values({1, nn / 2}, {1, nn / 2}, {1.0, -1.0, -1.0, 1.0});
    if (A.get_value(1, 1) != 4.0 || A.get_value(1, nn / 2) != -1.0 || A.get_value(nn / 2, 1) != -1.0 ||
        A.get_value(nn / 2, nn / 2) != 4.0) {
        std::printf(" %s: Setting multiple matrix entries with reallocation failed.\n", name);
        return 1;
    }
    if (on_device) {
        // y = A 1 by hand: every row sums its entries
        std::vector<dp> x((size_t)nn, 1.0), y((size_t)nn, 0.0);
        A.matvec(x.data(), y.data());
        for (int i = 1; i <= nn; i++) {
            dp want = 0.0;
            for (int j = 1; j <= nn; j++) want += A.get_value(i, j);
            if (y[(size_t)i - 1] != want) { std::printf(" %s: matvec after reallocation, row %d: %g vs %g\n", name, i, y[(size_t)i - 1], want); return 1; }
        }
    }
    if (verbose) std::printf(" o %s: setting unallocated entries works%s\n", name, on_device ? " (device matvec checked)" : "");
    return 0;
}

int main(int argc, char **argv)
{
    bool verbose = false, on_device = true;
    for (int a = 1; a < argc; a++) {
        if (!std::strcmp(argv[a], "-v") || !std::strcmp(argv[a], "-V") || !std::strcmp(argv[a], "--verbose")) verbose = true;
        if (!std::strcmp(argv[a], "--host-only")) on_device = false;
    }
    const int nn = 64;
    ll_graph g;
    g.init(nn, nn);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);
        const int j = i % nn + 1;
        g.add_edge(i, j);
        g.add_edge(j, i);
    }
    if (run<csr_matrix>(g, nn, "csr", verbose, on_device)) return 1;
    if (run<csc_matrix>(g, nn, "csc", verbose, on_device)) return 1;
    if (run<ellpack_matrix>(g, nn, "ellpack", verbose, on_device)) return 1;
    return 0;
}
