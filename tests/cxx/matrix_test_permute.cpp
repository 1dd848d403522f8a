// The permutation part of test/matrix_test_basics.f90 (:139-158, :364-392) restated against
// sigma.hpp: a random matrix B on a random graph, a random permutation p (Fisher-Yates as in the
// test), then
//     BP(:, p) = B ;  call A%right_permute(p)  ->  A%get_value(i, j) == BP(i, j) for all i, j
//     BP(p, :) = BP;  call A%left_permute(p)   ->  likewise
// for the three device formats.  Permutations are host-side index work in the reference and in the
// mirror (the device mirrors are dropped and rebuilt by the next matvec), so the program runs
// without a GPU with --host-only; with one, the permuted matrix's matvec is checked against BP x.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

template <class M>
static int run(const ll_graph &g, const std::vector<dp> &B, const std::vector<int> &p, int nn, const char *name,
               bool verbose, bool on_device, rng64 &rnd)
{
    M A;
    A.init(nn, nn);
    A.copy_graph(g);
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i)) A.set_value(i, j, B[(size_t)(i - 1) * nn + (j - 1)]);
    std::vector<dp> BP((size_t)nn * nn, 0.0), T((size_t)nn * nn, 0.0);
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn; j++) BP[(size_t)(i - 1) * nn + (p[(size_t)j - 1] - 1)] = B[(size_t)(i - 1) * nn + (j - 1)];   // BP(:, p) = B
    A.right_permute(p);
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn; j++)
            if (A.get_value(i, j) != BP[(size_t)(i - 1) * nn + (j - 1)]) { std::printf(" %s: Right-permutation failed at (%d,%d).\n", name, i, j); return 1; }
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn; j++) T[(size_t)(p[(size_t)i - 1] - 1) * nn + (j - 1)] = BP[(size_t)(i - 1) * nn + (j - 1)];   // BP(p, :) = BP
    BP.swap(T);
    A.left_permute(p);
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn; j++)
            if (A.get_value(i, j) != BP[(size_t)(i - 1) * nn + (j - 1)]) { std::printf(" %s: Left-permutation failed at (%d,%d).\n", name, i, j); return 1; }
    if (on_device) {
        std::vector<dp> x((size_t)nn), y((size_t)nn, 0.0);
        for (dp &v : x) v = rnd.next();
        A.matvec(x.data(), y.data());
        dp err = 0, scale = 0;
        for (int i = 0; i < nn; i++) {
            dp z = 0;
            for (int j = 0; j < nn; j++) z += BP[(size_t)i * nn + j] * x[(size_t)j];
            err = std::fmax(err, std::fabs(z - y[(size_t)i]));
            scale = std::fmax(scale, std::fabs(z));
        }
        if (err > 1.0e-15 * scale * nn) { std::printf(" %s: matvec of the permuted matrix failed: %g\n", name, err / scale); return 1; }
    }
    if (verbose) std::printf(" o %s: right and left permutation work%s\n", name, on_device ? " (device matvec checked)" : "");
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
    rng64 rnd(1234);
    ll_graph g;
    g.init(nn, nn);
    std::vector<dp> B((size_t)nn * nn, 0.0);
    for (int i = 1; i <= nn; i++) {
        g.add_edge(i, i);                                  // (keeps every ellpack row non-empty)
        for (int j = 1; j <= nn; j++)
            if (j != i && rnd.next() < 0.1) g.add_edge(i, j);
    }
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i)) B[(size_t)(i - 1) * nn + (j - 1)] = 2 * rnd.next() - 1;
    std::vector<int> p((size_t)nn);
    for (int i = 1; i <= nn; i++) p[(size_t)i - 1] = i;
    for (int i = nn; i >= 2; i--) {                        // :144-153
        const int j = (int)(rnd.next() * i) + 1;
        std::swap(p[(size_t)i - 1], p[(size_t)j - 1]);
    }
    if (run<csr_matrix>(g, B, p, nn, "csr", verbose, on_device, rnd)) return 1;
    if (run<csc_matrix>(g, B, p, nn, "csc", verbose, on_device, rnd)) return 1;
    if (run<ellpack_matrix>(g, B, p, nn, "ellpack", verbose, on_device, rnd)) return 1;
    return 0;
}
