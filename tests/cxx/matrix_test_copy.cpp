// test/matrix_test_copy.f90:60-146 restated against sigma.hpp: a random nn x nn/2
// matrix in each device format is copied -- straight and transposed -- into each
// device format with A%copy_matrix (built on the device here); every entry must
// agree to 1e-14 (:110-115, :130-135).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../sigma_b200/host/sigma.hpp"
#include "test_util.hpp"
using namespace sigma;

template <class M>
static std::unique_ptr<device_matrix> make_source(const ll_graph &g, int nn, rng64 rnd)
{
    auto A = std::make_unique<M>();
    A->init(nn, nn / 2);
    A->copy_graph(g);
    for (int i = 1; i <= nn; i++)
        for (int32_t j : g.get_neighbors(i)) A->set_value(i, j, rnd.next());
    return A;
}

template <class M>
static int check_copies(const char *from, const char *to, device_matrix &A, int nn)
{
    M B;
    B.init(nn, nn / 2);
    B.copy_matrix(A, false);
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn / 2; j++)
            if (std::fabs(A.get_value(i, j) - B.get_value(i, j)) > 1.0e-14) {
                std::printf(" Copying matrix %s -> %s failed %d %d\n", from, to, i, j);
                return 1;
            }
    M C;
    C.init(nn / 2, nn);
    C.copy_matrix(A, true);
    for (int i = 1; i <= nn; i++)
        for (int j = 1; j <= nn / 2; j++)
            if (std::fabs(A.get_value(i, j) - C.get_value(j, i)) > 1.0e-14) {
                std::printf(" Copying transpose matrix %s -> %s failed %d %d\n", from, to, i, j);
                return 1;
            }
    // the copies keep working as operators, and host mutation still reaches the device
    std::vector<dp> x(nn / 2, 1.0), y(nn), z(nn);
    A.matvec(x.data(), y.data());
    B.matvec(x.data(), z.data());
    for (int i = 0; i < nn; i++)
        if (std::fabs(y[i] - z[i]) > 1.0e-13) { std::printf(" matvec of the copy %s -> %s failed\n", from, to); return 1; }
    C.matvec_t(x.data(), z.data());
    for (int i = 0; i < nn; i++)
        if (std::fabs(y[i] - z[i]) > 1.0e-13) { std::printf(" matvec_t of the transposed copy %s -> %s failed\n", from, to); return 1; }
    B.scalar_multiply(3.0);
    B.matvec(x.data(), z.data());
    for (int i = 0; i < nn; i++)
        if (std::fabs(3.0 * y[i] - z[i]) > 1.0e-13) { std::printf(" stale mirror after mutating the copy %s -> %s\n", from, to); return 1; }
    return 0;
}

int main(int argc, char **argv)
{
    const bool verbose = argc > 1 && (!std::strcmp(argv[1], "-v") || !std::strcmp(argv[1], "-V") || !std::strcmp(argv[1], "--verbose"));
    const int nn = 64;
    const dp probability = std::log(1.0 * nn) / std::log(2.0) / nn * 2;
    rng64 rnd(31);
    ll_graph g;
    g.init(nn, nn / 2);
    for (int i = 1; i <= nn; i++) {
        for (int j = 1; j <= nn / 2; j++)
            if (rnd.next() < probability) g.add_edge(i, j);
        // the device ellpack format refuses rows without an edge (the reference would
        // read x(0), README.md:71-73); keep every row and column populated
        if (g.get_degree(i) == 0) g.add_edge(i, 1 + (i % (nn / 2)));
    }
    for (int j = 1; j <= nn / 2; j++) g.add_edge(j, j);
    if (verbose) std::printf(" %d edges\n", g.get_num_edges());

    const char *names[3] = {"csr", "csc", "ellpack"};
    for (int f1 = 0; f1 < 3; f1++) {
        std::unique_ptr<device_matrix> A = f1 == 0   ? make_source<csr_matrix>(g, nn, rnd)
                                           : f1 == 1 ? make_source<csc_matrix>(g, nn, rnd)
                                                     : make_source<ellpack_matrix>(g, nn, rnd);
        if (verbose) std::printf(" from %s\n", names[f1]);
        if (check_copies<csr_matrix>(names[f1], "csr", *A, nn)) return 1;
        if (check_copies<csc_matrix>(names[f1], "csc", *A, nn)) return 1;
        if (check_copies<ellpack_matrix>(names[f1], "ellpack", *A, nn)) return 1;
    }
    return 0;
}
