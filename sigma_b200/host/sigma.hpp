// sigma.hpp -- C++ host side above the C-ABI (include/sigma_b200.h), mirroring
// the reference's Fortran interface for the hot path: same type and procedure
// names, argument meaning and error behaviour (print + exit(1)), so programs
// written against SiGMA read the same here (tests/cxx/*.cpp are the reference's
// own test programs restated).  The reference is Fortran 2003 and the image has
// no Fortran compiler, so this header plays the role of the modified Fortran
// modules; fortran/sigma_b200_shim.f90 shows the same calls in Fortran.
//
//   reference                                       here (namespace sigma)
//   ---------------------------------------------   -------------------------------
//   type(ll_graph)   g%init / g%add_edge            ll_graph
//   convert_graph_type(g, "compressed sparse")      cs_graph::copy(ll_graph [, trans])
//   convert_graph_type(g, "ellpack")                ellpack_graph::copy(ll_graph)
//   type(csr_matrix|csc_matrix|ellpack_matrix)      csr_matrix, csc_matrix, ellpack_matrix
//   A%set_graph, zero, set_value, add_value,        same names
//   get_value, scalar_multiply
//   A%matvec / matvec_t / matvec_add / matvec_t_add same names (linear_operator)
//   A%set_solver / set_preconditioner / solve       same names
//   cg(tol), bicgstab(tol), jacobi()                same names -> linear_solver*
//   solver%setup / solve(A,x,b[,pc]) / destroy      same names
//   lanczos(A,T,Q), eigensolve(A,lambda,V)          same names
//
// Host mutators mark the device mirror dirty; the next matvec / solve re-uploads
// the values.  Index arrays are 1-based int32 exactly as the
// Fortran holds them.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../include/sigma_b200.h"

namespace sigma {

using dp = double;

// the reference's error convention: print, then `call exit(1)`
inline void sigb_check(int stat)
{
    if (stat != SIGB_OK) {
        std::printf(" %s\n Terminating.\n", sigb_last_error());
        std::exit(1);
    }
}

// ---------------------------------------------------------------------------
// graphs
// ---------------------------------------------------------------------------

// src/graph/formats/ll_graphs.f90: list-of-lists builder, insertion order kept
struct ll_graph {
    int n = 0, m = 0, ne = 0, max_d = 0;
    std::vector<std::vector<int32_t>> lists;   // lists[i-1] = neighbours of i, 1-based ids

    void init(int n_, int m_ = -1)
    {
        n = n_;
        m = m_ < 0 ? n_ : m_;
        ne = max_d = 0;
        lists.assign((size_t)n, {});
    }
    bool connected(int i, int j) const
    {
        for (int32_t c : lists[(size_t)i - 1]) if (c == j) return true;
        return false;
    }
    void add_edge(int i, int j)   // ll_add_edge :355-371
    {
        if (connected(i, j)) return;
        auto &l = lists[(size_t)i - 1];
        l.push_back(j);
        if ((int)l.size() > max_d) max_d = (int)l.size();
        ne++;
    }
    int get_degree(int i) const { return (int)lists[(size_t)i - 1].size(); }
    const std::vector<int32_t> &get_neighbors(int i) const { return lists[(size_t)i - 1]; }
    int get_num_edges() const { return ne; }
    int get_max_degree() const { return max_d; }
};

// device mirror of a pattern, shared by every matrix on that graph and
// released with the last owner (the reference's manual reference counts,
// src/graph/graph_interfaces.f90:345-363)
struct graph_mirror {
    sigb_graph_t h = nullptr;
    ~graph_mirror() { if (h) sigb_graph_release(h); }
};

// src/graph/formats/cs_graphs.f90
struct cs_graph {
    int n = 0, m = 0, ne = 0, max_d = 0;
    std::vector<int32_t> ptr, node;            // ptr(n+1), node(ne), 1-based
    std::shared_ptr<graph_mirror> mirror;      // dropped by every mutator

    // g%copy(h, trans) -> cs_graph_build :109-197: count, 1-based prefix sum,
    // first-free-slot insertion in the source iteration order (never sorted)
    void copy(const ll_graph &h, bool trans = false)
    {
        n = trans ? h.m : h.n;
        m = trans ? h.n : h.m;
        ne = h.ne;
        ptr.assign((size_t)n + 1, 0);
        for (int i = 1; i <= h.n; i++)
            for (int32_t j : h.lists[(size_t)i - 1]) ptr[(size_t)(trans ? j : i)] += 1;
        ptr[0] = 1;
        for (int i = 1; i <= n; i++) ptr[(size_t)i] += ptr[(size_t)i - 1];
        node.assign((size_t)ne, 0);
        std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);   // next free slot per line
        for (int i = 1; i <= h.n; i++)
            for (int32_t j : h.lists[(size_t)i - 1]) {
                const int r = trans ? j : i, c = trans ? i : j;
                node[(size_t)fill[(size_t)r - 1]++ - 1] = c;
            }
        max_d = 0;
        for (int i = 0; i < n; i++) max_d = std::max(max_d, ptr[(size_t)i + 1] - ptr[(size_t)i]);
        mirror.reset();
    }
    int find_edge(int i, int j) const   // 0-based position in node, -1 if absent
    {
        for (int k = ptr[(size_t)i - 1]; k <= ptr[(size_t)i] - 1; k++)
            if (node[(size_t)k - 1] == j) return k - 1;
        return -1;
    }
};

// src/graph/formats/ellpack_graphs.f90: node(max_d, n), padding = last neighbour
struct ellpack_graph {
    int n = 0, m = 0, ne = 0, max_d = 0;
    std::vector<int32_t> node, degrees;        // node[(i-1)*max_d + (k-1)] == node(k, i)
    std::shared_ptr<graph_mirror> mirror;

    void copy(const ll_graph &h)   // ellpack_graph_build :105-170
    {
        n = h.n;
        m = h.m;
        ne = h.ne;
        max_d = h.max_d;
        node.assign((size_t)n * max_d, 0);
        degrees.assign((size_t)n, 0);
        for (int i = 1; i <= n; i++) {
            int32_t *row = node.data() + (size_t)(i - 1) * max_d;
            for (int32_t j : h.lists[(size_t)i - 1]) {
                const int d = degrees[(size_t)i - 1];
                for (int l = d; l < max_d; l++) row[l] = j;   // g%node(d+1:, i) = j  (:164)
                degrees[(size_t)i - 1] = d + 1;
            }
        }
        mirror.reset();
    }
};

// ---------------------------------------------------------------------------
// linear operators
// ---------------------------------------------------------------------------
struct linear_solver;

// src/linear_operator/linear_operator_interface.f90:18-45
struct linear_operator {
    int nrow = 0, ncol = 0;
    linear_solver *solver = nullptr, *pc = nullptr;
    virtual ~linear_operator() {}
    virtual void matvec_add(const dp *x, dp *y) = 0;
    virtual void matvec_t_add(const dp *x, dp *y) = 0;
    virtual void matvec(const dp *x, dp *y)      // :185-194: y = 0 ; matvec_add
    {
        for (int i = 0; i < nrow; i++) y[i] = 0.0;
        matvec_add(x, y);
    }
    virtual void matvec_t(const dp *x, dp *y)    // :199-208
    {
        for (int i = 0; i < ncol; i++) y[i] = 0.0;
        matvec_t_add(x, y);
    }
    virtual dp get_value(int, int) { return 0.0; }
    inline void set_solver(linear_solver *s);          // :259-267
    inline void set_preconditioner(linear_solver *p);  // :272-280
    inline void solve(dp *x, const dp *b);             // :213-233
};

// common part of the device-mirrored matrices
struct device_matrix : linear_operator {
    sigb_matrix_t mirror = nullptr;
    bool dirty = true;
    std::vector<dp> val;
    ~device_matrix() override { if (mirror) sigb_matrix_destroy(mirror); }
    virtual void sync_mirror() = 0;
    void matvec_add(const dp *x, dp *y) override { sync_mirror(); sigb_check(sigb_matvec_add(mirror, 0, x, y)); }
    void matvec_t_add(const dp *x, dp *y) override { sync_mirror(); sigb_check(sigb_matvec_add(mirror, 1, x, y)); }
    void matvec(const dp *x, dp *y) override { sync_mirror(); sigb_check(sigb_matvec(mirror, 0, x, y)); }
    void matvec_t(const dp *x, dp *y) override { sync_mirror(); sigb_check(sigb_matvec(mirror, 1, x, y)); }
    void zero() { for (dp &v : val) v = 0.0; dirty = true; }
    void scalar_multiply(dp alpha) { for (dp &v : val) v *= alpha; dirty = true; }
    void upload()
    {
        if (dirty) {
            sigb_check(sigb_matrix_set_values(mirror, val.data(), (int64_t)val.size()));
            dirty = false;
        }
    }
};

// src/matrix/formats/cs_matrices.f90 (csr_matrix :112-151, csc_matrix :156-195)
template <bool COL>
struct cs_matrix : device_matrix {
    std::shared_ptr<cs_graph> g;

    void init(int nrow_, int ncol_) { nrow = nrow_; ncol = ncol_; }
    // A%set_graph(g): share the pattern (:259-289); for a csc_matrix the graph
    // holds the columns, i.e. it is the transposed pattern
    void set_graph(std::shared_ptr<cs_graph> g_)
    {
        const int gn = COL ? ncol : nrow, gm = COL ? nrow : ncol;
        if (g_->n != gn || g_->m != gm) {
            std::printf(" Attempted to set CS matrix connectivity structure to a graph of inconsistent dimensions\n Terminating.\n");
            std::exit(1);
        }
        g = std::move(g_);
        val.assign((size_t)g->ne, 0.0);
        if (mirror) { sigb_matrix_destroy(mirror); mirror = nullptr; }
        dirty = true;
    }
    // A%copy_graph(h): a csc_matrix copies the transpose (:234-254)
    void copy_graph(const ll_graph &h)
    {
        auto gg = std::make_shared<cs_graph>();
        gg->copy(h, COL);
        set_graph(gg);
    }
    int slot(int i, int j) const { return COL ? g->find_edge(j, i) : g->find_edge(i, j); }
    void missing(int i, int j) const
    {
        std::printf(" entry (%d,%d) is not in the sparsity pattern; the reallocation path of set_value is not part of this mirror\n Terminating.\n", i, j);
        std::exit(1);
    }
    void set_value(int i, int j, dp z) { const int k = slot(i, j); if (k < 0) missing(i, j); val[(size_t)k] = z; dirty = true; }
    void add_value(int i, int j, dp z) { const int k = slot(i, j); if (k < 0) missing(i, j); val[(size_t)k] += z; dirty = true; }
    dp get_value(int i, int j) override { const int k = slot(i, j); return k < 0 ? 0.0 : val[(size_t)k]; }

    void sync_mirror() override
    {
        if (!g->mirror) {
            g->mirror = std::make_shared<graph_mirror>();
            sigb_check(sigb_cs_graph_create(g->n, g->m, g->ptr.data(), g->node.data(), COL ? SIGB_COL : SIGB_ROW,
                                            &g->mirror->h));
        }
        if (!mirror) {
            sigb_check(sigb_matrix_create(g->mirror->h, &mirror));
            dirty = true;
        }
        upload();
    }
};
using csr_matrix = cs_matrix<false>;
using csc_matrix = cs_matrix<true>;

// src/matrix/formats/ellpack_matrices.f90:28-105
struct ellpack_matrix : device_matrix {
    std::shared_ptr<ellpack_graph> g;

    void init(int nrow_, int ncol_) { nrow = nrow_; ncol = ncol_; }
    void set_graph(std::shared_ptr<ellpack_graph> g_)
    {
        g = std::move(g_);
        val.assign((size_t)g->n * g->max_d, 0.0);
        if (mirror) { sigb_matrix_destroy(mirror); mirror = nullptr; }
        dirty = true;
    }
    void copy_graph(const ll_graph &h)
    {
        auto gg = std::make_shared<ellpack_graph>();
        gg->copy(h);
        set_graph(gg);
    }
    int slot(int i, int j) const   // scans the first degrees(i) slots (:444-466)
    {
        const int d = g->degrees[(size_t)i - 1];
        int found = -1;
        for (int k = 0; k < d; k++)
            if (g->node[(size_t)(i - 1) * g->max_d + k] == j) found = (i - 1) * g->max_d + k;
        return found;
    }
    void set_value(int i, int j, dp z)
    {
        const int k = slot(i, j);
        if (k < 0) { std::printf(" entry (%d,%d) is not in the sparsity pattern\n Terminating.\n", i, j); std::exit(1); }
        val[(size_t)k] = z;
        dirty = true;
    }
    void add_value(int i, int j, dp z)
    {
        const int k = slot(i, j);
        if (k < 0) { std::printf(" entry (%d,%d) is not in the sparsity pattern\n Terminating.\n", i, j); std::exit(1); }
        val[(size_t)k] += z;
        dirty = true;
    }
    dp get_value(int i, int j) override { const int k = slot(i, j); return k < 0 ? 0.0 : val[(size_t)k]; }

    void sync_mirror() override
    {
        if (!g->mirror) {
            g->mirror = std::make_shared<graph_mirror>();
            sigb_check(sigb_ell_graph_create(g->n, g->m, g->max_d, g->node.data(), g->degrees.data(), &g->mirror->h));
        }
        if (!mirror) {
            sigb_check(sigb_matrix_create(g->mirror->h, &mirror));
            dirty = true;
        }
        upload();
    }
};

// ---------------------------------------------------------------------------
// solvers (src/solver/*.f90)
// ---------------------------------------------------------------------------
struct linear_solver {
    int nn = 0;
    bool initialized = false;
    int iterations = 0;          // cg_solver%iterations / bicgstab_solver%iterations
    sigb_solver_t dev = nullptr;
    virtual ~linear_solver() { destroy(); }

    static device_matrix &mirrored(linear_operator &A)
    {
        auto *M = dynamic_cast<device_matrix *>(&A);
        if (!M) { std::printf(" this solver needs a csr/csc/ellpack matrix\n Terminating.\n"); std::exit(1); }
        return *M;
    }
    // solver%setup(A): the non-square check and its message live in the library
    // (cg_solvers.f90:61-65 -> SIGB_ERR_NONSQUARE -> print + exit(1))
    virtual void setup(linear_operator &A)
    {
        device_matrix &M = mirrored(A);
        M.sync_mirror();
        sigb_check(sigb_solver_setup(dev, M.mirror));
        nn = A.nrow;
        initialized = true;
        iterations = 0;
    }
    // call solver%solve(A, x, b [, pc])
    virtual void solve(linear_operator &A, dp *x, const dp *b, linear_solver *pc = nullptr)
    {
        device_matrix &M = mirrored(A);
        M.sync_mirror();
        sigb_check(sigb_solver_solve(dev, M.mirror, x, b, pc ? pc->dev : nullptr));
        int64_t it = 0;
        sigb_check(sigb_solver_get_info(dev, &it, nullptr, nullptr));
        iterations = (int)it;
    }
    void set_max_iterations(int64_t cap) { sigb_check(sigb_solver_set_max_iterations(dev, cap)); }
    bool capped() const { int c = 0; sigb_solver_get_info(dev, nullptr, nullptr, &c); return c != 0; }
    virtual void destroy()
    {
        if (dev) sigb_solver_destroy(dev);
        dev = nullptr;
        initialized = false;
    }
};

struct cg_solver : linear_solver {
    dp tolerance = 1e-16;
    void set_params(dp tol = -1.0) { tolerance = tol < 0 ? 1e-16 : tol; sigb_check(sigb_solver_set_params(dev, tol)); }
};
struct bicgstab_solver : linear_solver {
    dp tolerance = 1e-16;
    void set_params(dp tol = -1.0) { tolerance = tol < 0 ? 1e-16 : tol; sigb_check(sigb_solver_set_params(dev, tol)); }
};
struct jacobi_solver : linear_solver {};

// factory functions of the reference (cg_solvers.f90:36-47, bicgstab_solvers.f90:36-47,
// jacobi_solvers.f90:26-32); tolerance < 0 selects the default 1e-16
inline linear_solver *cg(dp tolerance = -1.0)
{
    auto *s = new cg_solver();
    sigb_check(sigb_cg_create(tolerance, &s->dev));
    s->tolerance = tolerance < 0 ? 1e-16 : tolerance;
    return s;
}
inline linear_solver *bicgstab(dp tolerance = -1.0)
{
    auto *s = new bicgstab_solver();
    sigb_check(sigb_bicgstab_create(tolerance, &s->dev));
    s->tolerance = tolerance < 0 ? 1e-16 : tolerance;
    return s;
}
inline linear_solver *jacobi()
{
    auto *s = new jacobi_solver();
    sigb_check(sigb_jacobi_create(&s->dev));
    return s;
}

inline void linear_operator::set_solver(linear_solver *s) { solver = s; s->setup(*this); }
inline void linear_operator::set_preconditioner(linear_solver *p) { pc = p; p->setup(*this); }
inline void linear_operator::solve(dp *x, const dp *b) { solver->solve(*this, x, b, pc); }

// ---------------------------------------------------------------------------
// eigensolver (src/eigensolver.f90)
// ---------------------------------------------------------------------------
// T is T(3, n) column-major, Q is Q(nrow, n) column-major.  The start vector is
// Q(:,1) on entry if use_q1 (un-normalised, as after :50-51), else the library
// draws it from `seed` (the reference uses a time-seeded RNG, :47-50).
inline void lanczos(linear_operator &A, int n, dp *T, dp *Q, bool use_q1 = false, uint64_t seed = 0)
{
    device_matrix &M = linear_solver::mirrored(A);
    M.sync_mirror();
    std::vector<dp> q1;
    if (use_q1) q1.assign(Q, Q + A.nrow);
    sigb_check(sigb_lanczos(M.mirror, n, use_q1 ? q1.data() : nullptr, seed, T, Q));
}
inline void eigensolve(linear_operator &A, int n, dp *lambda, dp *V, bool use_q1 = false, uint64_t seed = 0)
{
    device_matrix &M = linear_solver::mirrored(A);
    M.sync_mirror();
    std::vector<dp> q1;
    if (use_q1) q1.assign(V, V + A.nrow);
    sigb_check(sigb_eigensolve(M.mirror, n, use_q1 ? q1.data() : nullptr, seed, lambda, V));
}

// call B%set_solver(...) first; call generalized_lanczos(A, B, T, Q)   (eigensolver.f90:95-155)
inline void generalized_lanczos(linear_operator &A, linear_operator &B, int n, dp *T, dp *Q, bool use_q1 = false,
                                uint64_t seed = 0)
{
    device_matrix &MA = linear_solver::mirrored(A), &MB = linear_solver::mirrored(B);
    if (!B.solver) { std::printf(" generalized_lanczos: B has no solver set\n Terminating.\n"); std::exit(1); }
    MA.sync_mirror();
    MB.sync_mirror();
    std::vector<dp> q1;
    if (use_q1) q1.assign(Q, Q + A.nrow);
    sigb_check(sigb_generalized_lanczos(MA.mirror, MB.mirror, B.solver->dev, B.pc ? B.pc->dev : nullptr, n,
                                        use_q1 ? q1.data() : nullptr, seed, T, Q));
}
inline void generalized_eigensolve(linear_operator &A, linear_operator &B, int n, dp *lambda, dp *V,
                                   bool use_q1 = false, uint64_t seed = 0)
{
    device_matrix &MA = linear_solver::mirrored(A), &MB = linear_solver::mirrored(B);
    if (!B.solver) { std::printf(" generalized_eigensolve: B has no solver set\n Terminating.\n"); std::exit(1); }
    MA.sync_mirror();
    MB.sync_mirror();
    std::vector<dp> q1;
    if (use_q1) q1.assign(V, V + A.nrow);
    sigb_check(sigb_generalized_eigensolve(MA.mirror, MB.mirror, B.solver->dev, B.pc ? B.pc->dev : nullptr, n,
                                           use_q1 ? q1.data() : nullptr, seed, lambda, V));
}

}  // namespace sigma
