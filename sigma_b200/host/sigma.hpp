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
//   A%copy_matrix(B, trans)                         same name (built on the device)
//   A%matvec / matvec_t / matvec_add / matvec_t_add same names (linear_operator)
//   A%set_solver / set_preconditioner / solve       same names
//   cg(tol), bicgstab(tol), jacobi(), ldu()         same names -> linear_solver*
//   solver%setup / solve(A,x,b[,pc]) / destroy      same names
//   lanczos(A,T,Q), eigensolve(A,lambda,V)          same names
//   L = A + B, L = A * B, L = adjoint(A)            same (operator_sum/_product/_adjoint)
//   type(sparse_matrix) composite: set_dimensions,  sparse_matrix, same names
//   set_block_sizes, set_submatrix, set/add/get
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

// Single-process multi-GPU mode.  After sigma::use_gpus(ndev) (ndev <= 0: all visible GPUs) every
// square csr_matrix / csc_matrix / ellpack_matrix is mirrored as one row block per GPU (sigb_mgpu_csr_create;
// csc and ellpack through their rows in the order their own matvec loops accumulate them): A%matvec, A%matvec_add
// solver%solve(A, x, b [, pc]) with cg / bicgstab / jacobi, lanczos and eigensolve run on all of them, with
// the caller's whole vectors and no change to the calling program.  What a multi-GPU mirror cannot do
// (matvec_t, copies, expressions, ldu, the generalized eigensolvers) is what the library refuses for it.
inline int &gpus_in_use() { static int n = 0; return n; }
inline int use_gpus(int ndev = 0)
{
    sigb_check(sigb_mgpu_init(ndev));
    sigb_check(sigb_mgpu_device_count(&gpus_in_use()));
    return gpus_in_use();
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
    // device mirrors, dropped by every mutator: [0] as the row pattern of a
    // csr_matrix, [1] as the column pattern of a csc_matrix (one graph object can
    // serve both at once: test/matrix_test_composite.f90:176-186)
    std::shared_ptr<graph_mirror> mirror[2];

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
        mirror[0].reset();
        mirror[1].reset();
    }
    int find_edge(int i, int j) const   // 0-based position in node, -1 if absent
    {
        for (int k = ptr[(size_t)i - 1]; k <= ptr[(size_t)i] - 1; k++)
            if (node[(size_t)k - 1] == j) return k - 1;
        return -1;
    }
    // g%add_edge(i, j)   (cs_add_edge :400-442): the new neighbour goes to the END of line i,
    // everything behind it moves up by one.  Returns the 0-based position it took (-1 if the
    // edge was there already).  Host-side index work; the device mirrors are dropped.
    int add_edge(int i, int j)
    {
        if (find_edge(i, j) >= 0) return -1;
        const int indx = ptr[(size_t)i];                       // 1-based ptr(i + 1)
        node.insert(node.begin() + (indx - 1), (int32_t)j);
        for (int k = i; k <= n; k++) ptr[(size_t)k] += 1;      // ptr(i+1 .. n+1)
        ne++;
        max_d = std::max(max_d, ptr[(size_t)i] - ptr[(size_t)i - 1]);
        mirror[0].reset();
        mirror[1].reset();
        return indx - 1;
    }
    // g%left_permute(p, edge_p)   (cs_graph_left_permute :499-549): line i becomes line p(i), each
    // line keeps its stored order.  from[i-1] / to[i-1] / len[i-1] (0-based offsets) say where the
    // entries of old line i went -- the reference's compressed edge permutation.
    void left_permute(const std::vector<int> &p, std::vector<int> &from, std::vector<int> &to, std::vector<int> &len)
    {
        std::vector<int32_t> nptr((size_t)n + 1, 0), nnode((size_t)ne);
        from.assign((size_t)n, 0); to.assign((size_t)n, 0); len.assign((size_t)n, 0);
        for (int i = 1; i <= n; i++) nptr[(size_t)p[(size_t)i - 1]] = ptr[(size_t)i] - ptr[(size_t)i - 1];
        nptr[0] = 1;
        for (int i = 1; i <= n; i++) nptr[(size_t)i] += nptr[(size_t)i - 1];
        for (int i = 1; i <= n; i++) {
            const int d = ptr[(size_t)i] - ptr[(size_t)i - 1], src = ptr[(size_t)i - 1] - 1, dst = nptr[(size_t)p[(size_t)i - 1] - 1] - 1;
            for (int k = 0; k < d; k++) nnode[(size_t)dst + k] = node[(size_t)src + k];
            from[(size_t)i - 1] = src; to[(size_t)i - 1] = dst; len[(size_t)i - 1] = d;
        }
        ptr.swap(nptr);
        node.swap(nnode);
        mirror[0].reset();
        mirror[1].reset();
    }
    // g%right_permute(p)   (cs_graph_right_permute :554-570): the ids are relabelled in place
    void right_permute(const std::vector<int> &p)
    {
        for (int32_t &c : node) c = p[(size_t)c - 1];
        mirror[0].reset();
        mirror[1].reset();
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
    bool connected(int i, int j) const
    {
        for (int k = 0; k < degrees[(size_t)i - 1]; k++)
            if (node[(size_t)(i - 1) * max_d + k] == j) return true;
        return false;
    }
    // g%add_edge(i, j)   (ellpack_add_edge :379-413, add_edge_with_reallocation :600-639): fill the
    // rest of row i with j while there is room, otherwise widen every row by one slot (padding =
    // copy of the last neighbour).  Returns true when max_d grew.
    bool add_edge(int i, int j)
    {
        if (connected(i, j)) return false;
        const int k = degrees[(size_t)i - 1];
        bool widened = false;
        if (k < max_d) {
            for (int l = k; l < max_d; l++) node[(size_t)(i - 1) * max_d + l] = j;
        } else {
            std::vector<int32_t> wide((size_t)n * (max_d + 1), 0);
            for (int r = 0; r < n; r++) {
                for (int l = 0; l < max_d; l++) wide[(size_t)r * (max_d + 1) + l] = node[(size_t)r * max_d + l];
                if (max_d > 0) wide[(size_t)r * (max_d + 1) + max_d] = node[(size_t)r * max_d + max_d - 1];
            }
            wide[(size_t)(i - 1) * (max_d + 1) + max_d] = j;
            node.swap(wide);
            max_d += 1;
            widened = true;
        }
        degrees[(size_t)i - 1] = k + 1;
        ne++;
        mirror.reset();
        return widened;
    }
    // ellpack_graph_left_permute :486-518 / ellpack_graph_right_permute :523-541
    void left_permute(const std::vector<int> &p)
    {
        std::vector<int32_t> nnode(node.size()), ndeg(degrees.size());
        for (int i = 1; i <= n; i++) {
            for (int l = 0; l < max_d; l++) nnode[(size_t)(p[(size_t)i - 1] - 1) * max_d + l] = node[(size_t)(i - 1) * max_d + l];
            ndeg[(size_t)p[(size_t)i - 1] - 1] = degrees[(size_t)i - 1];
        }
        node.swap(nnode);
        degrees.swap(ndeg);
        mirror.reset();
    }
    void right_permute(const std::vector<int> &p)
    {
        for (int32_t &c : node) if (c != 0) c = p[(size_t)c - 1];
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
    int reference_count = 0;                     // :285-302
    linear_solver *solver = nullptr, *pc = nullptr;
    virtual ~linear_operator() {}
    // The device mirror of this operator, brought up to date (values re-uploaded
    // after host mutations).  Every operator of this header has one: a stored
    // matrix mirrors its arrays, an expression is built from its operands' mirrors.
    virtual sigb_matrix_t device_handle() = 0;
    // matvec_add / matvec_t_add (deferred in the reference, :35-36) and the
    // non-overridable matvec / matvec_t (:185-208; the zero-fill is fused away)
    virtual void matvec_add(const dp *x, dp *y) { sigb_check(sigb_matvec_add(device_handle(), 0, x, y)); }
    virtual void matvec_t_add(const dp *x, dp *y) { sigb_check(sigb_matvec_add(device_handle(), 1, x, y)); }
    virtual void matvec(const dp *x, dp *y) { sigb_check(sigb_matvec(device_handle(), 0, x, y)); }
    virtual void matvec_t(const dp *x, dp *y) { sigb_check(sigb_matvec(device_handle(), 1, x, y)); }
    // default get_value (:168-181): column j of the operator through a matvec
    virtual dp get_value(int i, int j)
    {
        std::vector<dp> x((size_t)ncol, 0.0), y((size_t)nrow, 0.0);
        x[(size_t)j - 1] = 1.0;
        matvec(x.data(), y.data());
        return y[(size_t)i - 1];
    }
    void add_reference() { reference_count++; }
    void remove_reference() { reference_count--; }
    virtual void destroy() {}
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
    sigb_matrix_t device_handle() override { sync_mirror(); return mirror; }
    virtual void set_value(int i, int j, dp z) = 0;
    virtual void add_value(int i, int j, dp z) = 0;
    virtual bool in_pattern(int i, int j) const = 0;
    // A batch of `call A%add_value(is(c), js(c), zs(c))` applied IN ORDER on the device
    // (an assembly loop, examples/fem.f90:43-47; add_multiple_values cs_matrices.f90:934-967);
    // bit-identical to issuing the calls one by one.  The host copy of the values is
    // refreshed from the device afterwards.
    void add_values(const std::vector<int32_t> &is, const std::vector<int32_t> &js, const std::vector<dp> &zs)
    {
        if (is.size() != js.size() || is.size() != zs.size()) {
            std::printf(" add_values: index and value arrays differ in length\n Terminating.\n");
            std::exit(1);
        }
        sync_mirror();
        sigb_check(sigb_matrix_add_values(mirror, (int64_t)is.size(), is.data(), js.data(), zs.data()));
        sigb_check(sigb_matrix_get_arrays(mirror, nullptr, nullptr, val.data()));
    }
    // call A%add_multiple_values(is, js, B)   (cs_matrices.f90:934-967, ellpack likewise): the
    // add_value stream (is(k), js(l), B(k, l)), k outer, l inner; B is given row by row
    // (B[k * size(js) + l]).  One device call per block; an assembly loop gathers the blocks
    // of a whole mesh into one add_values batch instead.
    void add_multiple_values(const std::vector<int32_t> &is, const std::vector<int32_t> &js, const std::vector<dp> &B)
    {
        if (B.size() != is.size() * js.size()) {
            std::printf(" add_multiple_values: B must be size(is) x size(js)\n Terminating.\n");
            std::exit(1);
        }
        std::vector<int32_t> ii, jj;
        ii.reserve(B.size());
        jj.reserve(B.size());
        bool all_in = true;
        for (int32_t i : is)
            for (int32_t j : js) { ii.push_back(i); jj.push_back(j); all_in = all_in && in_pattern(i, j); }
        if (all_in) { add_values(ii, jj, B); return; }
        // an entry outside the pattern makes the graph grow: the reference loop on the host
        for (size_t c = 0; c < B.size(); c++) add_value(ii[c], jj[c], B[c]);
    }
    void zero() { for (dp &v : val) v = 0.0; dirty = true; }
    void scalar_multiply(dp alpha) { for (dp &v : val) v *= alpha; dirty = true; }
    // multi-GPU mirrors of csc / ellpack matrices hold the ROWS of the matrix (the sharded path takes csr
    // blocks): mg_perm[k] = stored index of the k-th entry in row order, mg_val the values in that order
    std::vector<int64_t> mg_perm;
    std::vector<dp> mg_val;
    void upload()
    {
        if (dirty) {
            if (!mg_perm.empty()) {
                mg_val.resize(mg_perm.size());
                for (size_t k = 0; k < mg_perm.size(); k++) mg_val[k] = val[(size_t)mg_perm[k]];
                sigb_check(sigb_matrix_set_values(mirror, mg_val.data(), (int64_t)mg_val.size()));
            } else {
                sigb_check(sigb_matrix_set_values(mirror, val.data(), (int64_t)val.size()));
            }
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
    // set_matrix_value_with_reallocation (default_sparse_matrix_kernels.f90:176-229) for an entry
    // outside the pattern: the graph gains the edge (a csc_matrix adds (j, i), cs_matrices.f90:992),
    // every stored value moves to its new index and the new slot receives z.  Host-side, like the
    // reference; the device mirrors of the graph and of the matrix are dropped and rebuilt by the
    // next matvec / solve.  A graph shared with other matrices is cloned first (the reference
    // mutates the shared object and leaves the other matrices' value arrays behind).
    void grow(int i, int j, dp z)
    {
        if (i < 1 || i > nrow || j < 1 || j > ncol) {
            std::printf(" entry (%d,%d) is outside the %d x %d matrix\n Terminating.\n", i, j, nrow, ncol);
            std::exit(1);
        }
        if (g.use_count() > 1) g = std::make_shared<cs_graph>(*g);
        const int pos = COL ? g->add_edge(j, i) : g->add_edge(i, j);
        val.insert(val.begin() + pos, z);
        if (mirror) { sigb_matrix_destroy(mirror); mirror = nullptr; }
        dirty = true;
    }
    void set_value(int i, int j, dp z) override { const int k = slot(i, j); if (k < 0) { grow(i, j, z); return; } val[(size_t)k] = z; dirty = true; }
    void add_value(int i, int j, dp z) override { const int k = slot(i, j); if (k < 0) { grow(i, j, z); return; } val[(size_t)k] += z; dirty = true; }
    bool in_pattern(int i, int j) const override { return slot(i, j) >= 0; }
    dp get_value(int i, int j) override { const int k = slot(i, j); return k < 0 ? 0.0 : val[(size_t)k]; }
    // call A%left_permute(p) / A%right_permute(p)   (cs_matrices.f90:471-490): rows (columns) i move to
    // p(i).  A csr_matrix moves its lines for a left permutation and relabels its ids for a right one
    // (graph_leftperm / graph_rightperm, default_sparse_matrix_kernels.f90:234-277); a csc_matrix the
    // other way round (:186-187).  Host-side, like the reference; the device mirrors are dropped.
    void move_lines(const std::vector<int> &p)
    {
        std::vector<int> from, to, len;
        g->left_permute(p, from, to, len);
        std::vector<dp> nval(val.size());
        for (size_t b = 0; b < from.size(); b++)
            for (int k = 0; k < len[b]; k++) nval[(size_t)to[b] + k] = val[(size_t)from[b] + k];
        val.swap(nval);
    }
    void permute(const std::vector<int> &p, bool lines)
    {
        if (g.use_count() > 1) g = std::make_shared<cs_graph>(*g);
        if (lines) move_lines(p); else g->right_permute(p);
        if (mirror) { sigb_matrix_destroy(mirror); mirror = nullptr; }
        dirty = true;
    }
    void left_permute(const std::vector<int> &p) { permute(p, !COL); }
    void right_permute(const std::vector<int> &p) { permute(p, COL); }
    // call A%get_row(nodes, slice, k) / A%get_column(nodes, slice, k)   (cs_matrices.f90:370-392): the
    // stored line itself when the format keeps that direction (get_slice_contiguous,
    // default_sparse_matrix_kernels.f90:94-123), a scan over all lines otherwise
    // (get_slice_discontiguous :128-165).  Host-side reads of the mirror's own arrays.
    void line_slice(int k, std::vector<int32_t> &nodes, std::vector<dp> &slice) const
    {
        nodes.clear(); slice.clear();
        for (int q = g->ptr[(size_t)k - 1] - 1; q < g->ptr[(size_t)k] - 1; q++) { nodes.push_back(g->node[(size_t)q]); slice.push_back(val[(size_t)q]); }
    }
    void cross_slice(int k, std::vector<int32_t> &nodes, std::vector<dp> &slice) const
    {
        nodes.clear(); slice.clear();
        for (int l = 1; l <= g->n; l++) {
            const int q = g->find_edge(l, k);
            if (q >= 0) { nodes.push_back(l); slice.push_back(val[(size_t)q]); }
        }
    }
    void get_row(std::vector<int32_t> &nodes, std::vector<dp> &slice, int k) const { if (COL) cross_slice(k, nodes, slice); else line_slice(k, nodes, slice); }
    void get_column(std::vector<int32_t> &nodes, std::vector<dp> &slice, int k) const { if (COL) line_slice(k, nodes, slice); else cross_slice(k, nodes, slice); }

    // The ROWS of a csc_matrix, each in the order csc_matvec_add reaches it (columns ascending, a column's
    // entries in stored order: cs_matrices.f90:627-647 -- the stable transpose the one-GPU path builds on the
    // device), with perm[k] = stored index of the k-th entry in row order.  Host-side index work.
    void rows_in_matvec_order(std::vector<int32_t> &ptr1, std::vector<int32_t> &rnode, std::vector<int64_t> &perm) const
    {
        const int nlines = g->n, nrows = COL ? g->m : g->n;
        ptr1.assign((size_t)nrows + 1, 0);
        std::vector<int32_t> count((size_t)nrows + 1, 0);
        for (int q = 0; q < g->ne; q++) count[(size_t)g->node[(size_t)q]]++;
        ptr1[0] = 1;
        for (int i = 1; i <= nrows; i++) ptr1[(size_t)i] = ptr1[(size_t)i - 1] + count[(size_t)i];
        std::vector<int32_t> fill(ptr1.begin(), ptr1.end() - 1);                  // next free slot of row i at fill[i-1]
        rnode.assign((size_t)g->ne, 0);
        perm.assign((size_t)g->ne, 0);
        for (int j = 1; j <= nlines; j++)
            for (int q = g->ptr[(size_t)j - 1] - 1; q < g->ptr[(size_t)j] - 1; q++) {
                const int i = g->node[(size_t)q];
                const int32_t slot = fill[(size_t)i - 1]++;
                rnode[(size_t)slot - 1] = j;
                perm[(size_t)slot - 1] = q;
            }
    }

    void sync_mirror() override
    {
        if (!mirror && !COL && gpus_in_use() > 0 && g->n == g->m) {
            // multi-GPU mode: the pattern goes in whole, the library shards it (row blocks, halo and
            // send lists derived from the graph)
            sigb_check(sigb_mgpu_csr_create(g->n, g->ptr.data(), g->node.data(), &mirror));
            dirty = true;
        }
        if (!mirror && COL && gpus_in_use() > 0 && g->n == g->m) {
            // a csc_matrix in multi-GPU mode is sharded like a csr_matrix, through its rows (rows_in_matvec_order)
            std::vector<int32_t> ptr1, rnode;
            rows_in_matvec_order(ptr1, rnode, mg_perm);
            sigb_check(sigb_mgpu_csr_create(g->n, ptr1.data(), rnode.data(), &mirror));
            dirty = true;
        }
        if (!mirror) {
            std::shared_ptr<graph_mirror> &gm = g->mirror[COL ? 1 : 0];
            if (!gm) {
                gm = std::make_shared<graph_mirror>();
                sigb_check(sigb_cs_graph_create(g->n, g->m, g->ptr.data(), g->node.data(),
                                                COL ? SIGB_COL : SIGB_ROW, &gm->h));
            }
            sigb_check(sigb_matrix_create(gm->h, &mirror));
            dirty = true;
        }
        upload();
    }

    // call A%copy_matrix(B, trans)   (cs_matrix_copy_matrix :294-322): the graph is
    // built and the values are placed ON THE DEVICE (sigb_matrix_copy), then read
    // back into this object's own arrays; the device copy becomes the mirror.
    void copy_matrix(linear_operator &B, bool trans = false)
    {
        const int nr = trans ? B.ncol : B.nrow, nc = trans ? B.nrow : B.ncol;
        if (nrow != nr || ncol != nc) {
            std::printf(" Attempted to copy a matrix of inconsistent dimensions\n Terminating.\n");
            std::exit(1);
        }
        sigb_matrix_t h = nullptr;
        sigb_check(sigb_matrix_copy(B.device_handle(), COL ? SIGB_FMT_CSC : SIGB_FMT_CSR, trans ? 1 : 0, &h));
        int32_t n = 0, m = 0, md = 0;
        int64_t ne = 0;
        sigb_check(sigb_matrix_get_format(h, nullptr, &n, &m, &ne, &md));
        g = std::make_shared<cs_graph>();
        g->n = n;
        g->m = m;
        g->ne = (int)ne;
        g->max_d = md;
        g->ptr.assign((size_t)n + 1, 1);
        g->node.assign((size_t)ne, 0);
        val.assign((size_t)ne, 0.0);
        sigb_check(sigb_matrix_get_arrays(h, g->ptr.data(), g->node.data(), val.data()));
        if (mirror) sigb_matrix_destroy(mirror);
        mirror = h;
        dirty = false;
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
    // set_unallocated_matrix_value (ellpack_matrices.f90:801-825): widen val by one zero slot per
    // row when row i is full, add the edge, val(d + 1, i) = z.  Host-side; mirrors dropped.
    void grow(int i, int j, dp z)
    {
        if (i < 1 || i > nrow || j < 1 || j > ncol) {
            std::printf(" entry (%d,%d) is outside the %d x %d matrix\n Terminating.\n", i, j, nrow, ncol);
            std::exit(1);
        }
        if (g.use_count() > 1) g = std::make_shared<ellpack_graph>(*g);
        const int d = g->degrees[(size_t)i - 1], old_w = g->max_d;
        if (g->add_edge(i, j)) {
            std::vector<dp> wide((size_t)g->n * g->max_d, 0.0);
            for (int r = 0; r < g->n; r++)
                for (int l = 0; l < old_w; l++) wide[(size_t)r * g->max_d + l] = val[(size_t)r * old_w + l];
            val.swap(wide);
        }
        val[(size_t)(i - 1) * g->max_d + d] = z;
        if (mirror) { sigb_matrix_destroy(mirror); mirror = nullptr; }
        dirty = true;
    }
    void set_value(int i, int j, dp z) override
    {
        const int k = slot(i, j);
        if (k < 0) { grow(i, j, z); return; }
        val[(size_t)k] = z;
        dirty = true;
    }
    void add_value(int i, int j, dp z) override
    {
        const int k = slot(i, j);
        if (k < 0) { grow(i, j, z); return; }
        val[(size_t)k] += z;
        dirty = true;
    }
    bool in_pattern(int i, int j) const override { return slot(i, j) >= 0; }
    dp get_value(int i, int j) override { const int k = slot(i, j); return k < 0 ? 0.0 : val[(size_t)k]; }
    // ellpack_matrix_left_permute / _right_permute (ellpack_matrices.f90:601-630), host-side
    void left_permute(const std::vector<int> &p)
    {
        if (g.use_count() > 1) g = std::make_shared<ellpack_graph>(*g);
        std::vector<dp> nval(val.size());
        for (int i = 1; i <= g->n; i++)
            for (int l = 0; l < g->max_d; l++) nval[(size_t)(p[(size_t)i - 1] - 1) * g->max_d + l] = val[(size_t)(i - 1) * g->max_d + l];
        val.swap(nval);
        g->left_permute(p);
        if (mirror) { sigb_matrix_destroy(mirror); mirror = nullptr; }
        dirty = true;
    }
    void right_permute(const std::vector<int> &p)
    {
        if (g.use_count() > 1) g = std::make_shared<ellpack_graph>(*g);
        g->right_permute(p);
        if (mirror) { sigb_matrix_destroy(mirror); mirror = nullptr; }
        dirty = true;
    }
    // get_row: the first degrees(k) slots of row k; get_column: a scan over the rows
    void get_row(std::vector<int32_t> &nodes, std::vector<dp> &slice, int k) const
    {
        nodes.clear(); slice.clear();
        for (int l = 0; l < g->degrees[(size_t)k - 1]; l++) {
            nodes.push_back(g->node[(size_t)(k - 1) * g->max_d + l]);
            slice.push_back(val[(size_t)(k - 1) * g->max_d + l]);
        }
    }
    void get_column(std::vector<int32_t> &nodes, std::vector<dp> &slice, int k) const
    {
        nodes.clear(); slice.clear();
        for (int i = 1; i <= g->n; i++) {
            const int q = slot(i, k);
            if (q >= 0) { nodes.push_back(i); slice.push_back(val[(size_t)q]); }
        }
    }

    void sync_mirror() override
    {
        if (!mirror && gpus_in_use() > 0 && g->n == g->m) {
            // an ellpack_matrix in multi-GPU mode is sharded like a csr_matrix, through its rows without padding
            std::vector<int32_t> ptr1, rnode;
            rows_without_padding(ptr1, rnode, mg_perm);
            sigb_check(sigb_mgpu_csr_create(g->n, ptr1.data(), rnode.data(), &mirror));
            dirty = true;
        }
        if (!mirror) {
            if (!g->mirror) {
                g->mirror = std::make_shared<graph_mirror>();
                sigb_check(sigb_ell_graph_create(g->n, g->m, g->max_d, g->node.data(), g->degrees.data(),
                                                 &g->mirror->h));
            }
            sigb_check(sigb_matrix_create(g->mirror->h, &mirror));
            dirty = true;
        }
        upload();
    }

    // The rows of an ellpack_matrix without the padding slots (their values are zero: ellpack_matvec_add adds
    // 0 * x for them, ellpack_matrices.f90:655-658), with perm[k] = index into val of the k-th entry.
    void rows_without_padding(std::vector<int32_t> &ptr1, std::vector<int32_t> &rnode, std::vector<int64_t> &perm) const
    {
        const int n = g->n;
        ptr1.assign((size_t)n + 1, 1);
        rnode.clear();
        perm.clear();
        for (int i = 1; i <= n; i++) {
            for (int k = 0; k < g->degrees[(size_t)i - 1]; k++) {
                rnode.push_back(g->node[(size_t)(i - 1) * g->max_d + k]);
                perm.push_back((int64_t)(i - 1) * g->max_d + k);
            }
            ptr1[(size_t)i] = (int32_t)rnode.size() + 1;
        }
    }

    // call A%copy_matrix(B, trans)   (ellpack_matrix_copy_matrix :169-198), on the device
    void copy_matrix(linear_operator &B, bool trans = false)
    {
        const int nr = trans ? B.ncol : B.nrow, nc = trans ? B.nrow : B.ncol;
        if (nrow != nr || ncol != nc) {
            std::printf(" Attempted to copy a matrix of inconsistent dimensions\n Terminating.\n");
            std::exit(1);
        }
        sigb_matrix_t h = nullptr;
        sigb_check(sigb_matrix_copy(B.device_handle(), SIGB_FMT_ELLPACK, trans ? 1 : 0, &h));
        int32_t n = 0, m = 0, md = 0;
        int64_t ne = 0;
        sigb_check(sigb_matrix_get_format(h, nullptr, &n, &m, &ne, &md));
        g = std::make_shared<ellpack_graph>();
        g->n = n;
        g->m = m;
        g->ne = (int)ne;
        g->max_d = md;
        g->node.assign((size_t)n * md, 0);
        g->degrees.assign((size_t)n, 0);
        val.assign((size_t)n * md, 0.0);
        sigb_check(sigb_matrix_get_arrays(h, g->degrees.data(), g->node.data(), val.data()));
        if (mirror) sigb_matrix_destroy(mirror);
        mirror = h;
        dirty = false;
    }
};

// ---------------------------------------------------------------------------
// operator expressions (src/linear_operator/linear_operator_{sums,products,
// adjoints}.f90) and the block composite (src/matrix/sparse_matrix_composites.f90)
// ---------------------------------------------------------------------------
// Common part: the device expression is (re)built from the operands' current
// mirrors whenever one of them has been replaced (set_graph drops a mirror);
// refreshing the operands' values is all that is needed otherwise.
struct operator_expression : linear_operator {
    std::vector<linear_operator *> operands;
    std::vector<sigb_matrix_t> built_from;
    sigb_matrix_t handle = nullptr;
    ~operator_expression() override { if (handle) sigb_matrix_destroy(handle); }
    virtual void build(const std::vector<sigb_matrix_t> &h) = 0;
    sigb_matrix_t device_handle() override
    {
        std::vector<sigb_matrix_t> h;
        for (linear_operator *op : operands) h.push_back(op->device_handle());
        if (!handle || h != built_from) {
            if (handle) sigb_matrix_destroy(handle);
            handle = nullptr;
            build(h);
            built_from = h;
        }
        return handle;
    }
    // operator_sum_destroy linear_operator_sums.f90:136-159 (same for the others):
    // drop the references; an operand nobody else holds is destroyed with us
    void destroy() override
    {
        if (handle) { sigb_matrix_destroy(handle); handle = nullptr; }
        for (linear_operator *op : operands) {
            op->remove_reference();
            if (op->reference_count <= 0) { op->destroy(); delete op; }
        }
        operands.clear();
        built_from.clear();
        reference_count = 0;
    }
    void adopt(linear_operator &A) { operands.push_back(&A); A.add_reference(); }
};

struct operator_sum : operator_expression {
    void build(const std::vector<sigb_matrix_t> &h) override { sigb_check(sigb_operator_sum(h[0], h[1], &handle)); }
    dp get_value(int i, int j) override   // :79-95
    {
        dp z = 0.0;
        for (linear_operator *op : operands) z = z + op->get_value(i, j);
        return z;
    }
};
struct operator_product : operator_expression {
    void build(const std::vector<sigb_matrix_t> &h) override { sigb_check(sigb_operator_product(h[0], h[1], &handle)); }
};
struct operator_adjoint : operator_expression {
    void build(const std::vector<sigb_matrix_t> &h) override { sigb_check(sigb_operator_adjoint(h[0], &handle)); }
    dp get_value(int i, int j) override { return operands[0]->get_value(j, i); }   // :49-57
};

// add_operators (linear_operator_sums.f90:38-72), multiply_operators
// (linear_operator_products.f90:39-73), adjoint (linear_operator_adjoints.f90:28-44).
// Stack objects handed in as operands must carry a reference of their own
// (call A.add_reference() first) or destroy() would try to delete them.
inline linear_operator *add_operators(linear_operator &A, linear_operator &B)
{
    if (A.nrow != B.nrow || A.ncol != B.ncol) {
        std::printf(" Dimensions of operators to be summed are not consistent\n");
        std::exit(1);
    }
    auto *C = new operator_sum();
    C->nrow = A.nrow;
    C->ncol = A.ncol;
    C->adopt(A);
    C->adopt(B);
    return C;
}
inline linear_operator *multiply_operators(linear_operator &A, linear_operator &B)
{
    if (A.ncol != B.nrow) {
        std::printf(" Dimensions of operators to be multiplied are inconsistent\n");
        std::exit(1);
    }
    auto *C = new operator_product();
    C->nrow = A.nrow;
    C->ncol = B.ncol;
    C->adopt(A);
    C->adopt(B);
    return C;
}
inline linear_operator *adjoint(linear_operator &A)
{
    auto *B = new operator_adjoint();
    B->nrow = A.ncol;
    B->ncol = A.nrow;
    B->adopt(A);
    return B;
}
inline linear_operator *operator+(linear_operator &A, linear_operator &B) { return add_operators(A, B); }
inline linear_operator *operator*(linear_operator &A, linear_operator &B) { return multiply_operators(A, B); }

// type(sparse_matrix) as a composite of sub-matrices
// (sparse_matrix_composites.f90:41-49); every block is set with set_submatrix
struct sparse_matrix : operator_expression {
    int num_row_mats = 0, num_col_mats = 0;
    std::vector<int> row_ptr, col_ptr;     // 1-based block offsets, as in the reference

    void set_dimensions(int nrow_, int ncol_) { nrow = nrow_; ncol = ncol_; }     // :181-198
    void set_block_sizes(const std::vector<int> &rows, const std::vector<int> &cols)   // :226-262
    {
        num_row_mats = (int)rows.size();
        num_col_mats = (int)cols.size();
        row_ptr.assign(rows.size() + 1, 1);
        col_ptr.assign(cols.size() + 1, 1);
        for (size_t it = 0; it < rows.size(); it++) row_ptr[it + 1] = row_ptr[it] + rows[it];
        for (size_t jt = 0; jt < cols.size(); jt++) col_ptr[jt + 1] = col_ptr[jt] + cols[jt];
        operands.assign(rows.size() * cols.size(), nullptr);
    }
    linear_operator *&sub(int it, int jt) { return operands[(size_t)(it - 1) * num_col_mats + (jt - 1)]; }
    void set_submatrix(int it, int jt, linear_operator &B)   // :1031-1065
    {
        const int r = row_ptr[(size_t)it] - row_ptr[(size_t)it - 1], c = col_ptr[(size_t)jt] - col_ptr[(size_t)jt - 1];
        if (B.nrow != r || B.ncol != c) {
            std::printf(" Inconsistent dimensions for sub-matrix\n");
            std::exit(1);
        }
        sub(it, jt) = &B;
        B.add_reference();
    }
    int get_owning_row_matrix(int i) const      // :1235-1246
    {
        int it = 1;
        for (; it <= num_row_mats; it++)
            if (row_ptr[(size_t)it - 1] <= i && row_ptr[(size_t)it] > i) break;
        return it;
    }
    int get_owning_column_matrix(int j) const   // :1251-1262
    {
        int jt = 1;
        for (; jt <= num_col_mats; jt++)
            if (col_ptr[(size_t)jt - 1] <= j && col_ptr[(size_t)jt] > j) break;
        return jt;
    }
    dp get_value(int i, int j) override         // :465-485
    {
        const int it = get_owning_row_matrix(i), jt = get_owning_column_matrix(j);
        return sub(it, jt)->get_value(i - row_ptr[(size_t)it - 1] + 1, j - col_ptr[(size_t)jt - 1] + 1);
    }
    dp get(int it, int jt, int i, int j) { return sub(it, jt)->get_value(i, j); }   // get_submat_value :490-498
    device_matrix &leaf(int it, int jt)
    {
        auto *M = dynamic_cast<device_matrix *>(sub(it, jt));
        if (!M) { std::printf(" sub-matrix (%d,%d) is not a stored matrix\n Terminating.\n", it, jt); std::exit(1); }
        return *M;
    }
    void set(int it, int jt, int i, int j, dp z) { leaf(it, jt).set_value(i, j, z); }   // set_submat_value :904-912
    void add(int it, int jt, int i, int j, dp z) { leaf(it, jt).add_value(i, j, z); }   // add_submat_value :917-925
    sigb_matrix_t device_handle() override
    {
        for (linear_operator *op : operands)
            if (!op) { std::printf(" composite matrix has an unset sub-matrix\n Terminating.\n"); std::exit(1); }
        return operator_expression::device_handle();
    }
    void build(const std::vector<sigb_matrix_t> &h) override
    {
        std::vector<int32_t> rows((size_t)num_row_mats), cols((size_t)num_col_mats);
        for (int it = 0; it < num_row_mats; it++) rows[(size_t)it] = row_ptr[(size_t)it + 1] - row_ptr[(size_t)it];
        for (int jt = 0; jt < num_col_mats; jt++) cols[(size_t)jt] = col_ptr[(size_t)jt + 1] - col_ptr[(size_t)jt];
        sigb_check(sigb_composite_create(num_row_mats, num_col_mats, rows.data(), cols.data(), h.data(), &handle));
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

    // solver%setup(A): the non-square check and its message live in the library
    // (cg_solvers.f90:61-65 -> SIGB_ERR_NONSQUARE -> print + exit(1))
    virtual void setup(linear_operator &A)
    {
        sigb_check(sigb_solver_setup(dev, A.device_handle()));
        nn = A.nrow;
        initialized = true;
        iterations = 0;
    }
    // call solver%solve(A, x, b [, pc])
    virtual void solve(linear_operator &A, dp *x, const dp *b, linear_solver *pc = nullptr)
    {
        sigb_check(sigb_solver_solve(dev, A.device_handle(), x, b, pc ? pc->dev : nullptr));
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
// src/solver/ldu_solvers.f90:35-58: always incomplete, level 0 (:145,151)
struct sparse_ldu_solver : linear_solver {
    bool incomplete = true;
    int level = 0;
};

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

// ldu(incomplete, level)   (ldu_solvers.f90:73-86)
inline linear_solver *ldu(bool incomplete = true, int level = 0)
{
    (void)incomplete;
    (void)level;
    auto *s = new sparse_ldu_solver();
    sigb_check(sigb_ldu_create(&s->dev));
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
    std::vector<dp> q1;
    if (use_q1) q1.assign(Q, Q + A.nrow);
    sigb_check(sigb_lanczos(A.device_handle(), n, use_q1 ? q1.data() : nullptr, seed, T, Q));
}
inline void eigensolve(linear_operator &A, int n, dp *lambda, dp *V, bool use_q1 = false, uint64_t seed = 0)
{
    std::vector<dp> q1;
    if (use_q1) q1.assign(V, V + A.nrow);
    sigb_check(sigb_eigensolve(A.device_handle(), n, use_q1 ? q1.data() : nullptr, seed, lambda, V));
}

// call B%set_solver(...) first; call generalized_lanczos(A, B, T, Q)   (eigensolver.f90:95-155)
inline void generalized_lanczos(linear_operator &A, linear_operator &B, int n, dp *T, dp *Q, bool use_q1 = false,
                                uint64_t seed = 0)
{
    if (!B.solver) { std::printf(" generalized_lanczos: B has no solver set\n Terminating.\n"); std::exit(1); }
    std::vector<dp> q1;
    if (use_q1) q1.assign(Q, Q + A.nrow);
    sigb_check(sigb_generalized_lanczos(A.device_handle(), B.device_handle(), B.solver->dev, B.pc ? B.pc->dev : nullptr, n,
                                        use_q1 ? q1.data() : nullptr, seed, T, Q));
}
inline void generalized_eigensolve(linear_operator &A, linear_operator &B, int n, dp *lambda, dp *V,
                                   bool use_q1 = false, uint64_t seed = 0)
{
    if (!B.solver) { std::printf(" generalized_eigensolve: B has no solver set\n Terminating.\n"); std::exit(1); }
    std::vector<dp> q1;
    if (use_q1) q1.assign(V, V + A.nrow);
    sigb_check(sigb_generalized_eigensolve(A.device_handle(), B.device_handle(), B.solver->dev, B.pc ? B.pc->dev : nullptr, n,
                                           use_q1 ? q1.data() : nullptr, seed, lambda, V));
}

}  // namespace sigma
