// operators.cu -- operator expressions over device matrices (SURVEY.md 8f
// rank 2): the callers one level above the leaf matvec.
//
// Replaces the bodies of
//   operator_sum_matvec_add / _t_add        src/linear_operator/linear_operator_sums.f90:100-131
//   operator_product_matvec_add / _t_add    src/linear_operator/linear_operator_products.f90:78-150
//   operator_adjoint_matvec_add / _t_add    src/linear_operator/linear_operator_adjoints.f90:62-86
//   composite_matvec_add / _t_add           src/matrix/sparse_matrix_composites.f90:1076-1129
//   operator_sum_get_value, operator_adjoint_get_value, composite_mat_get_value
//                                           (diagonal only: what jacobi_setup reads)
// An expression is itself a sigb_matrix_t (every sparse matrix IS a
// linear_operator in the reference), so the solvers, the Lanczos loops and the
// matvec entry points take it unchanged.  All intermediate vectors stay on the
// device; every leaf runs the same SpMV kernels as a plain matvec, with the
// reference's accumulation order:
//   * `matvec` of a non-leaf is `y = 0 ; matvec_add` in the reference
//     (linear_operator_interface.f90:185-194).  The zero-fill pass is fused
//     away by letting the FIRST contribution to each y range overwrite
//     (MODE_SET) -- bit-identical, because a row sum that starts from +0.0 can
//     never be -0.0, so 0.0 + z == z exactly.
//   * operator_product keeps its two ping-pong scratch vectors z1 / z2
//     (linear_operator_products.f90:15,61) on the device.
#include <algorithm>

#include "dist.h"
#include "krylov.cuh"
#include "solvers.h"

namespace sigb {

namespace {

// y = y + z   (operator_product_matvec_add :109 / :146)
struct AddOp {
    static constexpr int ND = 0;
    static constexpr int NIN = 2;
    double *__restrict__ y;
    const double *__restrict__ z;
    const int *skip;
    __device__ bool begin() { return !(skip && *skip != 0); }
    __device__ void load(int64_t i, double *in) { in[0] = y[i]; in[1] = z[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { y[i] = add(in[0], in[1]); }
    __device__ double *out(int) { return nullptr; }
};

// what the fused SpMV epilogue does for a leaf (spmv_device.cuh emit_row), run
// as a separate pass behind an expression: y = scale * y ; u.y ; y.y
template <int NDOT>
struct PostOp {
    static constexpr int ND = NDOT;
    static constexpr int NIN = 3;
    double *__restrict__ y;
    const double *__restrict__ u, *__restrict__ scale;
    double *o0, *o1;
    const int *skip;
    __device__ bool begin() { return !(skip && *skip != 0); }
    __device__ void load(int64_t i, double *in)
    {
        in[0] = y[i];
        if (NDOT >= 1) in[1] = u[i];
        if (scale) in[2] = scale[i];
    }
    __device__ void compute(int64_t i, const double *in, double *acc)
    {
        double z = in[0];
        if (scale) { z = mul(in[2], z); y[i] = z; }
        if (NDOT >= 1) acc[0] = add(acc[0], mul(in[1], z));
        if (NDOT >= 2) acc[NDOT > 1 ? 1 : 0] = add(acc[NDOT > 1 ? 1 : 0], mul(z, z));
    }
    __device__ double *out(int d) { return d == 0 ? o0 : o1; }
};

// d = 0   /   idiag = 1 / d   (jacobi_solvers.f90:57-59)
struct FillOp {
    static constexpr int ND = 0;
    static constexpr int NIN = 0;
    double *__restrict__ y;
    double v;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t, double *) {}
    __device__ void compute(int64_t i, const double *, double *) { y[i] = v; }
    __device__ double *out(int) { return nullptr; }
};
struct RecipOp {
    static constexpr int ND = 0;
    static constexpr int NIN = 1;
    double *__restrict__ y;
    __device__ bool begin() { return true; }
    __device__ void load(int64_t i, double *in) { in[0] = y[i]; }
    __device__ void compute(int64_t i, const double *in, double *) { y[i] = 1.0 / in[0]; }
    __device__ double *out(int) { return nullptr; }
};

// out[g] (+)= A%get_value(g - roff, g - coff) for g in [lo, hi), all 1-based,
// on one leaf.  `line` is the stored line that holds the entry and `want` the
// id looked for in it: (row, column) for a csr_matrix / ellpack_matrix,
// (column, row) for a csc_matrix.  Last hit wins, 0 when absent
// (cs_matrices.f90:709-744, ellpack_matrices.f90:220-237).
__global__ void __launch_bounds__(kThreads)
diag_cs_kernel(const int32_t *__restrict__ ptr1, const int32_t *__restrict__ node1,
               const double *__restrict__ val, int64_t lo, int64_t hi, int64_t line_off, int64_t want_off,
               double *__restrict__ out, int accumulate)
{
    for (int64_t g = lo + blockIdx.x * (int64_t)kThreads + threadIdx.x; g < hi;
         g += (int64_t)gridDim.x * kThreads) {
        const int64_t line = g - line_off, want = g - want_off;
        double z = 0.0;
        for (int32_t k = ptr1[line - 1] - 1; k < ptr1[line] - 1; k++)
            if (node1[k] == want) z = val[k];
        out[g - 1] = accumulate ? add(out[g - 1], z) : z;
    }
}

__global__ void __launch_bounds__(kThreads)
diag_ell_kernel(const int32_t *__restrict__ node_sm, const double *__restrict__ val_sm,
                const int32_t *__restrict__ degrees, int32_t n_pad, int64_t lo, int64_t hi, int64_t roff,
                int64_t coff, double *__restrict__ out, int accumulate)
{
    for (int64_t g = lo + blockIdx.x * (int64_t)kThreads + threadIdx.x; g < hi;
         g += (int64_t)gridDim.x * kThreads) {
        const int64_t i = g - roff, j = g - coff;
        double z = 0.0;
        const int32_t d = degrees[i - 1];
        for (int32_t k = 0; k < d; k++)
            if (node_sm[(size_t)k * n_pad + (i - 1)] == j) z = val_sm[(size_t)k * n_pad + (i - 1)];
        out[g - 1] = accumulate ? add(out[g - 1], z) : z;
    }
}

// y_clean: no y(i) is -0.0.  True of any y an expression has already written
// (a row sum that starts from +0.0 is never -0.0, and y + z is -0.0 only when
// both are), which lets the later contributions skip tiles without entries.
int apply(sigb_matrix_t A, int trans, const double *x, double *y, bool add_to_y, const int *skip, bool y_clean);

int apply_product(sigb_matrix_t A, int trans, const double *x, double *y, bool add_to_y, const int *skip)
{
    OpInfo *op = A->op;
    const int nf = (int)op->kids.size();
    const double *src = x;
    for (int step = 0; step < nf; step++) {
        // last factor first (:97), or first factor first, transposed (:132)
        sigb_matrix_t P = op->kids[(size_t)(trans ? step : nf - 1 - step)];
        const bool last = (step == nf - 1);
        double *dst = (last && !add_to_y) ? y : ((step & 1) ? op->z2 : op->z1);
        SIGB_CHECK(apply(P, trans, src, dst, false, skip, false));
        src = dst;
    }
    if (add_to_y) {
        AddOp a{y, src, skip};
        SIGB_CHECK(launch_ew(a, trans ? A->ncol : A->nrow));
    }
    return SIGB_OK;
}

int apply(sigb_matrix_t A, int trans, const double *x, double *y, bool add_to_y, const int *skip, bool y_clean)
{
    OpInfo *op = A->op;
    if (!op) {
        DotSpec d;
        d.skip_flag = skip;
        d.y_no_negative_zero = add_to_y && y_clean;
        return matvec_dev(A, trans, x, y, MODE_SET, add_to_y, d);
    }
    switch (op->kind) {
    case OP_SUM:
        for (size_t k = 0; k < op->kids.size(); k++)
            SIGB_CHECK(apply(op->kids[k], trans, x, y, add_to_y || k > 0, skip, y_clean || k > 0));
        return SIGB_OK;
    case OP_ADJOINT:
        return apply(op->kids[0], !trans, x, y, add_to_y, skip, y_clean);
    case OP_PRODUCT:
        return apply_product(A, trans, x, y, add_to_y, skip);
    default: {
        const int nr = op->num_row_mats, nc = op->num_col_mats;
        if (!trans) {
            for (int it = 0; it < nr; it++)          // :1086
                for (int jt = 0; jt < nc; jt++)      // :1090
                    SIGB_CHECK(apply(op->kids[(size_t)it * nc + jt], 0, x + (op->col_ptr[(size_t)jt] - 1),
                                     y + (op->row_ptr[(size_t)it] - 1), add_to_y || jt > 0, skip, y_clean || jt > 0));
        } else {
            for (int jt = 0; jt < nc; jt++)          // :1115
                for (int it = 0; it < nr; it++)      // :1119
                    SIGB_CHECK(apply(op->kids[(size_t)it * nc + jt], 1, x + (op->row_ptr[(size_t)it] - 1),
                                     y + (op->col_ptr[(size_t)jt] - 1), add_to_y || it > 0, skip, y_clean || it > 0));
        }
        return SIGB_OK;
    }
    }
}

// out[g - 1] (+)= A%get_value(g - roff, g - coff), g in [lo, hi)
int diag_range(sigb_matrix_t A, int64_t roff, int64_t coff, int64_t lo, int64_t hi, double *out, bool accumulate)
{
    if (hi <= lo) return SIGB_OK;
    cudaStream_t st = ctx().stream;
    const int grid = ew_grid(hi - lo);
    OpInfo *op = A->op;
    if (!op) {
        sigb_graph_t g = A->g;
        SIGB_REQUIRE(!A->dist, SIGB_ERR_UNSUPPORTED, "get_value on a row-sharded operator inside an expression");
        if (g->kind == G_ELL)
            diag_ell_kernel<<<grid, kThreads, 0, st>>>(g->ell_node, A->val, g->ell_degrees, g->n_pad, lo, hi, roff,
                                                       coff, out, accumulate ? 1 : 0);
        else if (g->kind == G_CSR)
            diag_cs_kernel<<<grid, kThreads, 0, st>>>(g->stored.ptr, g->stored.node, A->val, lo, hi, roff, coff, out,
                                                      accumulate ? 1 : 0);
        else
            diag_cs_kernel<<<grid, kThreads, 0, st>>>(g->stored.ptr, g->stored.node, A->val, lo, hi, coff, roff, out,
                                                      accumulate ? 1 : 0);
        count_launch();
        SIGB_CUDA(cudaGetLastError());
        return SIGB_OK;
    }
    switch (op->kind) {
    case OP_SUM: {
        // z = 0 ; do k: z = z + summands(k)%get_value(i, j)   (linear_operator_sums.f90:89-93)
        double *dst = out;
        double *tmp = nullptr;
        if (accumulate) {   // a sum nested in a sum: its own total first, then added
            SIGB_CUDA(cudaMalloc((void **)&tmp, sizeof(double) * (size_t)(hi - lo)));
            dst = tmp - (lo - 1);
        }
        FillOp f{dst + (lo - 1), 0.0};
        int rc = launch_ew(f, hi - lo);
        for (size_t k = 0; rc == SIGB_OK && k < op->kids.size(); k++)
            rc = diag_range(op->kids[k], roff, coff, lo, hi, dst, true);
        if (rc == SIGB_OK && accumulate) {
            AddOp a{out + (lo - 1), tmp, nullptr};
            rc = launch_ew(a, hi - lo);
        }
        if (tmp) {
            cudaStreamSynchronize(st);
            cudaFree(tmp);
        }
        return rc;
    }
    case OP_ADJOINT:   // z = A%op%get_value(j, i)   (linear_operator_adjoints.f90:55)
        return diag_range(op->kids[0], coff, roff, lo, hi, out, accumulate);
    case OP_COMPOSITE: {
        // owning row / column block, then the block's own get_value with local
        // indices (sparse_matrix_composites.f90:474-482)
        const int nr = op->num_row_mats, nc = op->num_col_mats;
        for (int it = 0; it < nr; it++)
            for (int jt = 0; jt < nc; jt++) {
                const int64_t i1 = op->row_ptr[(size_t)it], i2 = op->row_ptr[(size_t)it + 1];   // [i1, i2)
                const int64_t j1 = op->col_ptr[(size_t)jt], j2 = op->col_ptr[(size_t)jt + 1];
                const int64_t l = std::max({lo, roff + i1, coff + j1});
                const int64_t h = std::min({hi, roff + i2, coff + j2});
                SIGB_CHECK(diag_range(op->kids[(size_t)it * nc + jt], roff + i1 - 1, coff + j1 - 1, l, h, out,
                                      accumulate));
            }
        return SIGB_OK;
    }
    default:
        set_error("get_value of an operator_product is undefined in the reference (the default "
                  "linear_operator_get_value multiplies an uninitialised vector, "
                  "linear_operator_interface.f90:168-181); jacobi cannot be set up on it");
        return SIGB_ERR_UNSUPPORTED;
    }
}

}  // namespace

// y = op(A) x (or y += op(A) x) for an expression, followed by what the fused
// SpMV epilogue does for a leaf: optional row scaling and up to two dots.
int op_matvec(sigb_matrix_t A, int trans, const double *x, double *y, bool add_to_y, const DotSpec &dot)
{
    SIGB_CHECK(apply(A, trans, x, y, add_to_y, dot.skip_flag, false));
    const int64_t n = trans ? A->ncol : A->nrow;
    if (dot.ndot == 0 && !dot.row_scale) return SIGB_OK;
    SIGB_REQUIRE(!add_to_y || !dot.row_scale, SIGB_ERR_ARG, "row scaling needs the overwrite form");
    if (dot.ndot == 0) {
        PostOp<0> p{y, nullptr, dot.row_scale, nullptr, nullptr, dot.skip_flag};
        return launch_ew(p, n);
    }
    if (dot.ndot == 1) {
        PostOp<1> p{y, dot.u, dot.row_scale, dot.out[0], nullptr, dot.skip_flag};
        return launch_ew(p, n);
    }
    PostOp<2> p{y, dot.u, dot.row_scale, dot.out[0], dot.out[1], dot.skip_flag};
    return launch_ew(p, n);
}

// idiag(i) = 1 / A%get_value(i, i) for an expression (jacobi_solvers.f90:55-59)
int op_jacobi_setup(sigb_matrix_t A, double *idiag)
{
    const int64_t n = A->nrow;
    SIGB_CHECK(diag_range(A, 0, 0, 1, n + 1, idiag, false));
    RecipOp r{idiag};
    return launch_ew(r, n);
}

static sigb_matrix_t make_op(int kind, int32_t nrow, int32_t ncol)
{
    sigb_matrix_t A = new sigb_matrix_s();
    A->nrow = nrow;
    A->ncol = ncol;
    A->op = new OpInfo();
    A->op->kind = kind;
    return A;
}

// C%summands(k)%ap => A ; call A%add_reference()   (linear_operator_sums.f90:64-67)
static void adopt(sigb_matrix_t parent, sigb_matrix_t kid)
{
    kid->refcount++;
    parent->op->kids.push_back(kid);
}

void op_destroy(sigb_matrix_t A)
{
    OpInfo *op = A->op;
    if (!op) return;
    for (sigb_matrix_t k : op->kids) sigb_matrix_destroy(k);   // remove_reference, destroy at 0
    cudaFree(op->z1);
    cudaFree(op->z2);
    delete op;
    A->op = nullptr;
}

}  // namespace sigb

using namespace sigb;

extern "C" {

static int check_operand(sigb_matrix_t A, const char *who)
{
    SIGB_REQUIRE(A, SIGB_ERR_ARG, "%s: null operator", who);
    SIGB_REQUIRE(!A->dist, SIGB_ERR_UNSUPPORTED, "%s: row-sharded operators cannot be part of an expression", who);
    return SIGB_OK;
}

int sigb_operator_sum(sigb_matrix_t A, sigb_matrix_t B, sigb_matrix_t *C)
{
    if (A && A->mg) { ::sigb::set_error("sigb_operator_sum: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(C, SIGB_ERR_ARG, "sigb_operator_sum: null output");
    SIGB_CHECK(check_operand(A, "sigb_operator_sum"));
    SIGB_CHECK(check_operand(B, "sigb_operator_sum"));
    SIGB_REQUIRE(A->nrow == B->nrow && A->ncol == B->ncol, SIGB_ERR_ARG,
                 "Dimensions of operators to be summed are not consistent");
    sigb_matrix_t S = make_op(OP_SUM, A->nrow, A->ncol);
    adopt(S, A);
    adopt(S, B);
    *C = S;
    return SIGB_OK;
}

int sigb_operator_product(sigb_matrix_t A, sigb_matrix_t B, sigb_matrix_t *C)
{
    if (A && A->mg) { ::sigb::set_error("sigb_operator_product: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(C, SIGB_ERR_ARG, "sigb_operator_product: null output");
    SIGB_CHECK(check_operand(A, "sigb_operator_product"));
    SIGB_CHECK(check_operand(B, "sigb_operator_product"));
    SIGB_REQUIRE(A->ncol == B->nrow, SIGB_ERR_ARG, "Dimensions of operators to be multiplied are inconsistent");
    sigb_matrix_t P = make_op(OP_PRODUCT, A->nrow, B->ncol);
    adopt(P, A);
    adopt(P, B);
    // temp_vec_size = maxval([A%nrow, A%ncol, B%nrow, B%ncol])   (:60)
    OpInfo *op = P->op;
    op->temp_vec_size = std::max({A->nrow, A->ncol, B->nrow, B->ncol});
    const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(op->temp_vec_size, 1);
    cudaError_t e = cudaMalloc((void **)&op->z1, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&op->z2, bytes);
    if (e != cudaSuccess) {
        sigb_matrix_destroy(P);
        return cuda_fail(e, "operator_product scratch", __FILE__, __LINE__);
    }
    *C = P;
    return SIGB_OK;
}

int sigb_operator_adjoint(sigb_matrix_t A, sigb_matrix_t *B)
{
    if (A && A->mg) { ::sigb::set_error("sigb_operator_adjoint: not available for a single-process multi-GPU operator"); return SIGB_ERR_UNSUPPORTED; }
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(B, SIGB_ERR_ARG, "sigb_operator_adjoint: null output");
    SIGB_CHECK(check_operand(A, "sigb_operator_adjoint"));
    sigb_matrix_t T = make_op(OP_ADJOINT, A->ncol, A->nrow);
    adopt(T, A);
    *B = T;
    return SIGB_OK;
}

int sigb_composite_create(int32_t num_row_mats, int32_t num_col_mats, const int32_t *rows, const int32_t *cols,
                          const sigb_matrix_t *blocks, sigb_matrix_t *A_out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(A_out && rows && cols && blocks && num_row_mats >= 1 && num_col_mats >= 1, SIGB_ERR_ARG,
                 "sigb_composite_create: bad argument");
    int64_t nrow = 0, ncol = 0;
    for (int it = 0; it < num_row_mats; it++) {
        SIGB_REQUIRE(rows[it] >= 0, SIGB_ERR_ARG, "sigb_composite_create: negative block size");
        nrow += rows[it];
    }
    for (int jt = 0; jt < num_col_mats; jt++) {
        SIGB_REQUIRE(cols[jt] >= 0, SIGB_ERR_ARG, "sigb_composite_create: negative block size");
        ncol += cols[jt];
    }
    SIGB_REQUIRE(nrow <= INT32_MAX && ncol <= INT32_MAX, SIGB_ERR_ARG, "sigb_composite_create: dimensions overflow int32");
    for (int it = 0; it < num_row_mats; it++)
        for (int jt = 0; jt < num_col_mats; jt++) {
            sigb_matrix_t Bk = blocks[(size_t)it * num_col_mats + jt];
            SIGB_CHECK(check_operand(Bk, "sigb_composite_create"));
            // composite_mat_set_submatrix :1042-1052
            SIGB_REQUIRE(Bk->nrow == rows[it] && Bk->ncol == cols[jt], SIGB_ERR_ARG,
                         "Inconsistent dimensions for sub-matrix (%d, %d): block is %d x %d, slot is %d x %d", it + 1,
                         jt + 1, Bk->nrow, Bk->ncol, rows[it], cols[jt]);
        }
    sigb_matrix_t S = make_op(OP_COMPOSITE, (int32_t)nrow, (int32_t)ncol);
    OpInfo *op = S->op;
    op->num_row_mats = num_row_mats;
    op->num_col_mats = num_col_mats;
    // row_ptr(1) = 1 ; row_ptr(it + 1) = row_ptr(it) + rows(it)   (:248-257)
    op->row_ptr.assign((size_t)num_row_mats + 1, 1);
    op->col_ptr.assign((size_t)num_col_mats + 1, 1);
    for (int it = 0; it < num_row_mats; it++) op->row_ptr[(size_t)it + 1] = op->row_ptr[(size_t)it] + rows[it];
    for (int jt = 0; jt < num_col_mats; jt++) op->col_ptr[(size_t)jt + 1] = op->col_ptr[(size_t)jt] + cols[jt];
    for (int k = 0; k < num_row_mats * num_col_mats; k++) adopt(S, blocks[k]);
    *A_out = S;
    return SIGB_OK;
}

int sigb_matrix_retain(sigb_matrix_t A)
{
    SIGB_REQUIRE(A, SIGB_ERR_ARG, "sigb_matrix_retain: null operator");
    A->refcount++;
    return SIGB_OK;
}

}  // extern "C"
