/*
 * sigma_oracle.c -- CPU restatement of the SiGMA (danshapero/sigma) sparse
 * matvec + Krylov hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call it.  The product path (sigma_b200/) never links it.
 *
 * Parity status: PINNED on the two deterministic known-answer tests the
 * reference ships (test/solver_test_diffusion_1d.f90: 64 CG iterations, misfit
 * 0.0; test/solver_test_advection_diffusion_1d.f90: BiCGSTAB, misfit <= 1e-8)
 * -- see tests/test_oracle_kat.py.  The reference itself (Fortran 2003) cannot
 * be compiled in this image (no Fortran compiler), so there is no oracle/_ref.
 * The eigensolve() tail calls LAPACK dstev, which is not vendored by the
 * reference and is unpinned there ("blas lapack", src/CMakeLists.txt:41):
 * orc_tridiag_eig below is our own implicit-QL restatement => "parity
 * unpinned" for that one boundary.
 *
 * Arithmetic rules (what a default gfortran build of the reference does on
 * x86-64): IEEE fp64, strict left-to-right accumulation, no FMA contraction,
 * no reassociation.  Compile with
 *     gcc -O2 -ffp-contract=off -fno-fast-math
 * All index arrays are 1-based int32 exactly as the Fortran holds them; array
 * element k of the Fortran array is C element [k-1].
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* ll_graph: list-of-lists builder graph                                     */
/* ------------------------------------------------------------------------ */

/*
 * Replays a sequence of g%add_edge(i, j) calls on an ll_graph
 * (src/graph/formats/ll_graphs.f90:355-371: push j onto row i's list unless
 * already connected) and then dumps the edges in the order the ll_graph edge
 * iterator returns them (ll_get_edges, src/graph/formats/ll_graphs.f90:295-345:
 * row by row, each row in insertion order).
 *
 * in : n rows, ne_in add_edge calls (ei[k], ej[k]) 1-based
 * out: out_i/out_j (capacity ne_in) = iteration-order edge stream,
 *      returns the number of distinct edges (g%ne).
 */
ORC_API int64_t orc_ll_graph_edges(int32_t n, int64_t ne_in, const int32_t *ei,
                                   const int32_t *ej, int32_t *out_i,
                                   int32_t *out_j, int32_t *max_d_out)
{
    int64_t *cnt = calloc((size_t)n + 2, sizeof(int64_t));
    int64_t *start = calloc((size_t)n + 2, sizeof(int64_t));
    int32_t *len = calloc((size_t)n + 1, sizeof(int32_t));
    int32_t *slots = malloc((size_t)(ne_in > 0 ? ne_in : 1) * sizeof(int32_t));
    int64_t k, ne = 0;
    int32_t i, max_d = 0;

    /* upper bound on each row's list length = number of add_edge calls */
    for (k = 0; k < ne_in; k++) cnt[ei[k]]++;
    start[1] = 0;
    for (i = 1; i <= n; i++) start[i + 1] = start[i] + cnt[i];

    for (k = 0; k < ne_in; k++) {
        int32_t r = ei[k], c = ej[k], l, found = 0;
        int32_t *row = slots + start[r];
        for (l = 0; l < len[r]; l++)               /* ll_connected */
            if (row[l] == c) { found = 1; break; }
        if (!found) {
            row[len[r]++] = c;                     /* lists(i)%push(j) */
            if (len[r] > max_d) max_d = len[r];
            ne++;
        }
    }

    k = 0;
    for (i = 1; i <= n; i++) {
        int32_t l;
        for (l = 0; l < len[i]; l++) {
            out_i[k] = i;
            out_j[k] = slots[start[i] + l];
            k++;
        }
    }
    if (max_d_out) *max_d_out = max_d;
    free(cnt); free(start); free(len); free(slots);
    return ne;
}

/* ------------------------------------------------------------------------ */
/* cs_graph                                                                  */
/* ------------------------------------------------------------------------ */

/*
 * cs_graph_build, src/graph/formats/cs_graphs.f90:109-197, fed by an edge
 * stream in the source graph's iteration order (copy_graph,
 * src/graph/graph_interfaces.f90:276-318; the 64-edge batching at :267 does
 * not change the order).  Pass 1 counts edges per row and prefix-sums into a
 * 1-based ptr (:144-156); pass 2 inserts each edge into the first free slot of
 * its row, skipping duplicates (:163-183).  trans swaps the roles of the two
 * endpoints (:122-125).
 *
 * ptr has n+1 entries, node has ne entries (ne = length of the stream).
 * Returns max_d (:191-194).  If the stream holds duplicates, unfilled slots
 * are left 0 (the reference would then prune them, :186-189; our callers never
 * pass duplicates, and the return value is negated to flag it).
 */
ORC_API int32_t orc_cs_graph_build(int32_t n, int64_t ne, const int32_t *src_i,
                                   const int32_t *src_j, int32_t trans,
                                   int32_t *ptr, int32_t *node)
{
    const int32_t *e1 = trans ? src_j : src_i;
    const int32_t *e2 = trans ? src_i : src_j;
    int64_t k;
    int32_t i, max_d = 0, holes = 0;

    for (i = 0; i <= n; i++) ptr[i] = 0;
    for (k = 0; k < ne; k++) ptr[e1[k]] += 1;      /* g%ptr(i + 1) += 1 */
    ptr[0] = 1;
    for (i = 1; i <= n; i++) ptr[i] = ptr[i] + ptr[i - 1];

    for (k = 0; k < ne; k++) node[k] = 0;
    for (k = 0; k < ne; k++) {
        int32_t r = e1[k], c = e2[k], l;
        for (l = ptr[r - 1]; l <= ptr[r] - 1; l++) {
            if (node[l - 1] == c) break;
            if (node[l - 1] == 0) { node[l - 1] = c; break; }
        }
    }
    for (k = 0; k < ne; k++) if (node[k] == 0) { holes = 1; break; }

    for (i = 1; i <= n; i++) {
        int32_t d = ptr[i] - ptr[i - 1];
        if (d > max_d) max_d = d;
    }
    return holes ? -max_d : max_d;
}

/* ------------------------------------------------------------------------ */
/* ellpack_graph                                                             */
/* ------------------------------------------------------------------------ */

/*
 * First pass of ellpack_graph_build,
 * src/graph/formats/ellpack_graphs.f90:128-141: per-row edge counts and their
 * maximum (g%max_d), which sizes node(max_d, n).
 */
ORC_API int32_t orc_ellpack_max_degree(int32_t n, int64_t ne,
                                       const int32_t *src_i,
                                       const int32_t *src_j, int32_t trans)
{
    const int32_t *e1 = trans ? src_j : src_i;
    int32_t *deg = calloc((size_t)n + 1, sizeof(int32_t));
    int32_t i, max_d = 0;
    int64_t k;
    for (k = 0; k < ne; k++) deg[e1[k]]++;
    for (i = 1; i <= n; i++) if (deg[i] > max_d) max_d = deg[i];
    free(deg);
    return max_d;
}

/*
 * Second pass of ellpack_graph_build,
 * src/graph/formats/ellpack_graphs.f90:143-168.  node is the Fortran array
 * node(max_d, n) in column-major order, i.e. node[(i-1)*max_d + (k-1)] is slot
 * k of row i.  Every insertion writes j into ALL remaining slots of the row
 * (g%node(d+1:, i) = j, :164), so padding slots end up holding a copy of the
 * row's last neighbour; rows with no edge keep node = 0 (:146).
 */
ORC_API void orc_ellpack_graph_build(int32_t n, int64_t ne,
                                     const int32_t *src_i,
                                     const int32_t *src_j, int32_t trans,
                                     int32_t max_d, int32_t *node,
                                     int32_t *degrees)
{
    const int32_t *e1 = trans ? src_j : src_i;
    const int32_t *e2 = trans ? src_i : src_j;
    int64_t k, tot = (int64_t)max_d * n;
    int32_t i;
    for (k = 0; k < tot; k++) node[k] = 0;
    for (i = 0; i < n; i++) degrees[i] = 0;
    for (k = 0; k < ne; k++) {
        int32_t r = e1[k], c = e2[k], l, d, connected = 0;
        int32_t *row = node + (int64_t)(r - 1) * max_d;
        d = degrees[r - 1];
        for (l = 0; l < d; l++)                    /* ellpack_connected :236 */
            if (row[l] == c) { connected = 1; break; }
        if (!connected) {
            for (l = d; l < max_d; l++) row[l] = c;
            degrees[r - 1] = d + 1;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* matrix entry mutators / accessors                                         */
/* ------------------------------------------------------------------------ */

/* csr_matrix_set_value / add_value, src/matrix/formats/cs_matrices.f90:840-891
 * (existing-entry branch only; returns 0 if (i,j) is not in the pattern, where
 * the reference would take the reallocation path).  For a csc_matrix call it
 * with (j, i) -- cs_matrices.f90:896-947 scans column j for row i. */
ORC_API int32_t orc_cs_set_value(const int32_t *ptr, const int32_t *node,
                                 double *val, int32_t i, int32_t j, double z,
                                 int32_t add)
{
    int32_t k, found = 0;
    for (k = ptr[i - 1]; k <= ptr[i] - 1; k++) {
        if (node[k - 1] == j) {
            val[k - 1] = add ? val[k - 1] + z : z;
            found = 1;
        }
    }
    return found;
}

/* csr_matrix_get_value, src/matrix/formats/cs_matrices.f90:709-724: no early
 * exit, 0 when absent. */
ORC_API double orc_cs_get_value(const int32_t *ptr, const int32_t *node,
                                const double *val, int32_t i, int32_t j)
{
    double z = 0.0;
    int32_t k;
    for (k = ptr[i - 1]; k <= ptr[i] - 1; k++)
        if (node[k - 1] == j) z = val[k - 1];
    return z;
}

/* ellpack_matrix_set_value / add_value,
 * src/matrix/formats/ellpack_matrices.f90:444-493: scans the first degrees(i)
 * slots only, so padding slots keep val = 0. */
ORC_API int32_t orc_ell_set_value(int32_t max_d, const int32_t *node,
                                  const int32_t *degrees, double *val,
                                  int32_t i, int32_t j, double z, int32_t add)
{
    int32_t k, d = degrees[i - 1], found = 0;
    int64_t base = (int64_t)(i - 1) * max_d;
    for (k = 0; k < d; k++) {
        if (node[base + k] == j) {
            val[base + k] = add ? val[base + k] + z : z;
            found = 1;
        }
    }
    return found;
}

/* ellpack_matrix_get_value, src/matrix/formats/ellpack_matrices.f90:220-237 */
ORC_API double orc_ell_get_value(int32_t max_d, const int32_t *node,
                                 const int32_t *degrees, const double *val,
                                 int32_t i, int32_t j)
{
    double z = 0.0;
    int32_t k, d = degrees[i - 1];
    int64_t base = (int64_t)(i - 1) * max_d;
    for (k = 0; k < d; k++)
        if (node[base + k] == j) z = val[base + k];
    return z;
}

/* ------------------------------------------------------------------------ */
/* matrix copy / format conversion                                           */
/* ------------------------------------------------------------------------ */

/*
 * copy_matrix_values, src/matrix/sparse_matrix_interfaces.f90:740-772: walk
 * the entries (si[k], sj[k], sv[k]) of the source in its iteration order and
 * `call A%set(i, j, z)` on the target, with (i, j) swapped when trans.  The
 * target graph was built beforehand from the same stream
 * (build_graph_from_matrix :692-735 -> orc_cs_graph_build /
 * orc_ellpack_graph_build).
 *
 * target_format: ORC_CSR (1): csr_matrix_set_value cs_matrices.f90:840-863
 *                ORC_CSC (2): csc_matrix_set_value :896-919 (scans column j)
 *                ORC_ELL (3): ellpack_matrix_set_value ellpack_matrices.f90:444-466
 * Returns the number of entries that were not found in the target pattern
 * (0 when the graph was built from the same stream).
 */
ORC_API int64_t orc_copy_matrix_values(int32_t target_format, int32_t max_d,
                                       const int32_t *ptr, const int32_t *node,
                                       const int32_t *degrees, double *val,
                                       int64_t ne, const int32_t *si,
                                       const int32_t *sj, const double *sv,
                                       int32_t trans)
{
    int64_t k, missing = 0;
    for (k = 0; k < ne; k++) {
        const int32_t i = trans ? sj[k] : si[k];
        const int32_t j = trans ? si[k] : sj[k];
        int32_t found;
        if (target_format == 1)      found = orc_cs_set_value(ptr, node, val, i, j, sv[k], 0);
        else if (target_format == 2) found = orc_cs_set_value(ptr, node, val, j, i, sv[k], 0);
        else                         found = orc_ell_set_value(max_d, node, degrees, val, i, j, sv[k], 0);
        if (!found) missing++;
    }
    return missing;
}

/*
 * A stream of `call A%add_value(i, j, z)` statements applied in order --
 * what an assembly loop does (examples/fem.f90:43-47) and what
 * add_multiple_values does over B(k, l) (csr_matrix_add_multiple_values
 * cs_matrices.f90:934-967).  Existing-entry branch of csr_matrix_add_value
 * (cs_matrices.f90:868-891), csc_matrix_add_value (:924-947, scans column j)
 * and ellpack_matrix_add_value (ellpack_matrices.f90:471-493): val = val + z.
 * Returns the number of calls whose (i, j) is not in the pattern (the reference
 * would take the reallocation path there; nothing is done for them here).
 */
ORC_API int64_t orc_add_values(int32_t format, int32_t max_d,
                               const int32_t *ptr, const int32_t *node,
                               const int32_t *degrees, double *val,
                               int64_t count, const int32_t *ci,
                               const int32_t *cj, const double *cz)
{
    int64_t c, missing = 0;
    for (c = 0; c < count; c++) {
        int32_t found;
        if (format == 1)      found = orc_cs_set_value(ptr, node, val, ci[c], cj[c], cz[c], 1);
        else if (format == 2) found = orc_cs_set_value(ptr, node, val, cj[c], ci[c], cz[c], 1);
        else                  found = orc_ell_set_value(max_d, node, degrees, val, ci[c], cj[c], cz[c], 1);
        if (!found) missing++;
    }
    return missing;
}

/* ------------------------------------------------------------------------ */
/* matvec kernels                                                            */
/* ------------------------------------------------------------------------ */

/* csr_matvec_add, src/matrix/formats/cs_matrices.f90:600-622 */
ORC_API void orc_csr_matvec_add(int32_t n, const int32_t *ptr,
                                const int32_t *node, const double *val,
                                const double *x, double *y)
{
    int32_t i, k;
    for (i = 1; i <= n; i++) {
        double z = 0.0;
        for (k = ptr[i - 1]; k <= ptr[i] - 1; k++) {
            int32_t j = node[k - 1];
            z = z + val[k - 1] * x[j - 1];
        }
        y[i - 1] = y[i - 1] + z;
    }
}

/* csc_matvec_add, src/matrix/formats/cs_matrices.f90:627-647 (also the
 * transpose kernel of a csr_matrix, :149) */
ORC_API void orc_csc_matvec_add(int32_t n, const int32_t *ptr,
                                const int32_t *node, const double *val,
                                const double *x, double *y)
{
    int32_t j, k;
    for (j = 1; j <= n; j++) {
        double z = x[j - 1];
        for (k = ptr[j - 1]; k <= ptr[j] - 1; k++) {
            int32_t i = node[k - 1];
            y[i - 1] = y[i - 1] + val[k - 1] * z;
        }
    }
}

/* ellpack_matvec_add, src/matrix/formats/ellpack_matrices.f90:640-665: all
 * max_d slots are multiplied, padding included. */
ORC_API void orc_ellpack_matvec_add(int32_t n, int32_t max_d,
                                    const int32_t *node, const double *val,
                                    const double *x, double *y)
{
    int32_t i, k;
    for (i = 1; i <= n; i++) {
        double z = 0.0;
        int64_t base = (int64_t)(i - 1) * max_d;
        for (k = 0; k < max_d; k++) {
            int32_t j = node[base + k];
            z = z + val[base + k] * x[j - 1];
        }
        y[i - 1] = y[i - 1] + z;
    }
}

/* ellpack_matvec_t_add, src/matrix/formats/ellpack_matrices.f90:670-693 */
ORC_API void orc_ellpack_matvec_t_add(int32_t n, int32_t max_d,
                                      const int32_t *node, const double *val,
                                      const double *x, double *y)
{
    int32_t i, j, k;
    for (j = 1; j <= n; j++) {
        double z = x[j - 1];
        int64_t base = (int64_t)(j - 1) * max_d;
        for (k = 0; k < max_d; k++) {
            i = node[base + k];
            y[i - 1] = y[i - 1] + val[base + k] * z;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* linear_operator dispatch                                                  */
/* ------------------------------------------------------------------------ */

/* Leaf formats, then the operator expressions of
 * src/linear_operator/linear_operator_{sums,products,adjoints}.f90 and the
 * block composite of src/matrix/sparse_matrix_composites.f90. */
enum { ORC_CSR = 1, ORC_CSC = 2, ORC_ELL = 3,
       ORC_SUM = 4, ORC_PRODUCT = 5, ORC_ADJOINT = 6, ORC_COMPOSITE = 7 };

typedef struct orc_matrix_s {
    int32_t format;        /* ORC_CSR ... ORC_COMPOSITE                      */
    int32_t nrow, ncol;
    int32_t max_d;         /* ELL width                                      */
    const int32_t *ptr;    /* cs: n+1 (n = nrow for CSR, ncol for CSC)       */
    const int32_t *node;   /* cs: ne ; ell: max_d*nrow column-major          */
    const int32_t *degrees;/* ell only: degrees(nrow)                        */
    const double *val;
    /* ---- operator expressions (unused by the leaf formats) -------------- */
    int32_t nkids;         /* num_summands / num_products / 1 / #blocks      */
    int32_t num_row_mats, num_col_mats;   /* composite                      */
    int32_t temp_vec_size; /* product: maxval([A%nrow,A%ncol,B%nrow,B%ncol]) */
    const struct orc_matrix_s *const *kids;
                           /* sum: summands(:) ; product: products(:) ;
                              adjoint: op ; composite: sub_mats(it, jt) at
                              [(it-1)*num_col_mats + (jt-1)]                 */
    const int32_t *row_ptr;/* composite: row_ptr(num_row_mats+1), 1-based    */
    const int32_t *col_ptr;/* composite: col_ptr(num_col_mats+1), 1-based    */
    double *z1, *z2;       /* product scratch, temp_vec_size doubles each    */
} orc_matrix;

static void zero(double *y, int32_t n) { int32_t i; for (i = 0; i < n; i++) y[i] = 0.0; }

ORC_API void orc_matvec(const orc_matrix *A, int32_t trans, const double *x,
                        double *y);

/* matvec_add / matvec_t_add dispatch.
 * Leaves: cs_matvec_add / cs_matvec_t_add,
 * src/matrix/formats/cs_matrices.f90:500-521 with the slot bindings at
 * :148-149 (CSR) and :192-193 (CSC); ellpack binds directly
 * (src/matrix/formats/ellpack_matrices.f90:78-79).
 * Expressions: operator_sum_matvec_add / _t_add
 * (src/linear_operator/linear_operator_sums.f90:100-131),
 * operator_product_matvec_add / _t_add
 * (src/linear_operator/linear_operator_products.f90:78-150),
 * operator_adjoint_matvec_add / _t_add
 * (src/linear_operator/linear_operator_adjoints.f90:62-86),
 * composite_matvec_add / _t_add
 * (src/matrix/sparse_matrix_composites.f90:1076-1129). */
ORC_API void orc_matvec_add(const orc_matrix *A, int32_t trans,
                            const double *x, double *y)
{
    int32_t i, k, it, jt;
    switch (A->format) {
    case ORC_CSR:
        if (!trans) orc_csr_matvec_add(A->nrow, A->ptr, A->node, A->val, x, y);
        else        orc_csc_matvec_add(A->nrow, A->ptr, A->node, A->val, x, y);
        break;
    case ORC_CSC:
        if (!trans) orc_csc_matvec_add(A->ncol, A->ptr, A->node, A->val, x, y);
        else        orc_csr_matvec_add(A->ncol, A->ptr, A->node, A->val, x, y);
        break;
    case ORC_ELL:
        if (!trans) orc_ellpack_matvec_add(A->nrow, A->max_d, A->node, A->val, x, y);
        else        orc_ellpack_matvec_t_add(A->nrow, A->max_d, A->node, A->val, x, y);
        break;
    case ORC_SUM:
        /* do k = 1, num_summands: call summands(k)%ap%matvec_add(x, y)  :108-110 */
        for (k = 0; k < A->nkids; k++) orc_matvec_add(A->kids[k], trans, x, y);
        break;
    case ORC_ADJOINT:
        /* call A%op%matvec_t_add(x, y) / matvec_add(x, y)   :69, :82 */
        orc_matvec_add(A->kids[0], !trans, x, y);
        break;
    case ORC_PRODUCT: {
        double *z1 = A->z1, *z2 = A->z2;
        /* z1(1:A%ncol) = x(1:A%ncol) ; z2(:) = 0   :93-94 / :128-129.  (The
         * transposed form copies A%ncol entries of an x that has A%nrow of
         * them -- the same thing for the square operators the reference
         * tests; the input length is used here.) */
        const int32_t nin = trans ? A->nrow : A->ncol;
        const int32_t nout = trans ? A->ncol : A->nrow;
        for (i = 0; i < nin; i++) z1[i] = x[i];
        zero(z2, A->temp_vec_size);
        if (!trans) {
            /* last factor first  :97-108 */
            for (k = A->nkids - 1; k >= 0; k--) {
                const orc_matrix *P = A->kids[k];
                zero(z2, A->temp_vec_size);          /* matvec zeroes all of z2 (:187 of the interface) */
                orc_matvec_add(P, 0, z1, z2);
                for (i = 0; i < P->nrow; i++) z1[i] = z2[i];
            }
        } else {
            /* first factor first, transposed  :132-143 */
            for (k = 0; k < A->nkids; k++) {
                const orc_matrix *P = A->kids[k];
                zero(z2, A->temp_vec_size);
                orc_matvec_add(P, 1, z1, z2);
                for (i = 0; i < P->ncol; i++) z1[i] = z2[i];
            }
        }
        for (i = 0; i < nout; i++) y[i] = y[i] + z2[i];   /* y = y + z2  :109 / :146 */
        zero(z1, A->temp_vec_size);                       /* :111-112 */
        zero(z2, A->temp_vec_size);
        break;
    }
    default: /* ORC_COMPOSITE */
        if (!trans) {
            for (it = 1; it <= A->num_row_mats; it++) {              /* :1086 */
                const int32_t i1 = A->row_ptr[it - 1];
                for (jt = 1; jt <= A->num_col_mats; jt++) {          /* :1090 */
                    const int32_t j1 = A->col_ptr[jt - 1];
                    const orc_matrix *Cm = A->kids[(it - 1) * A->num_col_mats + (jt - 1)];
                    orc_matvec_add(Cm, 0, x + (j1 - 1), y + (i1 - 1));   /* :1096 */
                }
            }
        } else {
            for (jt = 1; jt <= A->num_col_mats; jt++) {              /* :1115 */
                const int32_t j1 = A->col_ptr[jt - 1];
                for (it = 1; it <= A->num_row_mats; it++) {          /* :1119 */
                    const int32_t i1 = A->row_ptr[it - 1];
                    const orc_matrix *Cm = A->kids[(it - 1) * A->num_col_mats + (jt - 1)];
                    orc_matvec_add(Cm, 1, x + (i1 - 1), y + (j1 - 1));   /* :1125 */
                }
            }
        }
    }
}

/* linear_operator_matvec / matvec_t,
 * src/linear_operator/linear_operator_interface.f90:185-208: y = 0 (a full
 * pass) then matvec_add. */
ORC_API void orc_matvec(const orc_matrix *A, int32_t trans, const double *x,
                        double *y)
{
    zero(y, trans ? A->ncol : A->nrow);
    orc_matvec_add(A, trans, x, y);
}

/* A%get_value(i, j) for any operator.
 * csr_matrix_get_value cs_matrices.f90:709-724, csc_matrix_get_value :729-744
 * (scans column j for row i), ellpack_matrix_get_value
 * ellpack_matrices.f90:220-237, operator_sum_get_value
 * linear_operator_sums.f90:79-95, operator_adjoint_get_value
 * linear_operator_adjoints.f90:49-57, composite_mat_get_value
 * sparse_matrix_composites.f90:465-485 with get_owning_row/column_matrix
 * :1235-1262.  operator_product has no override: the default
 * linear_operator_get_value (linear_operator_interface.f90:168-181) multiplies
 * a vector that is uninitialised except for x(j) = 1 -- restated here with the
 * unit vector it evidently intends. */
ORC_API double orc_get_value(const orc_matrix *A, int32_t i, int32_t j)
{
    double z = 0.0;
    int32_t k, it, jt;
    switch (A->format) {
    case ORC_CSR: return orc_cs_get_value(A->ptr, A->node, A->val, i, j);
    case ORC_CSC: return orc_cs_get_value(A->ptr, A->node, A->val, j, i);
    case ORC_ELL: return orc_ell_get_value(A->max_d, A->node, A->degrees, A->val, i, j);
    case ORC_SUM:
        for (k = 0; k < A->nkids; k++) z = z + orc_get_value(A->kids[k], i, j);
        return z;
    case ORC_ADJOINT: return orc_get_value(A->kids[0], j, i);
    case ORC_COMPOSITE:
        for (it = 1; it <= A->num_row_mats; it++)
            if (A->row_ptr[it - 1] <= i && A->row_ptr[it] > i) break;
        for (jt = 1; jt <= A->num_col_mats; jt++)
            if (A->col_ptr[jt - 1] <= j && A->col_ptr[jt] > j) break;
        return orc_get_value(A->kids[(it - 1) * A->num_col_mats + (jt - 1)],
                             i - A->row_ptr[it - 1] + 1, j - A->col_ptr[jt - 1] + 1);
    default: {
        double *x = calloc((size_t)A->ncol, sizeof(double));
        double *y = calloc((size_t)A->nrow, sizeof(double));
        x[j - 1] = 1.0;
        orc_matvec(A, 0, x, y);
        z = y[i - 1];
        free(x);
        free(y);
        return z;
    }
    }
}

/* dot_product(a, b) / sum(a * b) as a non-fast-math gfortran evaluates them:
 * one accumulator, ascending index. */
static double dot(const double *a, const double *b, int32_t n)
{
    double s = 0.0;
    int32_t i;
    for (i = 0; i < n; i++) s = s + a[i] * b[i];
    return s;
}

/* ------------------------------------------------------------------------ */
/* Jacobi                                                                    */
/* ------------------------------------------------------------------------ */

/* jacobi_setup, src/solver/jacobi_solvers.f90:37-63:
 * idiag(i) = 1 / A%get_value(i, i) */
ORC_API void orc_jacobi_setup(const orc_matrix *A, double *idiag)
{
    int32_t i;
    for (i = 1; i <= A->nrow; i++) idiag[i - 1] = 1.0 / orc_get_value(A, i, i);
}

/* jacobi_solve, src/solver/jacobi_solvers.f90:68-81: x = idiag * b */
ORC_API void orc_jacobi_solve(int32_t n, const double *idiag, double *x,
                              const double *b)
{
    int32_t i;
    for (i = 0; i < n; i++) x[i] = idiag[i] * b[i];
}

/* ------------------------------------------------------------------------ */
/* CG                                                                        */
/* ------------------------------------------------------------------------ */

/*
 * cg_solve, src/solver/cg_solvers.f90:116-150.  work = 4*n doubles (p,q,r,z,
 * allocated/zeroed by cg_setup :52-90).  Returns the number of iterations
 * performed by THIS call (the reference accumulates into solver%iterations,
 * :145).  max_iter < 0 means "no cap" like the reference; a non-negative cap
 * is our safety net and is reported through *capped.
 */
ORC_API int64_t orc_cg_solve(const orc_matrix *A, double *x, const double *b,
                             double tolerance, int64_t max_iter, double *work,
                             double *res2_out, int32_t *capped)
{
    const int32_t n = A->nrow;
    double *p = work, *q = work + n, *r = work + 2 * (int64_t)n;
    double alpha, beta, res2, dpr;
    int64_t it = 0;
    int32_t i;

    if (capped) *capped = 0;
    orc_matvec(A, 0, x, q);                              /* :128 */
    for (i = 0; i < n; i++) r[i] = b[i] - q[i];          /* :129 */
    for (i = 0; i < n; i++) p[i] = r[i];                 /* :130 */
    res2 = dot(r, r, n);                                 /* :131 */

    while (sqrt(res2) > tolerance) {                     /* :133 */
        if (max_iter >= 0 && it >= max_iter) { if (capped) *capped = 1; break; }
        orc_matvec(A, 0, p, q);                          /* :134 */
        dpr = dot(p, q, n);                              /* :135 */
        alpha = res2 / dpr;                              /* :136 */
        for (i = 0; i < n; i++) x[i] = x[i] + alpha * p[i];   /* :137 */
        for (i = 0; i < n; i++) r[i] = r[i] - alpha * q[i];   /* :138 */

        dpr = dot(r, r, n);                              /* :140 */
        beta = dpr / res2;                               /* :141 */
        for (i = 0; i < n; i++) p[i] = r[i] + beta * p[i];    /* :142 */
        res2 = dpr;                                      /* :143 */
        it++;                                            /* :145 */
    }
    if (res2_out) *res2_out = res2;
    return it;
}

/*
 * cg_solve_pc, src/solver/cg_solvers.f90:155-194 with pc = jacobi
 * (pc%solve(A,z,r) -> jacobi_solve, src/solver/jacobi_solvers.f90:77).
 * Note res2 = r.z is both the recurrence scalar and the stopping quantity.
 */
ORC_API int64_t orc_cg_solve_jacobi(const orc_matrix *A, double *x,
                                    const double *b, const double *idiag,
                                    double tolerance, int64_t max_iter,
                                    double *work, double *res2_out,
                                    int32_t *capped)
{
    const int32_t n = A->nrow;
    double *p = work, *q = work + n, *r = work + 2 * (int64_t)n,
           *z = work + 3 * (int64_t)n;
    double alpha, beta, res2, dpr;
    int64_t it = 0;
    int32_t i;

    if (capped) *capped = 0;
    for (i = 0; i < n; i++) z[i] = x[i];                 /* :167 */
    orc_matvec(A, 0, z, q);                              /* :168 */
    for (i = 0; i < n; i++) r[i] = b[i] - q[i];          /* :169 */
    orc_jacobi_solve(n, idiag, z, r);                    /* :170 */
    for (i = 0; i < n; i++) p[i] = z[i];                 /* :171 */
    res2 = dot(r, z, n);                                 /* :172 */

    while (sqrt(res2) > tolerance) {                     /* :174 */
        if (max_iter >= 0 && it >= max_iter) { if (capped) *capped = 1; break; }
        orc_matvec(A, 0, p, q);
        dpr = dot(p, q, n);
        alpha = res2 / dpr;
        for (i = 0; i < n; i++) x[i] = x[i] + alpha * p[i];
        for (i = 0; i < n; i++) r[i] = r[i] - alpha * q[i];

        orc_jacobi_solve(n, idiag, z, r);                /* :181 */

        dpr = dot(r, z, n);                              /* :183 */
        beta = dpr / res2;
        for (i = 0; i < n; i++) p[i] = z[i] + beta * p[i];    /* :185 */
        res2 = dpr;
        it++;
    }
    if (res2_out) *res2_out = res2;
    return it;
}

/* ------------------------------------------------------------------------ */
/* ILDU(0): incomplete A ~= (I + L) D (I + U)                                */
/* ------------------------------------------------------------------------ */

/*
 * sparse_static_pattern_ldu_factorization, src/solver/ldu_solvers.f90:275-387.
 * L and U are csr matrices holding the STRICT lower / upper parts (unit
 * diagonals are implied); their patterns come from
 * incomplete_ldu_sparsity_pattern (:396-441): ll_graph add_edge(i, j) calls in
 * A's iteration order, i.e. row i of L / U lists row i's lower / upper
 * neighbours in the order A's iterator produced them -- unsorted, and the
 * elimination below walks them in that stored order (:330-331), exactly as
 * restated here.
 *
 * in : (ai, aj, av)[ne] = A's entry stream in iteration order
 * out: Lval, D, Uval (patterns Lptr/Lnode, Uptr/Unode given, 1-based)
 */
static double ldu_get(const int32_t *ptr, const int32_t *node, const double *val, int32_t i, int32_t j)
{
    return orc_cs_get_value(ptr, node, val, i, j);
}

ORC_API void orc_ldu_factor(int32_t n, int64_t ne, const int32_t *ai,
                            const int32_t *aj, const double *av,
                            const int32_t *Lptr, const int32_t *Lnode, double *Lval,
                            const int32_t *Uptr, const int32_t *Unode, double *Uval,
                            double *D)
{
    int64_t e;
    int32_t i, ind1, ind2;
    for (e = 0; e < Lptr[n] - 1; e++) Lval[e] = 0.0;                 /* call L%zero()  :302 */
    for (e = 0; e < Uptr[n] - 1; e++) Uval[e] = 0.0;                 /* call U%zero()  :303 */
    for (i = 0; i < n; i++) D[i] = 0.0;                              /* :304 */
    for (e = 0; e < ne; e++) {                                       /* copy A into L, D, U  :307-324 */
        if (ai[e] > aj[e])      orc_cs_set_value(Lptr, Lnode, Lval, ai[e], aj[e], av[e], 0);
        else if (aj[e] > ai[e]) orc_cs_set_value(Uptr, Unode, Uval, ai[e], aj[e], av[e], 0);
        else                    D[ai[e] - 1] = av[e];
    }
    for (i = 1; i <= n; i++) {                                       /* :331 */
        const int32_t lb = Lptr[i - 1] - 1, dl = Lptr[i] - 1 - lb;
        const int32_t ub = Uptr[i - 1] - 1, du = Uptr[i] - 1 - ub;
        for (ind1 = 0; ind1 < dl; ind1++) {                          /* :339 */
            const int32_t k = Lnode[lb + ind1];
            double Lik = ldu_get(Lptr, Lnode, Lval, i, k);           /* :342 */
            const double Uki = ldu_get(Uptr, Unode, Uval, k, i);     /* :343 */
            orc_cs_set_value(Lptr, Lnode, Lval, i, k, Lik / D[k - 1], 0);   /* :345 */
            Lik = Lik / D[k - 1];                                    /* :346 */
            for (ind2 = 0; ind2 < dl; ind2++) {                      /* :350 */
                const int32_t j = Lnode[lb + ind2];
                if (j > k) {
                    const double Ukj = ldu_get(Uptr, Unode, Uval, k, j);        /* :355 */
                    orc_cs_set_value(Lptr, Lnode, Lval, i, j, -(Lik * D[k - 1] * Ukj), 1);   /* :356 */
                }
            }
            D[i - 1] = D[i - 1] - Lik * D[k - 1] * Uki;              /* :361 */
            for (ind2 = 0; ind2 < du; ind2++) {                      /* :364 */
                const int32_t j = Unode[ub + ind2];
                const double Ukj = ldu_get(Uptr, Unode, Uval, k, j);            /* :366 */
                orc_cs_set_value(Uptr, Unode, Uval, i, j, -(Lik * D[k - 1] * Ukj), 1);       /* :367 */
            }
        }
        for (ind2 = 0; ind2 < du; ind2++) {                          /* :373 */
            const int32_t k = Unode[ub + ind2];
            const double Uik = ldu_get(Uptr, Unode, Uval, i, k);
            orc_cs_set_value(Uptr, Unode, Uval, i, k, Uik / D[i - 1], 0);       /* :376 */
        }
    }
}

/* ldu_solve, src/solver/ldu_solvers.f90:160-176: x = b ; (I + L) x = x ;
 * x = x / D ; (I + U) x = x, with lower_triangular_solve (:208-235) and
 * upper_triangular_solve (:240-263): rows in order, each row's entries in
 * stored order, z = z - val * x(j). */
ORC_API void orc_ldu_solve(int32_t n, const int32_t *Lptr, const int32_t *Lnode, const double *Lval,
                           const int32_t *Uptr, const int32_t *Unode, const double *Uval,
                           const double *D, double *x, const double *b)
{
    int32_t i, k;
    for (i = 0; i < n; i++) x[i] = b[i];
    for (i = 1; i <= n; i++) {
        double z = x[i - 1];
        for (k = Lptr[i - 1]; k <= Lptr[i] - 1; k++) z = z - Lval[k - 1] * x[Lnode[k - 1] - 1];
        x[i - 1] = z;
    }
    for (i = 0; i < n; i++) x[i] = x[i] / D[i];
    for (i = n; i >= 1; i--) {
        double z = x[i - 1];
        for (k = Uptr[i - 1]; k <= Uptr[i] - 1; k++) z = z - Uval[k - 1] * x[Unode[k - 1] - 1];
        x[i - 1] = z;
    }
}

/* cg_solve_pc, src/solver/cg_solvers.f90:155-194 with pc = sparse_ldu_solver. */
ORC_API int64_t orc_cg_solve_ldu(const orc_matrix *A, double *x, const double *b,
                                 const int32_t *Lptr, const int32_t *Lnode, const double *Lval,
                                 const int32_t *Uptr, const int32_t *Unode, const double *Uval,
                                 const double *D, double tolerance, int64_t max_iter,
                                 double *work, double *res2_out, int32_t *capped)
{
    const int32_t n = A->nrow;
    double *p = work, *q = work + n, *r = work + 2 * (int64_t)n, *z = work + 3 * (int64_t)n;
    double alpha, beta, res2, dpr;
    int64_t it = 0;
    int32_t i;

    if (capped) *capped = 0;
    for (i = 0; i < n; i++) z[i] = x[i];                 /* :167 */
    orc_matvec(A, 0, z, q);                              /* :168 */
    for (i = 0; i < n; i++) r[i] = b[i] - q[i];          /* :169 */
    orc_ldu_solve(n, Lptr, Lnode, Lval, Uptr, Unode, Uval, D, z, r);   /* :170 */
    for (i = 0; i < n; i++) p[i] = z[i];                 /* :171 */
    res2 = dot(r, z, n);                                 /* :172 */
    while (sqrt(res2) > tolerance) {                     /* :174 */
        if (max_iter >= 0 && it >= max_iter) { if (capped) *capped = 1; break; }
        orc_matvec(A, 0, p, q);
        dpr = dot(p, q, n);
        alpha = res2 / dpr;
        for (i = 0; i < n; i++) x[i] = x[i] + alpha * p[i];
        for (i = 0; i < n; i++) r[i] = r[i] - alpha * q[i];
        orc_ldu_solve(n, Lptr, Lnode, Lval, Uptr, Unode, Uval, D, z, r);   /* :181 */
        dpr = dot(r, z, n);
        beta = dpr / res2;
        for (i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        res2 = dpr;
        it++;
    }
    if (res2_out) *res2_out = res2;
    return it;
}

/* ------------------------------------------------------------------------ */
/* BiCGSTAB                                                                  */
/* ------------------------------------------------------------------------ */

/*
 * bicgstab_solve, src/solver/bicgstab_solvers.f90:124-177.
 * work = 8*n doubles (p,q,r,r0,v,s,t,z; bicgstab_setup :52-100).
 */
ORC_API int64_t orc_bicgstab_solve(const orc_matrix *A, double *x,
                                   const double *b, double tolerance,
                                   int64_t max_iter, double *work,
                                   double *res2_out, int32_t *capped)
{
    const int64_t n = A->nrow;
    double *p = work, *q = work + n, *r = work + 2 * n, *r0 = work + 3 * n,
           *v = work + 4 * n, *s = work + 5 * n, *t = work + 6 * n;
    double rho, rho_old, alpha, omega, beta, res2;
    int64_t it = 0;
    int32_t i;

    if (capped) *capped = 0;
    orc_matvec(A, 0, x, q);                              /* :140 */
    for (i = 0; i < n; i++) r0[i] = b[i] - q[i];         /* :141 */
    for (i = 0; i < n; i++) r[i] = r0[i];                /* :142 */
    rho = 1.0; rho_old = 1.0; alpha = 1.0; omega = 1.0;  /* :144-147 */
    for (i = 0; i < n; i++) v[i] = 0.0;                  /* :149 */
    for (i = 0; i < n; i++) p[i] = 0.0;                  /* :150 */
    res2 = dot(r, r, (int32_t)n);                        /* :152 */

    while (sqrt(res2) > tolerance) {                     /* :154 */
        if (max_iter >= 0 && it >= max_iter) { if (capped) *capped = 1; break; }
        rho = dot(r0, r, (int32_t)n);                    /* :155 */
        beta = rho / rho_old * alpha / omega;            /* :156 */
        for (i = 0; i < n; i++)                          /* :157 */
            p[i] = r[i] + beta * (p[i] - omega * v[i]);

        orc_matvec(A, 0, p, v);                          /* :159 */
        alpha = rho / dot(r0, v, (int32_t)n);            /* :160 */
        for (i = 0; i < n; i++) s[i] = r[i] - alpha * v[i];   /* :161 */

        orc_matvec(A, 0, s, t);                          /* :163 */
        omega = dot(s, t, (int32_t)n) / dot(t, t, (int32_t)n);  /* :164 */
        if (isnan(omega)) omega = 0.0;                   /* :165 */
        for (i = 0; i < n; i++)                          /* :166 */
            x[i] = x[i] + alpha * p[i] + omega * s[i];
        for (i = 0; i < n; i++) r[i] = s[i] - omega * t[i];   /* :167 */

        res2 = dot(r, r, (int32_t)n);                    /* :169 */
        rho_old = rho;                                   /* :170 */
        it++;                                            /* :172 */
    }
    if (res2_out) *res2_out = res2;
    return it;
}

/*
 * bicgstab_solve_pc, src/solver/bicgstab_solvers.f90:182-237 with pc = jacobi.
 * Left preconditioning, stops on the preconditioned residual, NO isnan guard.
 */
ORC_API int64_t orc_bicgstab_solve_jacobi(const orc_matrix *A, double *x,
                                          const double *b, const double *idiag,
                                          double tolerance, int64_t max_iter,
                                          double *work, double *res2_out,
                                          int32_t *capped)
{
    const int64_t n = A->nrow;
    double *p = work, *q = work + n, *r = work + 2 * n, *r0 = work + 3 * n,
           *v = work + 4 * n, *s = work + 5 * n, *t = work + 6 * n,
           *z = work + 7 * n;
    double rho, rho_old, alpha, omega, beta, res2;
    int64_t it = 0;
    int32_t i, nn = (int32_t)n;

    if (capped) *capped = 0;
    orc_matvec(A, 0, x, q);                              /* :199 */
    for (i = 0; i < n; i++) z[i] = b[i] - q[i];          /* :200 */
    orc_jacobi_solve(nn, idiag, r0, z);                  /* :201 */
    for (i = 0; i < n; i++) r[i] = r0[i];                /* :202 */
    rho = 1.0; rho_old = 1.0; alpha = 1.0; omega = 1.0;
    for (i = 0; i < n; i++) v[i] = 0.0;
    for (i = 0; i < n; i++) p[i] = 0.0;
    res2 = dot(r, r, nn);                                /* :212 */

    while (sqrt(res2) > tolerance) {                     /* :214 */
        if (max_iter >= 0 && it >= max_iter) { if (capped) *capped = 1; break; }
        rho = dot(r0, r, nn);                            /* :215 */
        beta = rho / rho_old * alpha / omega;            /* :216 */
        for (i = 0; i < n; i++)                          /* :217 */
            p[i] = r[i] + beta * (p[i] - omega * v[i]);
        orc_matvec(A, 0, p, z);                          /* :218 */
        orc_jacobi_solve(nn, idiag, v, z);               /* :219 */

        alpha = rho / dot(r0, v, nn);                    /* :221 */
        for (i = 0; i < n; i++) s[i] = r[i] - alpha * v[i];   /* :222 */
        orc_matvec(A, 0, s, z);                          /* :223 */
        orc_jacobi_solve(nn, idiag, t, z);               /* :224 */
        omega = dot(s, t, nn) / dot(t, t, nn);           /* :225 */
        for (i = 0; i < n; i++)                          /* :226 */
            x[i] = x[i] + alpha * p[i] + omega * s[i];
        for (i = 0; i < n; i++) r[i] = s[i] - omega * t[i];   /* :227 */

        rho_old = rho;                                   /* :229 */
        res2 = dot(r, r, nn);                            /* :230 */
        it++;
    }
    if (res2_out) *res2_out = res2;
    return it;
}

/*
 * bicgstab_solve_pc, src/solver/bicgstab_solvers.f90:182-237 with pc = ldu()
 * (call pc%solve(A, ., z) = ldu_solve, src/solver/ldu_solvers.f90:160-176).
 * The reference's tests never pair the two, but linear_solve_pc takes any
 * class(linear_solver) as pc (linear_operator_interface.f90:238-254), so a
 * caller may.  Same statements as the jacobi form above, another pc%solve.
 */
ORC_API int64_t orc_bicgstab_solve_ldu(const orc_matrix *A, double *x, const double *b,
                                       const int32_t *Lptr, const int32_t *Lnode, const double *Lval,
                                       const int32_t *Uptr, const int32_t *Unode, const double *Uval,
                                       const double *D, double tolerance, int64_t max_iter,
                                       double *work, double *res2_out, int32_t *capped)
{
    const int64_t n = A->nrow;
    double *p = work, *q = work + n, *r = work + 2 * n, *r0 = work + 3 * n,
           *v = work + 4 * n, *s = work + 5 * n, *t = work + 6 * n,
           *z = work + 7 * n;
    double rho, rho_old, alpha, omega, beta, res2;
    int64_t it = 0;
    int32_t i, nn = (int32_t)n;

    if (capped) *capped = 0;
    orc_matvec(A, 0, x, q);                              /* :199 */
    for (i = 0; i < n; i++) z[i] = b[i] - q[i];          /* :200 */
    orc_ldu_solve(nn, Lptr, Lnode, Lval, Uptr, Unode, Uval, D, r0, z);   /* :201 */
    for (i = 0; i < n; i++) r[i] = r0[i];                /* :202 */
    rho = 1.0; rho_old = 1.0; alpha = 1.0; omega = 1.0;  /* :204-207 */
    for (i = 0; i < n; i++) v[i] = 0.0;                  /* :209 */
    for (i = 0; i < n; i++) p[i] = 0.0;                  /* :210 */
    res2 = dot(r, r, nn);                                /* :212 */

    while (sqrt(res2) > tolerance) {                     /* :214 */
        if (max_iter >= 0 && it >= max_iter) { if (capped) *capped = 1; break; }
        rho = dot(r0, r, nn);                            /* :215 */
        beta = rho / rho_old * alpha / omega;            /* :216 */
        for (i = 0; i < n; i++)                          /* :217 */
            p[i] = r[i] + beta * (p[i] - omega * v[i]);
        orc_matvec(A, 0, p, z);                          /* :218 */
        orc_ldu_solve(nn, Lptr, Lnode, Lval, Uptr, Unode, Uval, D, v, z);   /* :219 */

        alpha = rho / dot(r0, v, nn);                    /* :221 */
        for (i = 0; i < n; i++) s[i] = r[i] - alpha * v[i];   /* :222 */
        orc_matvec(A, 0, s, z);                          /* :223 */
        orc_ldu_solve(nn, Lptr, Lnode, Lval, Uptr, Unode, Uval, D, t, z);   /* :224 */
        omega = dot(s, t, nn) / dot(t, t, nn);           /* :225 */
        for (i = 0; i < n; i++)                          /* :226 */
            x[i] = x[i] + alpha * p[i] + omega * s[i];
        for (i = 0; i < n; i++) r[i] = s[i] - omega * t[i];   /* :227 */

        rho_old = rho;                                   /* :229 */
        res2 = dot(r, r, nn);                            /* :230 */
        it++;
    }
    if (res2_out) *res2_out = res2;
    return it;
}

/* ------------------------------------------------------------------------ */
/* Lanczos / eigensolve                                                      */
/* ------------------------------------------------------------------------ */

/*
 * lanczos, src/eigensolver.f90:27-90, with the start vector supplied by the
 * caller (the reference draws it from a time-seeded RNG, :47-50; we take the
 * un-normalised vector q1 and apply :51 ourselves).
 * T is the Fortran array T(3, n) column-major: T[3*(i-1) + (row-1)].
 * Q is Q(nrow, n) column-major.  w = nrow doubles of scratch.
 */
ORC_API void orc_lanczos(const orc_matrix *A, int32_t n, const double *q1,
                         double *T, double *Q, double *w)
{
    const int64_t nr = A->nrow;
    double alpha, beta, nrm;
    int32_t i, k;
    int64_t l;
#define QC(c) (Q + (int64_t)((c) - 1) * nr)

    for (l = 0; l < 3 * (int64_t)n; l++) T[l] = 0.0;     /* :42 */
    for (l = 0; l < nr * n; l++) Q[l] = 0.0;             /* :43 */
    for (l = 0; l < nr; l++) w[l] = 0.0;                 /* :44 */

    for (l = 0; l < nr; l++) QC(1)[l] = q1[l];
    nrm = sqrt(dot(QC(1), QC(1), (int32_t)nr));          /* :52 */
    for (l = 0; l < nr; l++) QC(1)[l] = QC(1)[l] / nrm;

    orc_matvec(A, 0, QC(1), w);                          /* :55 */
    alpha = dot(QC(1), w, (int32_t)nr);                  /* :56 */
    for (l = 0; l < nr; l++) w[l] = w[l] - alpha * QC(1)[l];   /* :57 */
    beta = sqrt(dot(w, w, (int32_t)nr));                 /* :58 */
    for (l = 0; l < nr; l++) QC(2)[l] = w[l] / beta;     /* :59 */
    T[3 * 0 + 1] = alpha; T[3 * 0 + 2] = beta; T[3 * 0 + 0] = beta;

    for (i = 2; i <= n - 1; i++) {                       /* :67 */
        orc_matvec(A, 0, QC(i), w);                      /* :68 */
        alpha = dot(QC(i), w, (int32_t)nr);              /* :69 */
        for (l = 0; l < nr; l++)                         /* :70 */
            w[l] = w[l] - alpha * QC(i)[l] - beta * QC(i - 1)[l];
        for (k = 1; k <= i - 2; k++) {                   /* :74-76 */
            double c = dot(QC(k), w, (int32_t)nr);
            for (l = 0; l < nr; l++) w[l] = w[l] - c * QC(k)[l];
        }
        beta = sqrt(dot(w, w, (int32_t)nr));             /* :78 */
        for (l = 0; l < nr; l++) QC(i + 1)[l] = w[l] / beta;   /* :79 */
        T[3 * (i - 1) + 1] = alpha;
        T[3 * (i - 1) + 2] = beta;
        T[3 * (i - 1) + 0] = beta;
    }

    orc_matvec(A, 0, QC(n), w);                          /* :87 */
    T[3 * (n - 1) + 1] = dot(QC(n), w, (int32_t)nr);     /* :88 */
#undef QC
}

/*
 * Symmetric tridiagonal eigen-decomposition, standing in for LAPACK
 * dstev('V', ...) at src/eigensolver.f90:174.  Implicit QL with Wilkinson
 * shifts (the classical tql2 scheme), eigenvalues returned ascending in d,
 * eigenvectors in the columns of Z (n x n, column-major).  e has n entries;
 * e[0..n-2] is the sub-diagonal on entry.  Returns 0 on success.
 * LAPACK is un-vendored and unpinned in the reference => parity unpinned here.
 */
ORC_API int32_t orc_tridiag_eig(int32_t n, double *d, double *e, double *Z)
{
    int32_t i, j, k, l, m, iter;
    for (i = 0; i < n; i++)
        for (j = 0; j < n; j++) Z[(int64_t)j * n + i] = (i == j) ? 1.0 : 0.0;
    if (n == 1) return 0;
    e[n - 1] = 0.0;
    for (l = 0; l < n; l++) {
        iter = 0;
        do {
            for (m = l; m < n - 1; m++) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
            }
            if (m != l) {
                double g, r, s, c, p, f, b;
                if (iter++ == 60) return 1;
                g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                s = c = 1.0;
                p = 0.0;
                for (i = m - 1; i >= l; i--) {
                    f = s * e[i];
                    b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (k = 0; k < n; k++) {
                        double *zi1 = Z + (int64_t)(i + 1) * n + k;
                        double *zi = Z + (int64_t)i * n + k;
                        f = *zi1;
                        *zi1 = s * (*zi) + c * f;
                        *zi = c * (*zi) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    /* ascending order, like dstev */
    for (i = 0; i < n - 1; i++) {
        double p = d[k = i];
        for (j = i + 1; j < n; j++) if (d[j] < p) p = d[k = j];
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            for (j = 0; j < n; j++) {
                double tmp = Z[(int64_t)i * n + j];
                Z[(int64_t)i * n + j] = Z[(int64_t)k * n + j];
                Z[(int64_t)k * n + j] = tmp;
            }
        }
    }
    return 0;
}

/*
 * eigensolve, src/eigensolver.f90:160-184: lanczos -> dstev -> V = V*Q ->
 * sign normalisation by V(1,i)/|V(1,i)| -> lambda = T(2,:).
 * V is nrow x n column-major and holds the Ritz vectors on return; scratch
 * needs nrow + n*n + n doubles... allocated internally.
 */
ORC_API int32_t orc_eigensolve(const orc_matrix *A, int32_t n,
                               const double *q1, double *lambda, double *V)
{
    const int64_t nr = A->nrow;
    double *T = malloc(sizeof(double) * 3 * (size_t)n);
    double *w = malloc(sizeof(double) * (size_t)nr);
    double *d = malloc(sizeof(double) * (size_t)n);
    double *e = malloc(sizeof(double) * (size_t)n);
    double *Qm = malloc(sizeof(double) * (size_t)n * n);
    double *row = malloc(sizeof(double) * (size_t)n);
    int32_t i, j, k, info;
    int64_t l;

    orc_lanczos(A, n, q1, T, V, w);                      /* :172 */
    for (i = 0; i < n; i++) d[i] = T[3 * i + 1];
    for (i = 0; i < n - 1; i++) e[i] = T[3 * i + 2];
    info = orc_tridiag_eig(n, d, e, Qm);                 /* :174 */

    for (l = 0; l < nr; l++) {                           /* V = matmul(V, Q) :176 */
        for (j = 0; j < n; j++) {
            double s = 0.0;
            for (k = 0; k < n; k++) s = s + V[(int64_t)k * nr + l] * Qm[(int64_t)j * n + k];
            row[j] = s;
        }
        for (j = 0; j < n; j++) V[(int64_t)j * nr + l] = row[j];
    }
    for (i = 0; i < n; i++) {                            /* :178-180 */
        double sg = V[(int64_t)i * nr] / fabs(V[(int64_t)i * nr]);
        for (l = 0; l < nr; l++) V[(int64_t)i * nr + l] = sg * V[(int64_t)i * nr + l];
    }
    for (i = 0; i < n; i++) lambda[i] = d[i];            /* :182 */
    free(T); free(w); free(d); free(e); free(Qm); free(row);
    return info;
}

/*
 * generalized_lanczos, src/eigensolver.f90:95-155, for A x = lambda B x, with
 * `call B%solve(w, v)` resolved as the reference test sets it up
 * (test/eigensolver_test_generalized_lanczos.f90:150: B%set_solver(cg(tol)),
 * no preconditioner): linear_operator_solve
 * (src/linear_operator/linear_operator_interface.f90:213-233) -> cg_solve with
 * w (= A q_i at that point) as the initial guess.  q1 is the un-normalised
 * start vector (the reference draws it from the time-seeded RNG, :119-120).
 * T(3, n), Q(nrow, n) column-major like orc_lanczos.  The z(:, 0:n) array of the
 * reference (:108, deliberately indexed from 0, :112,136) is kept as is.
 * Returns the total number of inner CG iterations.
 */
ORC_API int64_t orc_generalized_lanczos(const orc_matrix *A, const orc_matrix *B, int32_t n,
                                        const double *q1, double cg_tol, int64_t cg_max_iter,
                                        double *T, double *Q)
{
    const int64_t nr = A->nrow;
    double *w = calloc((size_t)nr, sizeof(double)), *v = calloc((size_t)nr, sizeof(double));
    double *z = calloc((size_t)nr * (n + 1), sizeof(double));   /* z(:, 0:n) */
    double *work = calloc((size_t)nr * 4, sizeof(double));     /* cg_setup: p,q,r,z zeroed */
    double alpha = 0.0, beta = 0.0, d;
    int64_t l, inner = 0;
    int32_t i;
#define QC(c) (Q + (int64_t)((c) - 1) * nr)
#define ZC(c) (z + (int64_t)(c) * nr)

    for (l = 0; l < 3 * (int64_t)n; l++) T[l] = 0.0;
    for (l = 0; l < nr * n; l++) Q[l] = 0.0;

    for (l = 0; l < nr; l++) QC(1)[l] = q1[l];                 /* :119-120 */
    orc_matvec(B, 0, QC(1), w);                                /* :121 */
    d = sqrt(dot(w, QC(1), (int32_t)nr));                      /* :122 */
    for (l = 0; l < nr; l++) QC(1)[l] = QC(1)[l] / d;
    orc_matvec(B, 0, QC(1), ZC(1));                            /* :123 */

    for (i = 1; i <= n - 1; i++) {                             /* :128 */
        orc_matvec(A, 0, QC(i), w);                            /* :129 */
        for (l = 0; l < nr; l++) v[l] = w[l] - beta * ZC(i - 1)[l];   /* :130 */
        alpha = dot(v, QC(i), (int32_t)nr);                    /* :131 */
        for (l = 0; l < nr; l++) v[l] = v[l] - alpha * ZC(i)[l];      /* :132 */

        inner += orc_cg_solve(B, w, v, cg_tol, cg_max_iter, work, NULL, NULL);   /* :134 */

        beta = sqrt(dot(w, v, (int32_t)nr));                   /* :136 */
        for (l = 0; l < nr; l++) QC(i + 1)[l] = w[l] / beta;   /* :137 */
        for (l = 0; l < nr; l++) ZC(i + 1)[l] = v[l] / beta;   /* :138 */
        T[3 * (i - 1) + 1] = alpha;                            /* :140-142 */
        T[3 * (i - 1) + 2] = beta;
        T[3 * (i - 1) + 0] = beta;
    }
    orc_matvec(A, 0, QC(n), v);                                /* :145 */
    for (l = 0; l < nr; l++) v[l] = v[l] - beta * ZC(n)[l];    /* :146 */
    T[3 * (n - 1) + 1] = dot(QC(n), v, (int32_t)nr);           /* :147 (i == n after the loop) */
#undef QC
#undef ZC
    free(w); free(v); free(z); free(work);
    return inner;
}

/* ------------------------------------------------------------------------ */
/* Row-block partition + halo lists (index work for the multi-GPU path).     */
/* Not in the reference (it is serial); the single source of truth the       */
/* product's partitioner (sigma_b200/csrc/partition.cpp) must match           */
/* bit-exactly.  The seam it generalises is the block loop of                */
/* composite_matvec_add, src/matrix/sparse_matrix_composites.f90:1076-1100.   */
/* ------------------------------------------------------------------------ */

/* Row offsets part[0..P] (0-based, part[P] = n) balancing stored entries:
 * part[r] = smallest row i such that (ptr(i+1) - 1) >= r * nnz / P. */
ORC_API void orc_partition_rows(int32_t n, const int32_t *ptr, int32_t P,
                                int32_t *part)
{
    int64_t nnz = (int64_t)ptr[n] - 1;
    int32_t r, i = 0;
    part[0] = 0;
    for (r = 1; r < P; r++) {
        int64_t target = (nnz * r) / P;
        while (i < n && (int64_t)ptr[i] - 1 < target) i++;
        part[r] = i;
    }
    part[P] = n;
}

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/*
 * For the rank owning rows [lo, hi) (0-based) of a CSR matrix: halo = sorted
 * unique 1-based global column ids outside (lo, hi]; local_node = columns
 * renumbered 1-based into [owned | halo] (owned col c -> c - lo, halo col ->
 * (hi - lo) + 1 + position in halo).  ptr/node are the GLOBAL arrays.
 * halo must have room for every entry of the block; returns the halo length.
 */
ORC_API int32_t orc_halo_build(int32_t lo, int32_t hi, const int32_t *ptr,
                               const int32_t *node, int32_t *halo,
                               int32_t *local_node)
{
    int64_t k0 = ptr[lo] - 1, k1 = ptr[hi] - 1, k;
    int32_t nh = 0, i, m = 0;
    for (k = k0; k < k1; k++) {
        int32_t c = node[k];
        if (c <= lo || c > hi) halo[nh++] = c;
    }
    qsort(halo, (size_t)nh, sizeof(int32_t), cmp_i32);
    for (i = 0; i < nh; i++)
        if (i == 0 || halo[i] != halo[i - 1]) halo[m++] = halo[i];
    nh = m;
    for (k = k0; k < k1; k++) {
        int32_t c = node[k];
        if (c > lo && c <= hi) {
            local_node[k - k0] = c - lo;
        } else {
            int32_t a = 0, b = nh - 1;
            while (a < b) {
                int32_t mid = (a + b) / 2;
                if (halo[mid] < c) a = mid + 1; else b = mid;
            }
            local_node[k - k0] = (hi - lo) + 1 + a;
        }
    }
    return nh;
}
