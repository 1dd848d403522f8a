/*
 * sigma_b200.h -- C-ABI of the B200-native SpMV + Krylov hot path of SiGMA
 * (danshapero/sigma).
 *
 * The reference is Fortran 2003 and has no FFI for this path (its only C
 * header, include/graphs.h, is dead code that covers graphs alone; SURVEY.md
 * F4).  The seams this ABI sits behind are therefore the reference's
 * type-bound procedures; every entry point below names the reference
 * procedure whose BODY it replaces (file:line relative to the reference
 * root).  fortran/sigma_b200_shim.f90 holds the iso_c_binding stubs a
 * maintainer adds.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns an int status
 *    (SIGB_OK = 0).  The reference's error convention is
 *    `print *, ...; call exit(1)` (e.g. src/solver/cg_solvers.f90:61-65): the
 *    shim prints sigb_last_error() and exits 1 on a non-zero status.
 *  - all reals are IEEE fp64 (dp = kind(0.d0), src/types.f90:5); all indices
 *    are int32 and 1-BASED, passed exactly as the Fortran arrays hold them
 *    (cs_graph: ptr(n+1), node(ne), src/graph/formats/cs_graphs.f90:16;
 *    ellpack_graph: node(max_d, n) column-major + degrees(n),
 *    src/graph/formats/ellpack_graphs.f90:10-20).
 *  - pointers are HOST pointers unless the function name ends in _dev.
 *  - not thread-safe, like the reference (solvers are stateful,
 *    src/solver/cg_solvers.f90:13).  One process drives one GPU; multi-GPU
 *    runs use one process per GPU joined through a sigb_comm_t.
 *  - there is no CPU fallback: every compute entry fails with
 *    SIGB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SIGMA_B200_H
#define SIGMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGB_API __attribute__((visibility("default")))

/* ---- status codes ------------------------------------------------------ */
enum {
    SIGB_OK            = 0,
    SIGB_ERR_ARG       = 1,  /* bad argument (null handle, size mismatch ...) */
    SIGB_ERR_CUDA      = 2,  /* CUDA runtime failure / no usable device       */
    SIGB_ERR_STATE     = 3,  /* call order (solve before setup ...)           */
    SIGB_ERR_NONSQUARE = 4,  /* cg_setup/bicgstab_setup/jacobi_setup on a
                                non-square operator (cg_solvers.f90:61-65)     */
    SIGB_ERR_ISOLATED  = 5,  /* ELLPACK row with no edge: the reference would
                                read x(0) (README.md:71-73); rejected here     */
    SIGB_ERR_COMM      = 6,  /* NCCL / peer-memory failure                    */
    SIGB_ERR_UNSUPPORTED = 7
};

/* storage order of a compressed-sparse graph */
enum { SIGB_ROW = 0 /* csr_matrix */, SIGB_COL = 1 /* csc_matrix */ };

typedef struct sigb_graph_s  *sigb_graph_t;   /* device mirror of cs_graph / ellpack_graph */
typedef struct sigb_matrix_s *sigb_matrix_t;  /* device mirror of cs_matrix / ellpack_matrix */
typedef struct sigb_solver_s *sigb_solver_t;  /* cg_solver / bicgstab_solver / jacobi_solver */
typedef struct sigb_comm_s   *sigb_comm_t;    /* multi-GPU communicator */

/* ---- runtime ----------------------------------------------------------- */

/* Select the CUDA device this process drives and create the library stream.
 * device < 0: use LOCAL_RANK from the environment, else 0. */
SIGB_API int sigb_init(int device);
SIGB_API int sigb_finalize(void);
/* Message of the last failing call on this thread ("" if none). */
SIGB_API const char *sigb_last_error(void);
/* Library version string. */
SIGB_API const char *sigb_version(void);
/* Use an existing cudaStream_t for every launch (NULL restores the library's
 * own stream).  Lets a host time the library's kernels with its own events. */
SIGB_API int sigb_set_stream(void *cuda_stream);
SIGB_API int sigb_synchronize(void);
/* Number of kernels this library has launched since sigb_init (diagnostic). */
SIGB_API int64_t sigb_launch_count(void);

/* Device buffers (so a C / Fortran host needs no other CUDA binding). */
SIGB_API int sigb_dev_alloc(int64_t bytes, void **ptr_dev);
SIGB_API int sigb_dev_free(void *ptr_dev);
SIGB_API int sigb_copy_h2d(void *dst_dev, const void *src, int64_t bytes);
SIGB_API int sigb_copy_d2h(void *dst, const void *src_dev, int64_t bytes);

/* ---- graphs (sparsity patterns) ---------------------------------------- */

/* Mirror a cs_graph (src/graph/formats/cs_graphs.f90:11-60).
 *  n      : g%n  (rows for SIGB_ROW, columns for SIGB_COL)
 *  m      : g%m
 *  ptr1   : g%ptr(1:n+1), 1-based;  node1 : g%node(1:ne), 1-based, UNSORTED,
 *           stored order is preserved on the device.
 * For SIGB_COL the library also builds, once and on the device, the stable
 * transpose to CSR (row i's entries in ascending source column), which makes
 * the CSR kernel accumulate into each y(i) in exactly the order
 * csc_matvec_add does (src/matrix/formats/cs_matrices.f90:627-647). */
SIGB_API int sigb_cs_graph_create(int32_t n, int32_t m, const int32_t *ptr1,
                                  const int32_t *node1, int order,
                                  sigb_graph_t *g);

/* Mirror an ellpack_graph (src/graph/formats/ellpack_graphs.f90:10-61).
 *  node_cm : g%node(max_d, n) as stored (slot index fastest), padding slots
 *            included (copies of the row's last neighbour, :164);
 *  degrees : g%degrees(n).  Rows with degree 0 -> SIGB_ERR_ISOLATED.
 * The device copy is re-laid out slot-major (slot k of all rows contiguous). */
SIGB_API int sigb_ell_graph_create(int32_t n, int32_t m, int32_t max_d,
                                   const int32_t *node_cm,
                                   const int32_t *degrees, sigb_graph_t *g);

/* Reference counting like graph_interface add_reference/remove_reference
 * (src/graph/graph_interfaces.f90:345-363).  create returns refcount 1. */
/* The graph builders fed by an EDGE STREAM, on the device:
 *   cs_graph_build       src/graph/formats/cs_graphs.f90:109-197
 *   ellpack_graph_build  src/graph/formats/ellpack_graphs.f90:105-170
 * as copy_graph / convert_graph_type drive them (graph_interfaces.f90:276-318).
 * src_i / src_j: the source graph's edges in ITS iteration order (1-based; an
 * endpoint 0 in the second role is a null edge and is skipped); trans != 0 swaps
 * the roles of the endpoints.  Line l of the result holds its DISTINCT neighbours
 * in the order of their first appearance in the stream -- what the reference's
 * first-free-slot insertion with its duplicate check and final pruning produces,
 * bit for bit.  n lines, ids in 1..m.  order as in sigb_cs_graph_create.
 * The ellpack form pads every row with its last neighbour (:164) and refuses a
 * row without edges (SIGB_ERR_ISOLATED), like sigb_ell_graph_create. */
SIGB_API int sigb_cs_graph_build(int32_t n, int32_t m, int64_t count,
                                 const int32_t *src_i, const int32_t *src_j,
                                 int trans, int order, sigb_graph_t *g);
SIGB_API int sigb_ell_graph_build(int32_t n, int32_t m, int64_t count,
                                  const int32_t *src_i, const int32_t *src_j,
                                  int trans, sigb_graph_t *g);
SIGB_API int sigb_graph_retain(sigb_graph_t g);
SIGB_API int sigb_graph_release(sigb_graph_t g);

/* Read back the device-built transpose of a cs graph (index parity checks):
 * ptr_t1 has (other dimension)+1 entries, node_t1 has ne entries, 1-based.
 * Either may be NULL.  For SIGB_COL graphs this is the CSR form used by
 * matvec; for SIGB_ROW graphs it is the form used by matvec_t. */
SIGB_API int sigb_cs_graph_get_transpose(sigb_graph_t g, int32_t *ptr_t1,
                                         int32_t *node_t1);

/* ---- matrices ---------------------------------------------------------- */

/* A%set_graph(g): share the pattern, values zero
 * (cs_matrix_set_graph src/matrix/formats/cs_matrices.f90:259-289,
 *  ellpack_matrix_set_graph src/matrix/formats/ellpack_matrices.f90:141-164). */
SIGB_API int sigb_matrix_create(sigb_graph_t g, sigb_matrix_t *A);
/* Upload / refresh A%val (cs: ne doubles in node order; ellpack:
 * val(max_d, n) as stored, padding zeros included).  This is what the shim
 * calls when a host mutator (set_value/add_value/zero/scalar_multiply/...,
 * cs_matrices.f90:448-490,840-1099) has marked the mirror dirty. */
SIGB_API int sigb_matrix_set_values(sigb_matrix_t A, const double *val,
                                    int64_t count);
SIGB_API int sigb_matrix_destroy(sigb_matrix_t A);
SIGB_API int sigb_matrix_get_dims(sigb_matrix_t A, int32_t *nrow, int32_t *ncol,
                                  int64_t *nnz);
/* Values of the device-built transpose, in sigb_cs_graph_get_transpose order. */
SIGB_API int sigb_matrix_get_transpose_values(sigb_matrix_t A, double *val_t);

/* ---- matrix copy / format conversion on the device -----------------------
 * The step before the hot path: today an O(ne * d) host
 * scan per copy (cs_graph_build, src/graph/formats/cs_graphs.f90:163-183). */
enum { SIGB_FMT_CSR = 1, SIGB_FMT_CSC = 2, SIGB_FMT_ELLPACK = 3 };

/* call B%copy_matrix(A, trans) with B of the given format:
 * cs_matrix_copy_matrix (src/matrix/formats/cs_matrices.f90:294-322) /
 * ellpack_matrix_copy_matrix (src/matrix/formats/ellpack_matrices.f90:169-198)
 * = build_graph_from_matrix + copy_matrix_values
 * (src/matrix/sparse_matrix_interfaces.f90:692-772).  B gets its own graph,
 * whose arrays equal what the reference's builders produce bit for bit: every
 * target line holds its entries in the source's iteration order (first-free-
 * slot insertion, cs_graphs.f90:163-183; ellpack_graphs.f90:150-168 with
 * last-neighbour padding and val = 0 in the padding).  trans != 0 copies the
 * transpose.  An ellpack target with an empty row -> SIGB_ERR_ISOLATED.
 * A must be a stored csr / csc / ellpack matrix without duplicate edges. */
SIGB_API int sigb_matrix_copy(sigb_matrix_t A, int format, int trans,
                              sigb_matrix_t *B);
/* Shape of a stored matrix's arrays: n_lines = g%n, n_ids = g%m, ne = g%ne,
 * max_d = g%max_d (any pointer may be NULL). */
SIGB_API int sigb_matrix_get_format(sigb_matrix_t A, int *format,
                                    int32_t *n_lines, int32_t *n_ids,
                                    int64_t *ne, int32_t *max_d);
/* Read the arrays back so the host object can be filled in (or checked):
 *  csr / csc : ptr_or_degrees = g%ptr(n_lines+1), node = g%node(ne),
 *              val = A%val(ne), all 1-based as the Fortran holds them;
 *  ellpack   : ptr_or_degrees = g%degrees(n_lines), node = g%node(max_d, n),
 *              val = A%val(max_d, n) (slot index fastest).
 * Any pointer may be NULL. */
SIGB_API int sigb_matrix_get_arrays(sigb_matrix_t A, int32_t *ptr_or_degrees,
                                    int32_t *node, double *val);

/* Assembly on the device: apply `call A%add_value(i1[c], j1[c], z[c])` for
 * c = 0 .. count-1 IN ORDER (csr_matrix_add_value
 * src/matrix/formats/cs_matrices.f90:868-891, csc_matrix_add_value :924-947,
 * ellpack_matrix_add_value src/matrix/formats/ellpack_matrices.f90:471-493;
 * add_multiple_values is the same statement over B(k, l), cs_matrices.f90:
 * 934-967; the loop of examples/fem.f90:43-47 is such a stream).  Every stored
 * entry receives its contributions in ascending call index, so the values
 * are bit-identical to the serial loop.  All (i, j) must already be in the
 * sparsity pattern: otherwise SIGB_ERR_ARG and nothing is added (the
 * reference's reallocation path, cs_matrices.f90:888-890, is not mirrored).
 * The device values are then ahead of the host copy: read them back with
 * sigb_matrix_get_arrays. */
SIGB_API int sigb_matrix_add_values(sigb_matrix_t A, int64_t count,
                                    const int32_t *i1, const int32_t *j1,
                                    const double *z);

/* ---- matvec ------------------------------------------------------------ */

/* linear_operator%matvec / matvec_t
 * (src/linear_operator/linear_operator_interface.f90:185-208): y = op(A) x.
 * trans = 0: A, trans = 1: A^T.  The zero-fill pass is fused away. */
SIGB_API int sigb_matvec(sigb_matrix_t A, int trans, const double *x, double *y);
/* matvec_add / matvec_t_add seam: y = y + op(A) x
 * (csr_matvec_add cs_matrices.f90:600-622, csc_matvec_add :627-647,
 *  ellpack_matvec_add ellpack_matrices.f90:640-665, _t_add :670-693). */
SIGB_API int sigb_matvec_add(sigb_matrix_t A, int trans, const double *x,
                             double *y);
/* Same with device-resident vectors (add != 0 selects matvec_add). */
SIGB_API int sigb_matvec_dev(sigb_matrix_t A, int trans, const double *x_dev,
                             double *y_dev, int add);
/* Fused y = A x and *dot = x . y (the CG hot pair cg_solvers.f90:134-135),
 * device vectors, result copied to the host scalar (dot may be NULL: nothing
 * is copied and the call stays asynchronous). */
SIGB_API int sigb_matvec_dot_dev(sigb_matrix_t A, const double *x_dev,
                                 double *y_dev, double *dot);

/* ---- operator expressions ----------------------------------------------
 * In the reference every sparse matrix IS a linear_operator, and operators
 * compose lazily: `A + B`, `A * B`, `adjoint(A)` and the block composite
 * `sparse_matrix`.  Here an expression is again a sigb_matrix_t: it can be
 * passed to sigb_matvec*, sigb_solver_setup / _solve (cg, bicgstab; jacobi
 * where get_value is defined, i.e. not on products), sigb_lanczos*, and to
 * these constructors again.  Intermediate vectors stay on the device and every
 * leaf keeps the reference's accumulation order.  Each constructor takes a
 * reference on its operands (add_reference); sigb_matrix_destroy on an
 * expression drops them again (operator_sum_destroy
 * linear_operator_sums.f90:136-159).  Row-sharded operators cannot be
 * operands. */
typedef sigb_matrix_t sigb_operator_t;   /* readability only: same handle */

/* C = A + B: add_operators (src/linear_operator/linear_operator_sums.f90:38-72);
 * matvec_add runs the summands in order into the same y (:100-131). */
SIGB_API int sigb_operator_sum(sigb_operator_t A, sigb_operator_t B,
                               sigb_operator_t *C);
/* C = A * B: multiply_operators
 * (src/linear_operator/linear_operator_products.f90:39-73); the scratch vectors
 * z1, z2 of temp_vec_size doubles (:60-61) live on the device;
 * matvec_add: last factor first, y = y + z2 (:78-114); matvec_t_add: first
 * factor first, transposed (:119-150). */
SIGB_API int sigb_operator_product(sigb_operator_t A, sigb_operator_t B,
                                   sigb_operator_t *C);
/* B = adjoint(A) (src/linear_operator/linear_operator_adjoints.f90:28-44):
 * matvec_add <-> matvec_t_add of A (:62-86). */
SIGB_API int sigb_operator_adjoint(sigb_operator_t A, sigb_operator_t *B);
/* type(sparse_matrix) as a composite of num_row_mats x num_col_mats blocks
 * (src/matrix/sparse_matrix_composites.f90:41-49): set_block_sizes(rows, cols)
 * (:226-262) + set_submatrix(it, jt, blocks[(it-1)*num_col_mats + (jt-1)])
 * (:1031-1065).  Every block must be given and match its slot's dimensions
 * ("Inconsistent dimensions for sub-matrix").  matvec_add is the block-row
 * loop of composite_matvec_add (:1076-1100), matvec_t_add the block-column
 * loop (:1105-1129), on offset slices of the device vectors. */
SIGB_API int sigb_composite_create(int32_t num_row_mats, int32_t num_col_mats,
                                   const int32_t *rows, const int32_t *cols,
                                   const sigb_operator_t *blocks,
                                   sigb_operator_t *A);
/* add_reference (linear_operator_interface.f90:285-291); undone by
 * sigb_matrix_destroy, which frees the mirror when the count reaches zero. */
SIGB_API int sigb_matrix_retain(sigb_operator_t A);

/* ---- solvers ----------------------------------------------------------- */

/* cg(tolerance) (src/solver/cg_solvers.f90:36-47); tolerance < 0 selects the
 * reference default 1e-16 (cg_set_params :95-111). */
SIGB_API int sigb_cg_create(double tolerance, sigb_solver_t *s);
/* bicgstab(tolerance) (src/solver/bicgstab_solvers.f90:36-47). */
SIGB_API int sigb_bicgstab_create(double tolerance, sigb_solver_t *s);
/* jacobi() (src/solver/jacobi_solvers.f90:26-32). */
SIGB_API int sigb_jacobi_create(sigb_solver_t *s);

/* ldu(incomplete, level) (src/solver/ldu_solvers.f90:73-86): always the
 * incomplete factorisation of level 0, like the reference (:145,151).
 * A ~= (I + L) D (I + U) on the strict triangles of A's own pattern.
 *  - sigb_solver_setup(s, A): sparse_ldu_setup (:95-130) -- the patterns of
 *    L and U (incomplete_ldu_sparsity_pattern :396-441) and the level schedules
 *    once per pattern, the numeric factorisation
 *    (sparse_static_pattern_ldu_factorization :275-387) on every call.  A must
 *    be a stored csr / csc / ellpack matrix on one GPU.
 *  - sigb_solver_solve(s, A, x, b, NULL): ldu_solve (:160-176), x = b, forward
 *    solve with I + L, x = x / D, backward solve with I + U.
 *  - as the pc of a cg solver: cg_solve_pc (cg_solvers.f90:155-194) calls it
 *    once per iteration.
 * Rows that do not depend on each other run concurrently (level scheduling);
 * each row does the reference's arithmetic in the reference's order, so
 * factors and solves are bit-identical to the serial loops. */
SIGB_API int sigb_ldu_create(sigb_solver_t *s);
/* Sizes of the factors after setup: n rows, nL / nU stored entries of L / U,
 * and how many row levels the forward / backward sweeps take. */
SIGB_API int sigb_ldu_get_sizes(sigb_solver_t s, int32_t *n, int64_t *nL,
                                int64_t *nU, int32_t *n_forward_levels,
                                int32_t *n_backward_levels);
/* Read the factors back (parity checks): csr patterns 1-based (n + 1 / nL /
 * nU entries), values, D(n).  Any pointer may be NULL. */
SIGB_API int sigb_ldu_get_factors(sigb_solver_t s, int32_t *Lptr1,
                                  int32_t *Lnode1, double *Lval,
                                  int32_t *Uptr1, int32_t *Unode1,
                                  double *Uval, double *D);
/* Host-only index work behind the setup (no GPU needed; bit-exact contract with
 * the oracle's restatement of incomplete_ldu_sparsity_pattern): from the rows
 * of A in iteration order (ptr1 n+1, node1 ne, 1-based) build the patterns,
 * the destination of every entry of A in the combined value array
 * [ Lval | Uval | D ] (0-based) and the level schedules: rows (1-based)
 * grouped by level, level l = rows[lev[l] .. lev[l+1]).  Output arrays are
 * sized by the caller: ptr arrays and level pointers n + 1, node arrays and
 * dest ne, row lists n. */
SIGB_API int sigb_ldu_symbolic(int32_t n, const int32_t *ptr1,
                               const int32_t *node1, int32_t *Lptr1,
                               int32_t *Lnode1, int32_t *Uptr1, int32_t *Unode1,
                               int64_t *dest, int32_t *forward_rows,
                               int32_t *forward_lev, int32_t *n_forward_levels,
                               int32_t *backward_rows, int32_t *backward_lev,
                               int32_t *n_backward_levels);

/* solver%setup(A): cg_setup cg_solvers.f90:52-90, bicgstab_setup
 * bicgstab_solvers.f90:52-100 (allocate + zero work vectors, iterations = 0),
 * jacobi_setup jacobi_solvers.f90:37-63 (idiag(i) = 1 / A(i,i)). */
SIGB_API int sigb_solver_setup(sigb_solver_t s, sigb_matrix_t A);
/* set_params (cg_solvers.f90:95-111): tolerance < 0 -> 1e-16. */
SIGB_API int sigb_solver_set_params(sigb_solver_t s, double tolerance);
/* NOT in the reference: a safety cap on iterations per solve call (the
 * reference loops `do while (dsqrt(res2) > tolerance)` with no cap,
 * cg_solvers.f90:133).  cap < 0 (default) = no cap.  A capped solve still
 * returns SIGB_OK; query it with sigb_solver_get_info. */
SIGB_API int sigb_solver_set_max_iterations(sigb_solver_t s, int64_t cap);
/* NOT in the reference: which form of the CG loop (cg_solvers.f90:133-146)
 * runs -- 1: the whole loop as ONE persistent cooperative kernel, 0: three
 * kernels per iteration, -1 (default): the library chooses by shard size
 * (persistent at <= 3 M rows per GPU).  Same recurrence, statement order and
 * rounding either way; bench.py's parity gate uses it to check the small
 * instance with the kernels the full-size run takes. */
SIGB_API int sigb_solver_set_persistent(sigb_solver_t s, int mode);
/* NOT in the reference, a PARITY AID: with on != 0 every dot product of the cg /
 * bicgstab recurrence is summed strictly left to right by one thread -- the
 * serial `sum(a * b)` of the reference (cg_solvers.f90:135,140;
 * bicgstab_solvers.f90:155,160,164,168), rounded products, no FMA.  Everything else
 * in the solvers is element-wise and already reproduces the reference statement for
 * statement, so a strict-order solve must equal the serial loops BIT FOR BIT:
 * iterates, scalars and the stopping iteration (tests/test_gpu_solvers.py).  One
 * GPU, kernel-per-phase path; slow by construction (one thread per dot product). */
SIGB_API int sigb_solver_set_strict_order(sigb_solver_t s, int on);

/* solver%solve(A, x, b [, pc]): linear_solve / linear_solve_pc
 * (cg_solve cg_solvers.f90:116-150, cg_solve_pc :155-194, bicgstab_solve
 * bicgstab_solvers.f90:124-177, bicgstab_solve_pc :182-237, jacobi_solve
 * jacobi_solvers.f90:68-81, ldu_solve ldu_solvers.f90:160-176).  x is the
 * initial guess on entry and the solution on return.  pc may be NULL; the
 * device preconditioners are jacobi (cg, bicgstab) and ldu (cg, one GPU);
 * anything else is SIGB_ERR_UNSUPPORTED.  The whole iteration runs
 * on the device; the loop stops at the same test as the reference,
 * evaluated every iteration. */
SIGB_API int sigb_solver_solve(sigb_solver_t s, sigb_matrix_t A, double *x,
                               const double *b, sigb_solver_t pc);
SIGB_API int sigb_solver_solve_dev(sigb_solver_t s, sigb_matrix_t A,
                                   double *x_dev, const double *b_dev,
                                   sigb_solver_t pc);
/* iterations: solver%iterations, accumulated over solve calls since the last
 * setup (cg_solvers.f90:72,145).  res2: last value of the stopping quantity
 * squared.  capped: 1 if the last solve hit the safety cap. */
SIGB_API int sigb_solver_get_info(sigb_solver_t s, int64_t *iterations,
                                  double *res2, int *capped);
/* Copy a work vector back (diagnostics / parity): name is one of
 * "p","q","r","z","r0","v","s","t","idiag". */
SIGB_API int sigb_solver_get_vector(sigb_solver_t s, const char *name,
                                    double *out);
SIGB_API int sigb_solver_destroy(sigb_solver_t s);

/* ---- eigensolver ------------------------------------------------------- */

/* lanczos(A, T, Q) (src/eigensolver.f90:27-90): n = size(T, 2) steps.
 *  q1 : un-normalised start vector (nrow).  The reference draws it from a
 *       time-seeded RNG (:47-50); pass NULL to have the library draw one
 *       from `seed` (splitmix64 -> uniform [-1, 1)), else it is used as is
 *       and normalised as :51 does.
 *  T  : T(3, n) column-major;  Q : Q(nrow, n) column-major. */
SIGB_API int sigb_lanczos(sigb_matrix_t A, int32_t n, const double *q1,
                          uint64_t seed, double *T, double *Q);
/* lanczos with the basis resident on the device: Q_dev (nrow x n, column-major)
 * and the optional start vector q1_dev are DEVICE pointers; only T comes back to
 * the host.  Same loop (eigensolver.f90:27-90). */
SIGB_API int sigb_lanczos_dev(sigb_matrix_t A, int32_t n, const double *q1_dev,
                              uint64_t seed, double *T, double *Q_dev);
/* eigensolve(A, lambda, V) (src/eigensolver.f90:160-184): lanczos, symmetric
 * tridiagonal eigen-solve (stands in for LAPACK dstev, :174), V = V*Q on the
 * device, sign normalisation (:178-180).  lambda ascending. */
SIGB_API int sigb_eigensolve(sigb_matrix_t A, int32_t n, const double *q1,
                             uint64_t seed, double *lambda, double *V);

/* generalized_lanczos(A, B, T, Q) (src/eigensolver.f90:95-155) for
 * A x = lambda B x.  Every step runs `call B%solve(w, v)` (:134): b_solver is
 * the solver attached with B%set_solver (set up on B; the reference test uses
 * cg(1.0d-15), test/eigensolver_test_generalized_lanczos.f90:150) and b_pc the
 * optional preconditioner attached with B%set_preconditioner (NULL if none).
 * The solve runs on the device with w = A q_i as its initial guess, like the
 * reference facade (linear_operator_interface.f90:213-233). */
SIGB_API int sigb_generalized_lanczos(sigb_matrix_t A, sigb_matrix_t B,
                                      sigb_solver_t b_solver, sigb_solver_t b_pc,
                                      int32_t n, const double *q1, uint64_t seed,
                                      double *T, double *Q);
/* generalized_eigensolve(A, B, lambda, V) (src/eigensolver.f90:189-208):
 * generalized_lanczos, tridiagonal eigen-solve, V = V*Q (no sign fix). */
SIGB_API int sigb_generalized_eigensolve(sigb_matrix_t A, sigb_matrix_t B,
                                         sigb_solver_t b_solver, sigb_solver_t b_pc,
                                         int32_t n, const double *q1, uint64_t seed,
                                         double *lambda, double *V);

/* ---- multi-GPU: row-sharded operators ----------------------------------
 * Not in the reference (serial).  The seam is the block-row loop of
 * composite_matvec_add (src/matrix/sparse_matrix_composites.f90:1076-1100,
 * "This loop can be parallelized" :1086).  One process per GPU. */

/* Host-only index work (no GPU needed; bit-exact contract with
 * oracle/sigma_oracle.c orc_partition_rows / orc_halo_build). */
SIGB_API int sigb_partition_rows(int32_t n, const int32_t *ptr1, int32_t nparts,
                                 int32_t *part /* nparts+1, 0-based */);
/* For rows [lo, hi) (0-based) given that block's own ptr (hi-lo+1 entries,
 * 1-based, may start above 1) and GLOBAL 1-based column ids:
 *  halo       : out, sorted unique global columns outside the block
 *               (capacity = entries in the block); *nhalo its length;
 *  local_node : out, columns renumbered 1-based into [owned | halo]. */
SIGB_API int sigb_halo_build(int32_t lo, int32_t hi, const int32_t *ptr_blk1,
                             const int32_t *node_glob1, int32_t *halo,
                             int32_t *nhalo, int32_t *local_node);

/* Diagnostic, host-only: the row tiling the streaming CSR kernel uses for a
 * pattern (<= 2045 entries and <= 512 rows per tile; a longer row alone).
 * tiles holds {first row, end row, first entry, end entry} per tile, 0-based,
 * capacity n tiles. */
SIGB_API int sigb_debug_row_tiles(int32_t n, const int32_t *ptr1, int32_t *tiles,
                                  int32_t *ntiles);
/* The balanced tiling used for row-sharded operators: exactly m * groups tiles of
 * nearly equal entry counts when the limits allow it (else the greedy tiling). */
SIGB_API int sigb_debug_row_tiles_balanced(int32_t n, const int32_t *ptr1, int32_t groups,
                                           int32_t *tiles, int32_t *ntiles);
/* Diagnostic, needs the GPU: the same tiling built by the device
 * path (csrc/tiles_device.cu) from an uploaded copy of
 * ptr1; same output layout.  Must equal sigb_debug_row_tiles entry for entry. */
SIGB_API int sigb_debug_row_tiles_dev(int32_t n, const int32_t *ptr1, int32_t *tiles,
                                      int32_t *ntiles);

/* Diagnostic (host only, no GPU): the plan of a statically scheduled ILDU sweep (csrc/ldu_sweep.h) for the
 * strictly triangular factor given by its rows (ptr1 n+1, node1, 1-based); backward = 0 for L (rows ascending),
 * 1 for U (rows descending); levels = depth of the level schedule of the same sweep.  info[16] = eligible, R
 * (rows per chunk), sigma (skew), C (chunks), trips, W (ring depth), S_max, w16_max, nstage, stage_bytes,
 * threads, total (lo, hi), total_s (lo, hi), has_far.  The arrays may be null (a first call returns the sizes):
 * trip_table trips x 8 int32 (vlo, w, w16, S, off lo/hi, soff lo/hi), src and valmap total_s, cnt total.
 * tests/test_ldu_sweep_plan.py replays the device kernel on these arrays against the serial solves. */
SIGB_API int sigb_debug_ldu_sweep_plan(int32_t n, const int32_t *ptr1, const int32_t *node1, int backward,
                                       int64_t levels, int32_t *info, int32_t *trip_table, int32_t *src,
                                       uint8_t *cnt, int64_t *valmap);
/* Diagnostic: SM cycles per phase of the persistent CG kernel, accumulated since
 * the last call (then reset), for its first / middle / last CTA:
 * out[cta * 9 + k], k = 0 SpMV, 1 barrier of reduction 1, 2 cross-GPU part of
 * reduction 1, 3 residual update, 4 barrier of reduction 2, 5 cross-GPU part of
 * reduction 2, 6 direction update, 7 closing barrier, 8 iterations counted.
 * *supported = 0 and zeros unless the library was built with
 * -DSIGB_PHASE_TIMERS (csrc/Makefile: make VARIANT=_timers DEFS=-DSIGB_PHASE_TIMERS). */
SIGB_API int sigb_debug_cg_phase_cycles(unsigned long long *out27, int *supported);

/* Diagnostic, same builds: SM cycles of thread 0 of every CTA of the streaming
 * CSR kernel summed over the grid and the launches since the last call (then
 * reset): out[0] waiting for the staged tile, [1] products (gathers), [2] row
 * sums, [3] whole pass, [4] staged tiles processed, [5] CTA passes. */
SIGB_API int sigb_debug_spmv_tile_cycles(unsigned long long *out6, int *supported);

/* Communicator.  unique_id is SIGB_UNIQUE_ID_BYTES bytes produced by
 * sigb_comm_unique_id on rank 0 and broadcast by the host (torch.distributed,
 * MPI, a file ...). */
#define SIGB_UNIQUE_ID_BYTES 128
SIGB_API int sigb_comm_unique_id(void *unique_id);
SIGB_API int sigb_comm_create(const void *unique_id, int rank, int nranks,
                              sigb_comm_t *comm);
SIGB_API int sigb_comm_destroy(sigb_comm_t comm);
SIGB_API int sigb_comm_info(sigb_comm_t comm, int *rank, int *nranks,
                            int *peer_access);

/* Row block [part[rank], part[rank+1]) of a global n x n CSR matrix.
 *  ptr_blk1   : the block's slice of the global ptr (nloc+1 entries, 1-based,
 *               may start above 1)
 *  node_glob1 : GLOBAL 1-based column ids of the block's entries, stored order
 *  send_counts: nranks entries, how many owned rows rank q needs from us
 *  send_rows1 : those rows as 1-based LOCAL row ids, grouped by destination
 *               rank ascending, each group in ascending order
 * The halo list (what we need) is derived here from node_glob1 with
 * sigb_halo_build; the send lists are its mirror image on the owners, which
 * the host obtains with one all-to-all of the halo lists (torch.distributed,
 * MPI_Alltoallv ...; sigma_b200/distributed.py shows it) -- "graph-derived"
 * index work, no values involved.  The library splits the rows into interior
 * and boundary tiles and mirrors the renumbered local CSR on the device.
 * Values are uploaded with sigb_matrix_set_values (the block's val slice).
 * The resulting operator takes the vectors' OWNED slices in sigb_matvec /
 * sigb_matvec_add, sigb_solver_* (cg, bicgstab, jacobi), sigb_lanczos[_dev] and
 * sigb_eigensolve (the sign convention V(1, i) > 0 is taken from the rank that owns
 * global row 1); matvec_t, copies, expressions and ldu refuse it
 * (SIGB_ERR_UNSUPPORTED). */
SIGB_API int sigb_dist_csr_create(sigb_comm_t comm, int32_t n_global,
                                  const int32_t *part, const int32_t *ptr_blk1,
                                  const int32_t *node_glob1,
                                  const int32_t *send_counts,
                                  const int32_t *send_rows1, sigb_matrix_t *A);
/* ---- single-process multi-GPU mode ---------------------------------------
 * One caller thread, all GPUs of the box, through the same entry points as a
 * single GPU.  Replaces nothing in the reference (which is serial); the seam is
 * the block-row loop of composite_matvec_add
 * (src/matrix/sparse_matrix_composites.f90:1076-1100) behind the unchanged
 * solver interface (src/linear_operator/linear_operator_interface.f90:108-123):
 * a serial `use sigma` caller keeps calling A%matvec(x, y) and
 * solver%solve(A, x, b) on whole vectors.
 *
 *  sigb_mgpu_init(ndev)   ndev <= 0: every visible GPU.  One worker thread per GPU
 *                         inside the library, peer access between all pairs, the
 *                         peer-memory transport of the row-sharded operators (no
 *                         NCCL, no IPC handles).  SIGB_ERR_COMM when two of the
 *                         GPUs cannot address each other: there is no fallback.
 *  sigb_mgpu_csr_create   whole n x n CSR pattern in (1-based ptr / node as the
 *                         Fortran holds them): partition into row blocks balanced
 *                         by stored entries, halo lists and send lists are derived
 *                         here (bit-identical to sigb_partition_rows /
 *                         sigb_halo_build) and one row block goes to every GPU.
 * The handle is then used with WHOLE host arrays in sigb_matrix_set_values,
 * sigb_matvec, sigb_matvec_add, sigb_matrix_get_dims, sigb_solver_setup (cg,
 * bicgstab, jacobi), sigb_solver_solve (pc: jacobi), sigb_solver_get_info,
 * sigb_solver_get_vector, sigb_lanczos, sigb_eigensolve and the destroy calls; the
 * *_dev entry points, matvec_t, copies, expressions, ldu and the generalized
 * eigensolvers return SIGB_ERR_UNSUPPORTED for it. */
SIGB_API int sigb_mgpu_init(int ndev);
SIGB_API int sigb_mgpu_finalize(void);
SIGB_API int sigb_mgpu_device_count(int *ndev);
SIGB_API int sigb_mgpu_csr_create(int32_t n, const int32_t *ptr1, const int32_t *node1,
                                  sigb_matrix_t *A);

/* Halo / send lists of a distributed matrix (index parity checks). */
SIGB_API int sigb_dist_get_halo(sigb_matrix_t A, int32_t *nhalo, int32_t *halo);

#ifdef __cplusplus
}
#endif
#endif /* SIGMA_B200_H */
