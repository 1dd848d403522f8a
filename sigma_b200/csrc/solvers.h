// solvers.h -- solver handle and the driver entry points shared by
// solvers.cu, api.cu and comm.cu.
#pragma once

#include "internal.h"

namespace sigb {
struct KState;
struct LduInfo;   // ldu.cu
enum SolverKind { S_CG = 0, S_BICGSTAB = 1, S_JACOBI = 2, S_LDU = 3 };
}  // namespace sigb

struct sigb_solver_s {
    int kind = 0;
    double tol = 1e-16;        // cg_set_params default, cg_solvers.f90:106
    bool params_set = false;
    bool initialized = false;  // work vectors allocated (cg_setup :78-81)
    int64_t cap = -1;          // safety cap per solve (not in the reference)
    int persistent = -1;       // CG loop form: -1 library's choice, 0 kernel per phase, 1 one persistent kernel
    int strict_order = 0;      // parity aid: dot products summed strictly left to right, like the serial reference
    int32_t nn = 0;            // solver%nn: owned rows
    int64_t nvec = 0;          // allocated length of each work vector (nn + halo room)
    int64_t iterations = 0;    // solver%iterations (accumulates until the next setup)
    double res2 = 0.0;
    int capped = 0;
    double *work = nullptr;    // cg: p,q,r,z ; bicgstab: p,q,r,r0,v,s,t,z ; jacobi: idiag
    int nwork = 0;
    sigb::KState *state = nullptr;       // device
    sigb::KState *state_host = nullptr;  // pinned
    double *xb = nullptr;      // device staging for host-pointer solves (x | b)
    int64_t xb_len = 0;
    sigb_matrix_t A = nullptr; // operator given to setup
    // persistent CG kernel (cg_persistent.cu)
    unsigned long long *bar = nullptr;   // grid barrier counter
    double *pers_partials = nullptr;     // 2 buffers x 2 values x kMaxGrid CTA partial sums
    sigb::LduInfo *ldu = nullptr;        // sparse_ldu_solver: factors and level schedules (ldu.cu)
    std::vector<sigb_solver_s *> sub;    // single-process multi-GPU mode: the same solver on every row block (mgpu.cu)
};

namespace sigb {

// ILDU(0) (ldu.cu): setup = pattern + schedules once, numeric factorisation every call;
// apply = x <- (I+U)^-1 D^-1 (I+L)^-1 b, a no-op when *skip_flag != 0
int ldu_setup_dev(sigb_solver_t s, sigb_matrix_t A);
int ldu_apply_dev(sigb_solver_t s, double *x, const double *b, const int *skip_flag);
void ldu_destroy_dev(sigb_solver_t s);
int ldu_sizes(sigb_solver_t s, int32_t *n, int64_t *nL, int64_t *nU, int32_t *nflev, int32_t *nblev);
int ldu_read(sigb_solver_t s, int32_t *Lptr, int32_t *Lnode, double *Lval, int32_t *Uptr, int32_t *Unode,
             double *Uval, double *D);
int jacobi_setup_dev(sigb_solver_t s, sigb_matrix_t A);
int jacobi_apply_dev(sigb_solver_t s, double *x, const double *b);
int cg_solve_dev(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc);
int bicgstab_solve_dev(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b,
                       sigb_solver_t pc);
int lanczos_dev(sigb_matrix_t A, int32_t n, const double *q1, uint64_t seed, int64_t row_offset,
                double *T, double *Q, double *w, KState *st);
int generalized_lanczos_dev(sigb_matrix_t A, sigb_matrix_t B, sigb_solver_t bs, sigb_solver_t bpc, int32_t n,
                            const double *q1, uint64_t seed, int64_t row_offset, double *T, double *Q, double *w,
                            double *v, double *zbuf, KState *st);
int tridiag_eig_host(int n, double *d, double *e, double *Z);
int ritz_vectors_dev(sigb_matrix_t A, double *V, double *V2, const double *Qm_dev, int64_t nr, int32_t n,
                     double *first_row_dev);
size_t kstate_bytes();

// ---- single-process multi-GPU mode (mgpu.cu) --------------------------------
int mgpu_solver_setup(sigb_solver_t s, sigb_matrix_t A);
int mgpu_solver_solve(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc);
int mgpu_solver_get_vector(sigb_solver_t s, const char *name, double *out);
int mgpu_lanczos(sigb_matrix_t A, int32_t n, const double *q1, uint64_t seed, double *T, double *Q, double *lambda,
                 bool ritz);
void mgpu_solver_free(sigb_solver_t s);

// ---- persistent cooperative CG kernel (cg_persistent.cu) -------------------
struct PersistComm {          // all-reduce endpoints of a row-sharded operator
    void *red = nullptr;      // RedWin* of this rank
    void *peer_red[kMaxRanks] = {};
    int me = 0, nranks = 1;
};
struct CsrKernelArgs;
int cg_persistent_run(sigb_solver_t s, const CsrView &V, const double *val, const DotSpec &halo, double *x,
                      const double *b, double *p, double *q, double *r, double *z, const double *idiag, int64_t n,
                      const PersistComm &pcomm, long long max_iters);
// Row-sharded operators: fills the all-reduce endpoints and the halo spec the
// persistent kernel needs; *eligible = false when the operator uses a transport
// the kernel cannot drive (NCCL).
int dist_persist_info(sigb_matrix_t A, PersistComm *pc, DotSpec *halo, bool *eligible);

// y = A x (MODE_SET) with fused dots.  For row-sharded operators this also
// performs the halo exchange; x_has_halo says x has room for (and may receive)
// the halo entries behind its owned part.
int solver_matvec(sigb_matrix_t A, const double *x, double *y, const DotSpec &dot,
                  bool x_has_halo);
// Sum `count` contiguous device doubles over all ranks (no-op on one GPU).
// skip_flag: device int; when non-zero at execution time the reduction is a
// no-op (iterations launched past the stopping test), on every rank alike -- on the
// peer-memory transport.  On the NCCL transport the collective cannot be skipped by a
// device flag (every rank has to enter it), so past the latch it sums the already
// reduced words once more: the Krylov scalars (pq, rr, rho ...) are UNDEFINED after
// the latch there.  Nothing reads them: iters, final_res2, capped and done[] are
// latched by the kernel that evaluates the stopping test, a new solve starts from
// push_state, and the persistent kernel (the only form that resumes from rr) is not
// used with that transport.
bool dist_red_fuse(sigb_matrix_t A, RedFuse *rf);   // peer-memory transport: reductions finished inside their producers (comm.cu)
int dist_allreduce(sigb_matrix_t A, double *vals, int count, const int *skip_flag = nullptr);
int dist_allreduce2(sigb_matrix_t A, double *a, double *b, const int *skip_flag = nullptr);
// extra elements a work vector needs behind its owned part (0 on one GPU)
int64_t dist_halo_len(sigb_matrix_t A);

}  // namespace sigb
