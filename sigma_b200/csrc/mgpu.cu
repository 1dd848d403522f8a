// mgpu.cu -- single-process multi-GPU mode: ONE caller thread, all GPUs of the box.
//
// The reference is serial: a caller holds one csr_matrix and calls A%matvec(x, y) and
// solver%solve(A, x, b) on whole vectors (linear_operator_interface.f90:108-123,185-208).  The seam
// for several GPUs is the block-row loop of composite_matvec_add
// (sparse_matrix_composites.f90:1076-1100, "This loop can be parallelized" :1086): rows are
// independent, so the operator is split into contiguous row blocks, one per GPU.  Round 1 reached
// that only through one PROCESS per GPU with the index plan exchanged by the host program
// (sigma_b200/distributed.py); here the same is done behind the unchanged C-ABI:
//
//   sigb_mgpu_init(ndev)            one worker thread per GPU (each with its own context of this
//                                   library: device, stream, scratch), peer access enabled between
//                                   all pairs, an in-process rank group instead of NCCL / IPC handles
//   sigb_mgpu_csr_create(n, ptr, node, &A)
//                                   partition (rows balanced by stored entries), halo lists, send
//                                   lists -- all derived from the pattern in-library, bit-identical to
//                                   the per-process plan (sigb_partition_rows, sigb_halo_build) -- and
//                                   one row-sharded operator per GPU behind ONE sigb_matrix_t
//   sigb_matrix_set_values / sigb_matvec / sigb_matvec_add / sigb_solver_setup / sigb_solver_solve /
//   sigb_solver_get_info / sigb_solver_get_vector / *_destroy
//                                   take that handle with WHOLE host arrays; every worker serves its
//                                   row block through the per-GPU code paths (communication CTAs,
//                                   persistent CG kernel, in-kernel all-reduces) unchanged.
// The data path between the GPUs is the peer-memory transport of comm.cu; the host threads only
// launch and wait.
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "dist.h"
#include "internal.h"
#include "solvers.h"

namespace sigb {

namespace {

struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;     // set by the caller, cleared by the worker
    bool has_job = false, quit = false, done = false;
    int status = SIGB_OK;
    std::string error;
};

struct Mgpu {
    bool on = false;
    int ndev = 0;
    std::vector<Worker *> w;
    LocalGroup *grp = nullptr;
    std::vector<sigb_comm_t> comm;
};

Mgpu &mg()
{
    static Mgpu m;
    return m;
}

void worker_main(Worker *w)
{
    for (;;) {
        std::function<int()> job;
        {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) return;
            job = w->job;
        }
        const int rc = job();
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->status = rc;
            w->error = rc == SIGB_OK ? "" : sigb_last_error();
            w->has_job = false;
            w->done = true;
        }
        w->cv.notify_all();
    }
}

// fn(rank) on every worker, concurrently; the first failing rank's status and message win
int run_all(const std::function<int(int)> &fn)
{
    Mgpu &M = mg();
    for (int r = 0; r < M.ndev; r++) {
        Worker *w = M.w[(size_t)r];
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->job = [fn, r]() { return fn(r); };
            w->has_job = true;
            w->done = false;
        }
        w->cv.notify_all();
    }
    int rc = SIGB_OK;
    for (int r = 0; r < M.ndev; r++) {
        Worker *w = M.w[(size_t)r];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->status != SIGB_OK && rc == SIGB_OK) {
            rc = w->status;
            set_error("GPU %d: %s", r, w->error.c_str());
        }
    }
    return rc;
}

}  // namespace

// one row-sharded operator per GPU behind a single handle
struct MgpuMatrix {
    int32_t n = 0;
    int64_t nnz = 0;
    std::vector<int32_t> part;       // row offsets, ndev + 1
    std::vector<int64_t> ent_off;    // offsets of the row blocks in the stored-entry arrays, ndev + 1
    std::vector<sigb_matrix_t> shard;
};

bool mgpu_active() { return mg().on; }
int mgpu_devices() { return mg().ndev; }

void mgpu_matrix_free(sigb_matrix_t A)
{
    MgpuMatrix *G = A->mg;
    if (!G) return;
    run_all([&](int r) { return G->shard[(size_t)r] ? sigb_matrix_destroy(G->shard[(size_t)r]) : SIGB_OK; });
    delete G;
    A->mg = nullptr;
}

int mgpu_set_values(sigb_matrix_t A, const double *val, int64_t count)
{
    MgpuMatrix *G = A->mg;
    SIGB_REQUIRE(count == G->nnz, SIGB_ERR_ARG, "sigb_matrix_set_values: got %lld values, graph has %lld edges",
                 (long long)count, (long long)G->nnz);
    return run_all([&](int r) {
        return sigb_matrix_set_values(G->shard[(size_t)r], val + G->ent_off[(size_t)r],
                                      G->ent_off[(size_t)r + 1] - G->ent_off[(size_t)r]);
    });
}

int mgpu_matvec(sigb_matrix_t A, int trans, const double *x, double *y, bool add)
{
    MgpuMatrix *G = A->mg;
    SIGB_REQUIRE(!trans, SIGB_ERR_UNSUPPORTED, "sigb_matvec: matvec_t of a multi-GPU operator is not available");
    return run_all([&](int r) {
        const int32_t lo = G->part[(size_t)r];
        return add ? sigb_matvec_add(G->shard[(size_t)r], 0, x + lo, y + lo) : sigb_matvec(G->shard[(size_t)r], 0, x + lo, y + lo);
    });
}

int64_t mgpu_nnz(sigb_matrix_t A) { return A->mg->nnz; }

// solver%setup(A) on every GPU: one solver of the same kind per row block
int mgpu_solver_setup(sigb_solver_t s, sigb_matrix_t A)
{
    MgpuMatrix *G = A->mg;
    SIGB_REQUIRE(s->kind == S_CG || s->kind == S_BICGSTAB || s->kind == S_JACOBI, SIGB_ERR_UNSUPPORTED,
                 "sigb_solver_setup: cg, bicgstab and jacobi are available on a multi-GPU operator");
    const int ndev = mg().ndev;
    if ((int)s->sub.size() != ndev) {
        SIGB_REQUIRE(s->sub.empty(), SIGB_ERR_STATE, "sigb_solver_setup: solver belongs to another multi-GPU configuration");
        s->sub.assign((size_t)ndev, nullptr);
    }
    SIGB_CHECK(run_all([&](int r) {
        sigb_solver_t &q = s->sub[(size_t)r];
        if (!q) {
            int rc = s->kind == S_CG ? sigb_cg_create(s->tol, &q)
                                     : (s->kind == S_BICGSTAB ? sigb_bicgstab_create(s->tol, &q) : sigb_jacobi_create(&q));
            if (rc != SIGB_OK) return rc;
        }
        return sigb_solver_setup(q, G->shard[(size_t)r]);
    }));
    s->nn = A->nrow;
    s->iterations = 0;
    s->A = A;
    s->initialized = true;
    return SIGB_OK;
}

int mgpu_solver_solve(sigb_solver_t s, sigb_matrix_t A, double *x, const double *b, sigb_solver_t pc)
{
    MgpuMatrix *G = A->mg;
    SIGB_REQUIRE((int)s->sub.size() == mg().ndev && s->initialized, SIGB_ERR_STATE,
                 "sigb_solver_solve: solver%%setup(A) has not been called");
    SIGB_REQUIRE(s->nn == A->nrow, SIGB_ERR_ARG, "sigb_solver_solve: solver was set up for nn = %d, operator has %d rows",
                 s->nn, A->nrow);
    if (pc)
        SIGB_REQUIRE(pc->kind == S_JACOBI && (int)pc->sub.size() == mg().ndev && pc->initialized, SIGB_ERR_STATE,
                     "sigb_solver_solve: the preconditioner of a multi-GPU solve is jacobi, set up on the same operator");
    SIGB_CHECK(run_all([&](int r) {
        sigb_solver_t q = s->sub[(size_t)r];
        q->tol = s->tol;
        q->cap = s->cap;
        q->persistent = s->persistent;
        const int32_t lo = G->part[(size_t)r];
        return sigb_solver_solve(q, G->shard[(size_t)r], x + lo, b + lo, pc ? pc->sub[(size_t)r] : nullptr);
    }));
    // every rank computes bit-identical scalars and stops at the same iteration
    sigb_solver_t q0 = s->sub[0];
    s->iterations = q0->iterations;
    s->res2 = q0->res2;
    s->capped = q0->capped;
    return SIGB_OK;
}

// lanczos / eigensolve on every GPU: each worker runs the row-sharded form on its block (identical T and Ritz
// values on every rank: the dot products are all-reduced bit-identically), writing its rows of the basis into
// a buffer of its own; the columns are then gathered into the caller's nrow x n column-major array.
int mgpu_lanczos(sigb_matrix_t A, int32_t n, const double *q1, uint64_t seed, double *T, double *Q, double *lambda,
                 bool ritz)
{
    MgpuMatrix *G = A->mg;
    const int ndev = mg().ndev;
    std::vector<std::vector<double>> Qr((size_t)ndev), Tr((size_t)ndev);
    SIGB_CHECK(run_all([&](int r) {
        const int32_t lo = G->part[(size_t)r], nloc = G->part[(size_t)r + 1] - lo;
        Qr[(size_t)r].assign((size_t)std::max(nloc, 1) * n, 0.0);
        Tr[(size_t)r].assign((size_t)3 * n, 0.0);
        sigb_matrix_t S = G->shard[(size_t)r];
        return ritz ? sigb_eigensolve(S, n, q1 ? q1 + lo : nullptr, seed, Tr[(size_t)r].data(), Qr[(size_t)r].data())
                    : sigb_lanczos(S, n, q1 ? q1 + lo : nullptr, seed, Tr[(size_t)r].data(), Qr[(size_t)r].data());
    }));
    if (ritz) {
        if (lambda) memcpy(lambda, Tr[0].data(), sizeof(double) * (size_t)n);     // sigb_eigensolve returns lambda there
    } else if (T) {
        memcpy(T, Tr[0].data(), sizeof(double) * 3 * (size_t)n);
    }
    if (Q)
        for (int r = 0; r < ndev; r++) {
            const int32_t lo = G->part[(size_t)r], nloc = G->part[(size_t)r + 1] - lo;
            for (int32_t j = 0; j < n; j++)
                memcpy(Q + (size_t)j * G->n + lo, Qr[(size_t)r].data() + (size_t)j * nloc, sizeof(double) * (size_t)nloc);
        }
    return SIGB_OK;
}

int mgpu_solver_get_vector(sigb_solver_t s, const char *name, double *out)
{
    MgpuMatrix *G = s->A ? s->A->mg : nullptr;
    SIGB_REQUIRE(G && !s->sub.empty(), SIGB_ERR_STATE, "sigb_solver_get_vector: no multi-GPU setup");
    return run_all([&](int r) { return sigb_solver_get_vector(s->sub[(size_t)r], name, out + G->part[(size_t)r]); });
}

void mgpu_solver_free(sigb_solver_t s)
{
    if (s->sub.empty()) return;
    if (mg().on && (int)s->sub.size() == mg().ndev)
        run_all([&](int r) { return sigb_solver_destroy(s->sub[(size_t)r]); });
    s->sub.clear();
}

int64_t mgpu_launch_count()
{
    if (!mg().on) return 0;
    std::vector<int64_t> c((size_t)mg().ndev, 0);
    run_all([&](int r) { c[(size_t)r] = ctx().launches; return SIGB_OK; });
    int64_t total = 0;
    for (int64_t v : c) total += v;
    return total;
}

}  // namespace sigb

using namespace sigb;

extern "C" {

int sigb_mgpu_init(int ndev)
{
    Mgpu &M = mg();
    if (M.on) {
        SIGB_REQUIRE(ndev <= 0 || ndev == M.ndev, SIGB_ERR_STATE, "sigb_mgpu_init: already running on %d GPUs", M.ndev);
        return SIGB_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("sigb_mgpu_init: no CUDA device is usable (%s); this library has no CPU path", cudaGetErrorString(e));
        return SIGB_ERR_CUDA;
    }
    if (ndev <= 0) ndev = count;
    SIGB_REQUIRE(ndev <= count, SIGB_ERR_ARG, "sigb_mgpu_init: %d GPUs asked for, %d visible", ndev, count);
    SIGB_REQUIRE(ndev <= kMaxRanks, SIGB_ERR_UNSUPPORTED, "sigb_mgpu_init: at most %d GPUs (one box)", kMaxRanks);
    for (int a = 0; a < ndev; a++)
        for (int b = 0; b < ndev; b++) {
            if (a == b) continue;
            int can = 0;
            SIGB_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
            SIGB_REQUIRE(can, SIGB_ERR_COMM, "sigb_mgpu_init: GPU %d cannot address the memory of GPU %d (no NVLink / peer access); "
                         "there is no fallback transport in single-process mode", a, b);
        }
    M.ndev = ndev;
    M.grp = local_group_create(ndev);
    M.w.clear();
    for (int r = 0; r < ndev; r++) {
        Worker *w = new Worker();
        w->th = std::thread(worker_main, w);
        M.w.push_back(w);
    }
    M.comm.assign((size_t)ndev, nullptr);
    M.on = true;
    const int rc = run_all([&](int r) {
        SIGB_CHECK(sigb_init(r));          // this thread's context of the library, bound to GPU r
        for (int q = 0; q < ndev; q++) {
            if (q == r) continue;
            const cudaError_t pe = cudaDeviceEnablePeerAccess(q, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(pe, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
        local_group_barrier(M.grp);        // every device can be written before the first window is shared
        return comm_create_local(M.grp, r, ndev, &M.comm[(size_t)r]);
    });
    if (rc != SIGB_OK) {
        sigb_mgpu_finalize();
        return rc;
    }
    return SIGB_OK;
}

int sigb_mgpu_finalize(void)
{
    Mgpu &M = mg();
    if (!M.on) return SIGB_OK;
    run_all([&](int r) {
        if (M.comm[(size_t)r]) sigb_comm_destroy(M.comm[(size_t)r]);
        M.comm[(size_t)r] = nullptr;
        return sigb_finalize();
    });
    for (Worker *w : M.w) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
        }
        w->cv.notify_all();
        w->th.join();
        delete w;
    }
    M.w.clear();
    local_group_destroy(M.grp);
    M.grp = nullptr;
    M.on = false;
    M.ndev = 0;
    return SIGB_OK;
}

int sigb_mgpu_device_count(int *ndev)
{
    SIGB_REQUIRE(ndev, SIGB_ERR_ARG, "sigb_mgpu_device_count: null output");
    *ndev = mg().on ? mg().ndev : 0;
    return SIGB_OK;
}

int sigb_mgpu_csr_create(int32_t n, const int32_t *ptr1, const int32_t *node1, sigb_matrix_t *out)
{
    Mgpu &M = mg();
    SIGB_REQUIRE(M.on, SIGB_ERR_STATE, "sigb_mgpu_csr_create: call sigb_mgpu_init first");
    SIGB_REQUIRE(n >= 0 && ptr1 && out, SIGB_ERR_ARG, "sigb_mgpu_csr_create: bad argument");
    SIGB_REQUIRE(ptr1[0] == 1, SIGB_ERR_ARG, "sigb_mgpu_csr_create: ptr is 1-based (ptr(1) = 1)");
    for (int32_t i = 0; i < n; i++)
        SIGB_REQUIRE(ptr1[i] <= ptr1[i + 1], SIGB_ERR_ARG, "sigb_mgpu_csr_create: ptr must be monotone (row %d)", i + 1);
    const int64_t ne = (int64_t)ptr1[n] - 1;
    SIGB_REQUIRE(ne == 0 || node1, SIGB_ERR_ARG, "sigb_mgpu_csr_create: null node array");
    // (checked here for all row blocks at once: the per-GPU creation below is collective, and a rank that
    //  returned early on bad input would leave the others waiting)
    for (int64_t k = 0; k < ne; k++)
        SIGB_REQUIRE(node1[k] >= 1 && node1[k] <= n, SIGB_ERR_ARG, "sigb_mgpu_csr_create: column id %d of entry %lld outside 1..%d",
                     node1[k], (long long)k + 1, n);
    const int P = M.ndev;
    MgpuMatrix *G = new MgpuMatrix();
    G->n = n;
    G->nnz = ne;
    G->part.assign((size_t)P + 1, 0);
    G->ent_off.assign((size_t)P + 1, 0);
    G->shard.assign((size_t)P, nullptr);
    int rc = sigb_partition_rows(n, ptr1, P, G->part.data());
    if (rc != SIGB_OK) { delete G; return rc; }
    for (int r = 0; r <= P; r++) G->ent_off[(size_t)r] = (int64_t)ptr1[G->part[(size_t)r]] - 1;

    // halo list of every rank (what it reads and does not own) -> send lists (their mirror image on
    // the owners): pure index work on the pattern, the same functions the per-process plan uses
    std::vector<std::vector<int32_t>> halo((size_t)P), send_rows((size_t)P), send_cnt((size_t)P);
    for (int r = 0; r < P && rc == SIGB_OK; r++) {
        const int32_t lo = G->part[(size_t)r], hi = G->part[(size_t)r + 1];
        const int64_t cnt = G->ent_off[(size_t)r + 1] - G->ent_off[(size_t)r];
        std::vector<int32_t> h((size_t)std::max<int64_t>(cnt, 1)), local((size_t)std::max<int64_t>(cnt, 1));
        int32_t nh = 0;
        rc = sigb_halo_build(lo, hi, ptr1 + lo, node1 ? node1 + G->ent_off[(size_t)r] : nullptr, h.data(), &nh, local.data());
        h.resize((size_t)nh);
        halo[(size_t)r].swap(h);
    }
    if (rc != SIGB_OK) { delete G; return rc; }
    for (int r = 0; r < P; r++) {          // rank r sends to q the rows of q's halo that r owns, grouped by q ascending
        const int32_t lo = G->part[(size_t)r], hi = G->part[(size_t)r + 1];
        send_cnt[(size_t)r].assign((size_t)P, 0);
        for (int q = 0; q < P; q++) {
            if (q == r) continue;
            for (int32_t c : halo[(size_t)q])
                if (c > lo && c <= hi) {
                    send_rows[(size_t)r].push_back(c - lo);
                    send_cnt[(size_t)r][(size_t)q]++;
                }
        }
    }
    rc = run_all([&](int r) {
        const int32_t lo = G->part[(size_t)r];
        return sigb_dist_csr_create(M.comm[(size_t)r], n, G->part.data(), ptr1 + lo,
                                    node1 ? node1 + G->ent_off[(size_t)r] : nullptr, send_cnt[(size_t)r].data(),
                                    send_rows[(size_t)r].empty() ? nullptr : send_rows[(size_t)r].data(),
                                    &G->shard[(size_t)r]);
    });
    sigb_matrix_t A = new sigb_matrix_s();
    A->mg = G;
    A->nrow = n;
    A->ncol = n;
    if (rc != SIGB_OK) {
        mgpu_matrix_free(A);
        delete A;
        return rc;
    }
    *out = A;
    return SIGB_OK;
}

}  // extern "C"
