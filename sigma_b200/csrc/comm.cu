// comm.cu -- multi-GPU communicator and row-sharded (distributed) operators.
//
// Not in the reference, which is serial; the seam is the block-row loop of
// composite_matvec_add (src/matrix/sparse_matrix_composites.f90:1076-1100,
// "This loop can be parallelized" :1086) with its x(j1:j2) / y(i1:i2) slices.
//
// One process per GPU.  Rank r owns the contiguous rows [part[r], part[r+1]) of
// a square global operator; its block is stored as a local CSR whose columns
// are renumbered [owned | halo] (partition.cpp).  One SpMV is
//     publish the owned entries other ranks need
//     interior tiles (no halo column)          -- overlaps the transfer
//     boundary tiles, gathering halo columns from the landing buffer
// and every Krylov dot product is completed by an all-reduce of 1-3 doubles.
//
// Two transports:
//  * peer memory (default): every rank exposes a small window through CUDA IPC;
//    NVLink/NVSwitch makes it load/store addressable by all peers.  The push
//    kernel stores halo entries straight into the consumers' landing buffers
//    and publishes a sequence number; the boundary SpMV kernel itself waits on
//    those flags (kernels_spmv.cu) and acknowledges consumption.  All-reduces
//    are one warp: store the partial into every peer's inbox, wait for the P
//    inbox entries, add them in rank order -- every rank gets bit-identical
//    sums, so all ranks take the same stopping decision.  A few microseconds
//    per collective instead of a library launch per message.
//  * NCCL (SIGB_TRANSPORT=nccl, or when IPC mapping is unavailable): grouped
//    ncclSend/ncclRecv on an auxiliary stream + ncclAllReduce.
// NCCL is always used for bootstrap (exchanging IPC handles and layouts).
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "device_utils.cuh"
#include "dist.h"
#include "solvers.h"

namespace sigb {

// Ranks that live in ONE process (single-process multi-device mode, mgpu.cu: one host thread per
// GPU): the bootstrap exchanges go through host memory and the windows are shared as plain
// pointers (peer access is enabled between the devices) -- no NCCL, no IPC handles.
struct LocalGroup {
    int nranks = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    unsigned generation = 0;
    std::vector<std::vector<char>> slot;   // one per rank
    void barrier()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned gen = generation;
        if (++arrived == nranks) {
            arrived = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen; });
        }
    }
};

}  // namespace sigb

struct sigb_comm_s {
    sigb::LocalGroup *grp = nullptr;   // non-null: the ranks are threads of this process
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;   // NCCL halo exchange stream
    cudaEvent_t ev_pack = nullptr, ev_halo = nullptr;
    bool p2p = false;
    sigb::RedWin *red = nullptr;                     // this rank's inbox
    sigb::RedWin *peer_red[sigb::kMaxRanks] = {};    // peers' inboxes (peer-mapped)
    char *stage = nullptr;                           // device staging for bootstrap all-gathers
    size_t stage_bytes = 0;
};

namespace sigb {

#define SIGB_NCCL(call)                                                                  \
    do {                                                                                 \
        ncclResult_t r__ = (call);                                                       \
        if (r__ != ncclSuccess) {                                                        \
            set_error("NCCL error %d (%s) in %s at %s:%d", (int)r__, ncclGetErrorString(r__), #call, \
                      __FILE__, __LINE__);                                               \
            return SIGB_ERR_COMM;                                                        \
        }                                                                                \
    } while (0)

struct DistInfo {
    sigb_comm_t comm = nullptr;
    int32_t n_global = 0, lo = 0, hi = 0, nloc = 0, nhalo = 0;
    std::vector<int32_t> halo_host;              // sorted unique global column ids (1-based)
    std::vector<int> send_cnt, send_off, recv_cnt, recv_off;
    int total_send = 0;
    int32_t *send_rows = nullptr;   // device, 1-based local rows, grouped by destination
    double *sendbuf = nullptr;      // device, packed values (NCCL transport)
    double *halo = nullptr;         // device, landing buffer (NCCL transport)
    double *dot_tmp = nullptr;      // device, 2 doubles: interior partial sums
    // peer-memory transport
    HaloWin *win = nullptr;         // our window (flags + two landing buffers)
    int64_t stride = 0;             // entries between the two landing buffers
    HaloSync sync;                  // what boundary launches need
};

namespace {

// ---- bootstrap: all-gather of a few bytes per rank over NCCL ---------------
int allgather_bytes(sigb_comm_t c, const void *mine, size_t bytes, void *all)
{
    if (c->grp) {
        LocalGroup *g = c->grp;
        g->slot[(size_t)c->rank].assign((const char *)mine, (const char *)mine + bytes);
        g->barrier();
        for (int q = 0; q < c->nranks; q++) memcpy((char *)all + (size_t)q * bytes, g->slot[(size_t)q].data(), bytes);
        g->barrier();   // nobody overwrites its slot before everybody has read it
        return SIGB_OK;
    }
    const size_t need = bytes * (size_t)(c->nranks + 1);
    if (c->stage_bytes < need) {
        cudaFree(c->stage);
        SIGB_CUDA(cudaMalloc((void **)&c->stage, need));
        c->stage_bytes = need;
    }
    cudaStream_t st = ctx().stream;
    char *send = c->stage, *recv = c->stage + bytes;
    SIGB_CUDA(cudaMemcpyAsync(send, mine, bytes, cudaMemcpyHostToDevice, st));
    SIGB_NCCL(ncclAllGather(send, recv, bytes, ncclChar, c->nccl, st));
    SIGB_CUDA(cudaMemcpyAsync(all, recv, bytes * c->nranks, cudaMemcpyDeviceToHost, st));
    SIGB_CUDA(cudaStreamSynchronize(st));
    return SIGB_OK;
}

// Expose `ptr` (a cudaMalloc allocation) to all ranks; peers[q] receives the
// address it is mapped at in this process.  Collective.
int share_window(sigb_comm_t c, void *ptr, void **peers)
{
    if (c->grp) {   // one address space, peer access enabled: the pointers themselves
        std::vector<void *> all((size_t)c->nranks);
        SIGB_CHECK(allgather_bytes(c, &ptr, sizeof(void *), all.data()));
        for (int q = 0; q < c->nranks; q++) peers[q] = all[(size_t)q];
        return SIGB_OK;
    }
    cudaIpcMemHandle_t mine;
    SIGB_CUDA(cudaIpcGetMemHandle(&mine, ptr));
    std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
    SIGB_CHECK(allgather_bytes(c, &mine, sizeof(mine), all.data()));
    for (int q = 0; q < c->nranks; q++) {
        if (q == c->rank) {
            peers[q] = ptr;
        } else {
            SIGB_CUDA(cudaIpcOpenMemHandle(&peers[q], all[q], cudaIpcMemLazyEnablePeerAccess));
        }
    }
    return SIGB_OK;
}

// ---- peer-memory all-reduce: one warp -----------------------------------------
struct RedArgs {
    double *vals[3];
    int count;
    RedWin *win;
    RedWin *peer[kMaxRanks];
    int me, nranks;
    const int *skip_flag;
    FaultBlock *fault;
};

// One warp.  Lane q < nranks stores this rank's partial sums into rank q's
// inbox and then polls the entry rank q stored into ours; lanes c < count add
// the nranks contributions in rank order, so every rank computes bit-identical
// totals and all ranks take the same stopping decision.  A slot is reused every
// kRedSlots reductions; an all-reduce is a full barrier, so no rank can be that
// far ahead.
__global__ void red_kernel(const __grid_constant__ RedArgs a)
{
    if (a.skip_flag != nullptr && *a.skip_flag != 0) return;
    const int lane = threadIdx.x;
    const unsigned long long s = a.win->red_seq + 1;
    const int slot = (int)(s & (kRedSlots - 1));
    const unsigned int flag = (unsigned int)s;
    if (lane < a.nranks) {
        RedEntry *e = a.peer[lane]->red[slot][a.me];
        for (int c = 0; c < a.count; c++) red_entry_store(&e[c], *a.vals[c], flag);
    }
    __syncwarp();
    for (int c = 0; c < a.count; c++) {
        const double v = warp_rank_sum(a.win, slot, c, a.nranks, flag, a.fault);   // rank order, all lanes alike
        if (lane == 0) *a.vals[c] = v;
    }
    __syncwarp();
    if (lane == 0) a.win->red_seq = s;
}

// ---- NCCL transport: pack into a contiguous send buffer -----------------------
__global__ void __launch_bounds__(kThreads)
pack_kernel(const double *__restrict__ x, const int32_t *__restrict__ rows1, int n,
            double *__restrict__ out, const int *skip_flag)
{
    if (skip_flag != nullptr && *skip_flag != 0) return;
    for (int k = blockIdx.x * kThreads + threadIdx.x; k < n; k += gridDim.x * kThreads)
        out[k] = x[rows1[k] - 1];
}

}  // namespace

int64_t dist_global_n(sigb_matrix_t A) { return A->dist ? A->dist->n_global : A->nrow; }
int64_t dist_row_offset(sigb_matrix_t A) { return A->dist ? A->dist->lo : 0; }
int64_t dist_halo_len(sigb_matrix_t) { return 0; }  // halos land in the operator's own buffers

int dist_persist_info(sigb_matrix_t A, PersistComm *pc, DotSpec *halo, bool *eligible)
{
    *pc = PersistComm();
    *halo = DotSpec();
    *eligible = true;
    DistInfo *D = A->dist;
    if (!D) return SIGB_OK;
    sigb_comm_t C = D->comm;
    halo->nloc = D->nloc;
    if (C->nranks == 1) return SIGB_OK;
    if (!C->p2p) { *eligible = false; return SIGB_OK; }   // NCCL transport: host-driven kernels only
    pc->red = C->red;
    for (int q = 0; q < kMaxRanks; q++) pc->peer_red[q] = C->peer_red[q];
    pc->me = C->rank;
    pc->nranks = C->nranks;
    if (D->total_send > 0 || D->nhalo > 0) halo->sync = &D->sync;
    return SIGB_OK;
}

static void free_dist(DistInfo *D)
{
    if (!D) return;
    cudaDeviceSynchronize();
    if (D->win) {
        if (!D->comm->grp)
            for (int q = 0; q < D->comm->nranks; q++)
                if (q != D->comm->rank && D->sync.peer[q]) cudaIpcCloseMemHandle(D->sync.peer[q]);
        cudaFree(D->win);
    }
    cudaFree(D->send_rows);
    cudaFree(D->sendbuf);
    cudaFree(D->halo);
    cudaFree(D->dot_tmp);
    delete D;
}

int dist_destroy(sigb_matrix_t A)
{
    free_dist(A->dist);
    A->dist = nullptr;
    return SIGB_OK;
}

static int allreduce_ptrs(sigb_comm_t C, double *const *vals, int count, const int *skip_flag)
{
    if (C->p2p) {
        RedArgs a;
        for (int c = 0; c < 3; c++) a.vals[c] = c < count ? vals[c] : nullptr;
        a.count = count;
        a.win = C->red;
        for (int q = 0; q < kMaxRanks; q++) a.peer[q] = C->peer_red[q];
        a.me = C->rank;
        a.nranks = C->nranks;
        a.skip_flag = skip_flag;
        a.fault = ctx().fault_dev;
        red_kernel<<<1, 32, 0, ctx().stream>>>(a);
        count_launch();
        SIGB_CUDA(cudaGetLastError());
        return SIGB_OK;
    }
    if (count > 1) SIGB_NCCL(ncclGroupStart());
    for (int c = 0; c < count; c++)
        SIGB_NCCL(ncclAllReduce(vals[c], vals[c], 1, ncclDouble, ncclSum, C->nccl, ctx().stream));
    if (count > 1) SIGB_NCCL(ncclGroupEnd());
    return SIGB_OK;
}

// Peer-memory transport: endpoints for kernels that finish their own reduction across the GPUs
// (the last CTA of the producing kernel stores the local sum into the peers' inboxes and adds the
// ranks' contributions in rank order -- device_utils.cuh grid_reduce); false = NCCL transport or a
// single rank, use dist_allreduce after the kernel.  Round-2 A/B at 2 GPUs, full size: 226.4 us per
// CG iteration against 233.2 us with the separate one-warp launches (profiles/r2_visit_b_2gpu_summary.txt).
bool dist_red_fuse(sigb_matrix_t A, RedFuse *rf)
{
    *rf = RedFuse();
    DistInfo *D = A->dist;
    if (!D || D->comm->nranks == 1 || !D->comm->p2p) return false;
    sigb_comm_t C = D->comm;
    rf->win = C->red;
    for (int q = 0; q < kMaxRanks; q++) rf->peer[q] = C->peer_red[q];
    rf->me = C->rank;
    rf->nranks = C->nranks;
    rf->fault = ctx().fault_dev;
    return true;
}

int dist_allreduce(sigb_matrix_t A, double *vals, int count, const int *skip_flag)
{
    DistInfo *D = A->dist;
    if (!D || D->comm->nranks == 1) return SIGB_OK;
    double *ptrs[3] = {vals, vals + 1, vals + 2};
    return allreduce_ptrs(D->comm, ptrs, count, skip_flag);
}

int dist_allreduce2(sigb_matrix_t A, double *a, double *b, const int *skip_flag)
{
    DistInfo *D = A->dist;
    if (!D || D->comm->nranks == 1) return SIGB_OK;
    double *ptrs[3] = {a, b, nullptr};
    return allreduce_ptrs(D->comm, ptrs, 2, skip_flag);
}

// y = A x (or y += A x) for the local row block; x holds the owned entries.
int dist_matvec(sigb_matrix_t A, const double *x, double *y, bool add, const DotSpec &dot_in, bool)
{
    DistInfo *D = A->dist;
    sigb_comm_t C = D->comm;
    const CsrView &V = A->g->stored;
    const SpmvMode mode = add ? MODE_ADD_AFTER : MODE_SET;
    cudaStream_t main = ctx().stream;
    DotSpec dot = dot_in;
    dot.nloc = D->nloc;

    const bool exchange = C->nranks > 1 && (D->total_send > 0 || D->nhalo > 0);
    const bool p2p = exchange && C->p2p;
    if (p2p) {
        // ONE kernel does push + interior + (wait) + boundary + acknowledge
        DotSpec db = dot;
        db.sync = &D->sync;
        return launch_csr_spmv(V, A->val, x, y, mode, db, 0, main, 0);
    } else if (exchange) {
        if (D->total_send > 0) {
            int grid = (D->total_send + kThreads - 1) / kThreads;
            grid = std::min(grid, ctx().num_sms * 4);
            pack_kernel<<<grid, kThreads, 0, main>>>(x, D->send_rows, D->total_send, D->sendbuf, dot.skip_flag);
            count_launch();
            SIGB_CUDA(cudaGetLastError());
        }
        SIGB_CUDA(cudaEventRecord(C->ev_pack, main));
        SIGB_CUDA(cudaStreamWaitEvent(C->stream, C->ev_pack, 0));
        SIGB_NCCL(ncclGroupStart());
        for (int q = 0; q < C->nranks; q++) {
            if (D->send_cnt[q] > 0)
                SIGB_NCCL(ncclSend(D->sendbuf + D->send_off[q], (size_t)D->send_cnt[q], ncclDouble, q, C->nccl, C->stream));
            if (D->recv_cnt[q] > 0)
                SIGB_NCCL(ncclRecv(D->halo + D->recv_off[q], (size_t)D->recv_cnt[q], ncclDouble, q, C->nccl, C->stream));
        }
        SIGB_NCCL(ncclGroupEnd());
        SIGB_CUDA(cudaEventRecord(C->ev_halo, C->stream));
    }

    if (V.n_boundary == 0) {
        // nothing here depends on a halo: one launch over all tiles
        if (exchange) SIGB_CUDA(cudaStreamWaitEvent(main, C->ev_halo, 0));
        return launch_csr_spmv(V, A->val, x, y, mode, dot, 0, main, 0);
    }
    // NCCL: interior tiles overlap the exchange; their dot partials wait in dot_tmp
    DotSpec di = dot;
    if (dot.ndot > 0) {
        di.out[0] = D->dot_tmp;
        di.out[1] = D->dot_tmp + 1;
    }
    if (V.n_interior > 0 || dot.ndot > 0)
        SIGB_CHECK(launch_csr_spmv(V, A->val, x, y, mode, di, 1, main, 0));
    if (exchange) SIGB_CUDA(cudaStreamWaitEvent(main, C->ev_halo, 0));
    DotSpec db = dot;
    db.halo = D->halo;
    if (dot.ndot > 0) {
        db.addend[0] = D->dot_tmp;
        db.addend[1] = dot.ndot > 1 ? D->dot_tmp + 1 : nullptr;
    }
    return launch_csr_spmv(V, A->val, x, y, mode, db, 2, main, 0);
}

// ---- ranks as threads of one process (mgpu.cu) ---------------------------------
LocalGroup *local_group_create(int nranks)
{
    LocalGroup *g = new LocalGroup();
    g->nranks = nranks;
    g->slot.resize((size_t)nranks);
    return g;
}
void local_group_destroy(LocalGroup *g) { delete g; }
void local_group_barrier(LocalGroup *g) { g->barrier(); }

// Collective over the threads of the group; the calling thread's device is its rank's GPU and peer
// access to the other devices has been enabled.  Peer-memory transport only.
int comm_create_local(LocalGroup *grp, int rank, int nranks, sigb_comm_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(grp && out && nranks >= 1 && nranks <= kMaxRanks && rank >= 0 && rank < nranks, SIGB_ERR_ARG,
                 "comm_create_local: bad argument");
    sigb_comm_t c = new sigb_comm_s();
    c->grp = grp;
    c->rank = rank;
    c->nranks = nranks;
    int ok = 1;
    if (nranks > 1) {
        if (cudaMalloc((void **)&c->red, sizeof(RedWin)) != cudaSuccess) ok = 0;
        if (ok && cudaMemset(c->red, 0, sizeof(RedWin)) != cudaSuccess) ok = 0;
        void *peers[kMaxRanks] = {};
        share_window(c, c->red, peers);
        std::vector<int> oks((size_t)nranks);
        allgather_bytes(c, &ok, sizeof(int), oks.data());
        for (int q = 0; q < nranks; q++) ok = ok && oks[(size_t)q];
        if (ok) {
            for (int q = 0; q < nranks; q++) c->peer_red[q] = (RedWin *)peers[q];
            c->p2p = true;
        }
    }
    if (!ok) {
        cudaGetLastError();
        cudaFree(c->red);
        delete c;
        set_error("comm_create_local: could not allocate the all-reduce window");
        return SIGB_ERR_CUDA;
    }
    *out = c;
    return SIGB_OK;
}

}  // namespace sigb

using namespace sigb;

extern "C" {

int sigb_comm_unique_id(void *unique_id)
{
    SIGB_REQUIRE(unique_id, SIGB_ERR_ARG, "sigb_comm_unique_id: null buffer");
    static_assert(sizeof(ncclUniqueId) <= SIGB_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    SIGB_NCCL(ncclGetUniqueId(&id));
    memset(unique_id, 0, SIGB_UNIQUE_ID_BYTES);
    memcpy(unique_id, &id, sizeof(id));
    return SIGB_OK;
}

int sigb_comm_create(const void *unique_id, int rank, int nranks, sigb_comm_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(unique_id && out && nranks >= 1 && rank >= 0 && rank < nranks, SIGB_ERR_ARG,
                 "sigb_comm_create: bad argument");
    SIGB_REQUIRE(nranks <= kMaxRanks, SIGB_ERR_UNSUPPORTED, "sigb_comm_create: at most %d ranks (one box)", kMaxRanks);
    sigb_comm_t c = new sigb_comm_s();
    struct CommGuard {            // an early return releases what has been built so far
        sigb_comm_t c;
        ~CommGuard() { if (c) sigb_comm_destroy(c); }
    } guard{c};
    c->rank = rank;
    c->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    SIGB_NCCL(ncclCommInitRank(&c->nccl, nranks, id, rank));
    SIGB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SIGB_CUDA(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
    SIGB_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));

    // peer-memory transport: share the all-reduce inbox; every rank must succeed
    const char *tr = getenv("SIGB_TRANSPORT");
    int want = (nranks > 1 && !(tr && !strcmp(tr, "nccl"))) ? 1 : 0;
    if (want) {
        int ok = 1;
        if (cudaMalloc((void **)&c->red, sizeof(RedWin)) != cudaSuccess) ok = 0;
        if (ok) cudaMemset(c->red, 0, sizeof(RedWin));
        void *peers[kMaxRanks] = {};
        if (ok && share_window(c, c->red, peers) != SIGB_OK) ok = 0;
        cudaGetLastError();
        std::vector<int> oks((size_t)nranks);
        SIGB_CHECK(allgather_bytes(c, &ok, sizeof(int), oks.data()));
        for (int q = 0; q < nranks; q++) ok = ok && oks[q];
        if (ok) {
            for (int q = 0; q < nranks; q++) c->peer_red[q] = (RedWin *)peers[q];
            c->p2p = true;
        }
    }
    guard.c = nullptr;
    *out = c;
    return SIGB_OK;
}

int sigb_comm_destroy(sigb_comm_t c)
{
    if (!c) return SIGB_OK;
    cudaDeviceSynchronize();
    if (c->red) {
        if (!c->grp)
            for (int q = 0; q < c->nranks; q++)
                if (q != c->rank && c->peer_red[q]) cudaIpcCloseMemHandle(c->peer_red[q]);
        cudaFree(c->red);
    }
    cudaFree(c->stage);
    if (c->nccl) ncclCommDestroy(c->nccl);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->ev_pack) cudaEventDestroy(c->ev_pack);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    delete c;
    return SIGB_OK;
}

int sigb_comm_info(sigb_comm_t c, int *rank, int *nranks, int *peer_access)
{
    SIGB_REQUIRE(c, SIGB_ERR_ARG, "sigb_comm_info: null communicator");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    if (peer_access) *peer_access = c->p2p ? 1 : 0;
    return SIGB_OK;
}

int sigb_dist_csr_create(sigb_comm_t comm, int32_t n_global, const int32_t *part,
                         const int32_t *ptr_blk1, const int32_t *node_glob1,
                         const int32_t *send_counts, const int32_t *send_rows1, sigb_matrix_t *out)
{
    SIGB_CHECK(require_init());
    SIGB_REQUIRE(comm && part && ptr_blk1 && out && send_counts, SIGB_ERR_ARG, "sigb_dist_csr_create: bad argument");
    const int P = comm->nranks, me = comm->rank;
    SIGB_REQUIRE(part[0] == 0 && part[P] == n_global, SIGB_ERR_ARG, "sigb_dist_csr_create: part must span 0..n_global");
    for (int q = 0; q < P; q++)
        SIGB_REQUIRE(part[q] <= part[q + 1], SIGB_ERR_ARG, "sigb_dist_csr_create: part must be monotone (part[%d] = %d > part[%d] = %d)",
                     q, part[q], q + 1, part[q + 1]);
    const int32_t lo = part[me], hi = part[me + 1], nloc = hi - lo;
    for (int32_t i = 0; i < nloc; i++)
        SIGB_REQUIRE(ptr_blk1[i] <= ptr_blk1[i + 1], SIGB_ERR_ARG, "sigb_dist_csr_create: ptr must be monotone (row %d)", i + 1);
    const int64_t ne = (int64_t)ptr_blk1[nloc] - ptr_blk1[0];
    SIGB_REQUIRE(ne == 0 || node_glob1, SIGB_ERR_ARG, "sigb_dist_csr_create: null node array");
    for (int64_t k = 0; k < ne; k++)
        SIGB_REQUIRE(node_glob1[k] >= 1 && node_glob1[k] <= n_global, SIGB_ERR_ARG,
                     "sigb_dist_csr_create: column id %d of entry %lld outside 1..%d", node_glob1[k], (long long)k + 1, n_global);

    DistInfo *D = new DistInfo();
    // an early return (bad argument, CUDA failure) releases what has been built so far
    struct Guard {
        DistInfo *D;
        sigb_graph_t g;
        ~Guard() { free_dist(D); if (g) sigb_graph_release(g); }
    } guard{D, nullptr};
    D->comm = comm;
    D->n_global = n_global;
    D->lo = lo;
    D->hi = hi;
    D->nloc = nloc;

    // halo list + local numbering (bit-exact index work, partition.cpp)
    std::vector<int32_t> halo((size_t)std::max<int64_t>(ne, 1)), local((size_t)std::max<int64_t>(ne, 1));
    int32_t nhalo = 0;
    int rc = sigb_halo_build(lo, hi, ptr_blk1, node_glob1, halo.data(), &nhalo, local.data());
    if (rc != SIGB_OK) return rc;
    halo.resize((size_t)nhalo);
    D->nhalo = nhalo;
    D->halo_host = halo;
    D->recv_cnt.assign(P, 0);
    D->recv_off.assign(P, 0);
    D->send_cnt.assign(P, 0);
    D->send_off.assign(P, 0);
    for (int32_t h = 0, q = 0; h < nhalo; h++) {
        while (halo[h] > part[q + 1]) q++;   // owner of 1-based column c: part[q] < c <= part[q+1]
        D->recv_cnt[q]++;
    }
    for (int q = 1; q < P; q++) D->recv_off[q] = D->recv_off[q - 1] + D->recv_cnt[q - 1];
    int total_send = 0;
    for (int q = 0; q < P; q++) {
        SIGB_REQUIRE(send_counts[q] >= 0 && (q != me || send_counts[q] == 0), SIGB_ERR_ARG,
                     "sigb_dist_csr_create: bad send count for rank %d", q);
        D->send_cnt[q] = send_counts[q];
        D->send_off[q] = total_send;
        total_send += send_counts[q];
    }
    D->total_send = total_send;
    SIGB_REQUIRE(total_send == 0 || send_rows1, SIGB_ERR_ARG, "sigb_dist_csr_create: null send list");
    for (int k = 0; k < total_send; k++)
        SIGB_REQUIRE(send_rows1[k] >= 1 && send_rows1[k] <= nloc, SIGB_ERR_ARG,
                     "sigb_dist_csr_create: send row %d outside the owned block", send_rows1[k]);

    // local CSR mirror: ptr rebased to 1, columns in [owned | halo] numbering
    std::vector<int32_t> ptr((size_t)nloc + 1);
    for (int32_t i = 0; i <= nloc; i++) ptr[i] = ptr_blk1[i] - ptr_blk1[0] + 1;
    sigb_graph_t g = nullptr;
    rc = sigb_cs_graph_create(nloc, nloc + nhalo, ptr.data(), local.data(), SIGB_ROW, &g);
    if (rc != SIGB_OK) return rc;
    guard.g = g;

    // communication CTAs of the fused SpMV (spmv_device.cuh): ~8 entries per thread; at most 32 CTAs (7 % of the
    // persistent grid) -- a graph without locality sends nearly all of x (config 5: 17.5 M entries per rank)
    const int push_ctas = (P > 1 && comm->p2p && total_send > 0) ? std::min(32, (total_send + 8 * kThreads - 1) / (8 * kThreads)) : 0;
    // interior / boundary tile lists; the tiling is balanced over the compute CTAs of the persistent CG kernel
    std::vector<TileDesc> tiles, ti, tb;
    TileShape shape;
    build_tiles_balanced(ptr.data(), nloc, persistent_grid_ctas() - push_ctas, tiles, &shape);
    for (const TileDesc &t : tiles) {
        bool boundary = false;
        for (int32_t k = t.ks; k < t.ke && !boundary; k++) boundary = local[k] > nloc;
        (boundary ? tb : ti).push_back(t);
    }
    // the device tile table is re-ordered: interior tiles first, boundary last
    CsrView &V = g->stored;
    cudaStream_t st = ctx().stream;
    std::vector<TileDesc> ordered(ti);
    ordered.insert(ordered.end(), tb.begin(), tb.end());
    SIGB_CUDA(cudaStreamSynchronize(st));
    cudaFree(V.tiles);                           // the greedy table of sigb_cs_graph_create
    V.tiles = nullptr;
    SIGB_CHECK(upload_tiles(V, ordered));
    V.tile_nnz = shape.nnz;
    V.n_interior = (int32_t)ti.size();
    V.n_boundary = (int32_t)tb.size();
    V.tiles_interior = V.tiles;                  // views into the same table
    V.tiles_boundary = V.tiles + V.n_interior;
    SIGB_CUDA(cudaMalloc((void **)&D->send_rows, sizeof(int32_t) * std::max(total_send, 1)));
    SIGB_CUDA(cudaMalloc((void **)&D->dot_tmp, sizeof(double) * 2));
    SIGB_CUDA(cudaMemsetAsync(D->dot_tmp, 0, sizeof(double) * 2, st));
    if (total_send > 0)
        SIGB_CUDA(cudaMemcpyAsync(D->send_rows, send_rows1, sizeof(int32_t) * total_send, cudaMemcpyHostToDevice, st));

    if (comm->p2p) {
        // window = flags + two landing buffers; tell every rank where our slice
        // of ITS landing buffer starts (its recv offset for us) and its stride
        D->stride = ((int64_t)nhalo + 15) & ~15LL;
        const size_t bytes = sizeof(HaloWin) + sizeof(double) * 2 * (size_t)std::max<int64_t>(D->stride, 16);
        SIGB_CUDA(cudaMalloc((void **)&D->win, bytes));
        SIGB_CUDA(cudaMemsetAsync(D->win, 0, bytes, st));
        SIGB_CUDA(cudaStreamSynchronize(st));
        void *peers[kMaxRanks] = {};
        SIGB_CHECK(share_window(comm, D->win, peers));
        struct Layout { int64_t stride; int32_t recv_off[kMaxRanks]; int32_t recv_cnt[kMaxRanks]; } mine{}, all[kMaxRanks];
        mine.stride = D->stride;
        for (int q = 0; q < P; q++) { mine.recv_off[q] = D->recv_off[q]; mine.recv_cnt[q] = D->recv_cnt[q]; }
        SIGB_CHECK(allgather_bytes(comm, &mine, sizeof(Layout), all));
        D->sync.win = D->win;
        D->sync.me = me;
        D->sync.halo_base = reinterpret_cast<const double *>(D->win + 1);
        D->sync.halo_stride = D->stride;
        D->sync.send_rows = D->send_rows;
        D->sync.total_send = total_send;
        D->sync.push_ctas = push_ctas;
        // fewer than a quarter of the tiles are free of halo columns: nothing to hide the transfer behind,
        // every CTA of a stand-alone SpMV pushes its share first (HaloSync::push_all)
        D->sync.push_all = (push_ctas > 0 && ti.size() * 4 < tiles.size()) ? 1 : 0;
        const int forced = env_int("SIGB_PUSH_ALL", -1);      // A/B runs: 0 / 1 whatever the pattern
        if (forced >= 0 && push_ctas > 0) D->sync.push_all = forced ? 1 : 0;
        for (int q = 0; q < kMaxRanks; q++) {
            D->sync.peer[q] = q < P ? (HaloWin *)peers[q] : nullptr;
            D->sync.dst[q] = nullptr;
            D->sync.dst_stride[q] = 0;
        }
        for (int q = 0; q <= kMaxRanks; q++) D->sync.send_off[q] = q < P ? D->send_off[q] : total_send;
        for (int q = 0; q < P; q++) {
            if (D->recv_cnt[q] > 0) D->sync.src_mask |= 1u << q;
            if (D->send_cnt[q] > 0) {
                SIGB_REQUIRE(all[q].recv_cnt[me] == D->send_cnt[q], SIGB_ERR_ARG,
                             "sigb_dist_csr_create: rank %d expects %d entries from rank %d, send list has %d", q,
                             all[q].recv_cnt[me], me, D->send_cnt[q]);
                D->sync.dst_mask |= 1u << q;
                // our slice of rank q's landing buffer 0
                D->sync.dst[q] = reinterpret_cast<double *>((HaloWin *)peers[q] + 1) + all[q].recv_off[me];
                D->sync.dst_stride[q] = all[q].stride;
            }
        }
    } else {
        SIGB_CUDA(cudaMalloc((void **)&D->sendbuf, sizeof(double) * std::max(total_send, 1)));
        SIGB_CUDA(cudaMalloc((void **)&D->halo, sizeof(double) * std::max(nhalo, 1)));
        SIGB_CUDA(cudaMemsetAsync(D->halo, 0, sizeof(double) * std::max(nhalo, 1), st));
    }
    SIGB_CUDA(cudaStreamSynchronize(st));

    sigb_matrix_t A = nullptr;
    rc = sigb_matrix_create(g, &A);
    if (rc != SIGB_OK) return rc;
    guard.g = nullptr;
    sigb_graph_release(g);   // the matrix holds its own reference
    guard.D = nullptr;
    A->dist = D;
    A->nrow = nloc;
    A->ncol = nloc;   // owned columns; the halo is internal
    *out = A;
    return SIGB_OK;
}

int sigb_dist_get_halo(sigb_matrix_t A, int32_t *nhalo, int32_t *halo)
{
    SIGB_REQUIRE(A && A->dist, SIGB_ERR_ARG, "sigb_dist_get_halo: not a row-sharded operator");
    if (nhalo) *nhalo = A->dist->nhalo;
    if (halo)
        for (int32_t i = 0; i < A->dist->nhalo; i++) halo[i] = A->dist->halo_host[i];
    return SIGB_OK;
}

}  // extern "C"
