// comm.cu -- multi-GPU communicator and row-sharded operators (placeholder:
// the single-GPU path is complete; the sharded path lands next).
#include "dist.h"
#include "solvers.h"

namespace sigb {

struct DistInfo { int dummy; };

int dist_matvec(sigb_matrix_t, const double *, double *, bool, const DotSpec &, bool)
{
    set_error("row-sharded operators are not built yet");
    return SIGB_ERR_UNSUPPORTED;
}
int dist_destroy(sigb_matrix_t) { return SIGB_OK; }
int64_t dist_global_n(sigb_matrix_t A) { return A->nrow; }
int64_t dist_row_offset(sigb_matrix_t) { return 0; }
int64_t dist_halo_len(sigb_matrix_t) { return 0; }
int dist_allreduce(sigb_matrix_t, double *, int) { return SIGB_OK; }
int dist_allreduce2(sigb_matrix_t, double *, double *) { return SIGB_OK; }

}  // namespace sigb

using namespace sigb;
extern "C" {
int sigb_comm_unique_id(void *) { set_error("not built yet"); return SIGB_ERR_UNSUPPORTED; }
int sigb_comm_create(const void *, int, int, sigb_comm_t *) { set_error("not built yet"); return SIGB_ERR_UNSUPPORTED; }
int sigb_comm_destroy(sigb_comm_t) { return SIGB_OK; }
int sigb_comm_info(sigb_comm_t, int *, int *, int *) { set_error("not built yet"); return SIGB_ERR_UNSUPPORTED; }
int sigb_dist_csr_create(sigb_comm_t, int32_t, const int32_t *, const int32_t *, const int32_t *, sigb_matrix_t *) { set_error("not built yet"); return SIGB_ERR_UNSUPPORTED; }
int sigb_dist_get_halo(sigb_matrix_t, int32_t *, int32_t *) { set_error("not built yet"); return SIGB_ERR_UNSUPPORTED; }
}
