// dist.h -- row-sharded (multi-GPU) operators: entry points used by api.cu
// and solvers.cu.  Implemented in comm.cu.
#pragma once

#include "internal.h"

namespace sigb {

int dist_matvec(sigb_matrix_t A, const double *x, double *y, bool add, const DotSpec &dot,
                bool x_has_halo);
int dist_destroy(sigb_matrix_t A);
int64_t dist_global_n(sigb_matrix_t A);
int64_t dist_row_offset(sigb_matrix_t A);


// ranks as threads of one process (single-process multi-device mode, mgpu.cu)
struct LocalGroup;
LocalGroup *local_group_create(int nranks);
void local_group_destroy(LocalGroup *g);
void local_group_barrier(LocalGroup *g);
int comm_create_local(LocalGroup *grp, int rank, int nranks, sigb_comm_t *out);

}  // namespace sigb
