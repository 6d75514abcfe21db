// comm.cu -- multi-GPU bootstrap of libpmg.so (one process per GPU).
// Round-1 state: the row-slab partition arithmetic (pmg_partition_rows, solver.cu) is final and tested on
// CPU with gloo; the NCCL halo-exchange path is not wired into the solver yet, so these entry points
// report PMG_ERR_UNSUPPORTED instead of pretending.
#include <cstring>

#include "pmg_internal.h"

extern "C" {

pmg_status pmg_comm_unique_id(unsigned char id[PMG_COMM_ID_BYTES])
{
    if (id) std::memset(id, 0, PMG_COMM_ID_BYTES);
    return PMG_ERR_UNSUPPORTED;
}

pmg_status pmg_comm_init(const unsigned char id[PMG_COMM_ID_BYTES], int rank, int n_ranks, int device)
{
    (void)id; (void)rank; (void)n_ranks; (void)device;
    return PMG_ERR_UNSUPPORTED;
}

pmg_status pmg_comm_finalize(void) { return PMG_OK; }

}  // extern "C"
