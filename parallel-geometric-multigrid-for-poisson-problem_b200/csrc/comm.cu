// comm.cu -- multi-GPU plumbing of libpmg.so: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference has no distributed path at all (no MPI/NCCL/threads/streams); this is
// new work for BASELINE configs 3-5.  Levels are partitioned into contiguous ROW SLABS (pmg_partition_rows);
// the only data-path communication is
//   * one halo exchange of 8 rows per level visit on the way down (x on the finest level, the restricted
//     right-hand side on the coarser ones) -- nothing on the way up, the fused passes recompute the halo
//     rows they need;
//   * gather / scatter of the first agglomerated level to / from rank 0;
//   * an all-gather of one double per rank per cycle for the residual norm (summed in rank order on every
//     rank, so all ranks take the same convergence decision).
// NCCL is bound at run time with dlopen("libnccl.so.2") so that single-GPU users need no NCCL at all and a
// process that already loaded an NCCL (e.g. torch's bundled copy) shares that instance.
#include <dlfcn.h>

#include <cstring>
#include <string>
#include <vector>

#include "pmg_internal.h"

namespace pmg {

extern thread_local std::string g_last_error;

namespace {

// the handful of NCCL entry points used, with the ABI of nccl.h 2.x
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclFloat64 = 8 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *(*GetLastError)(ncclComm_t) = nullptr;
};

NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
double *g_agree = nullptr;  // 1 + n_ranks doubles, allocated with the communicator: comm_all_agree must never need memory
int g_rank = 0, g_nranks = 1, g_device = 0;

bool load_nccl(std::string &err)
{
    if (g_nccl.handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return false;
    }
#define PMG_SYM(field, name)                                           \
    *(void **)(&g_nccl.field) = dlsym(h, name);                       \
    if (!g_nccl.field) {                                              \
        err = std::string("NCCL symbol missing: ") + name;            \
        return false;                                                  \
    }
    PMG_SYM(GetUniqueId, "ncclGetUniqueId")
    PMG_SYM(CommInitRank, "ncclCommInitRank")
    PMG_SYM(CommDestroy, "ncclCommDestroy")
    PMG_SYM(GroupStart, "ncclGroupStart")
    PMG_SYM(GroupEnd, "ncclGroupEnd")
    PMG_SYM(Send, "ncclSend")
    PMG_SYM(Recv, "ncclRecv")
    PMG_SYM(AllGather, "ncclAllGather")
    PMG_SYM(GetErrorString, "ncclGetErrorString")
#undef PMG_SYM
    g_nccl.handle = h;
    return true;
}

pmg_status nccl_fail(const char *what, ncclResult_t r)
{
    g_last_error = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return PMG_ERR_COMM;
}

#define PMG_NCCL(call)                                  \
    do {                                                \
        ncclResult_t r_ = (call);                       \
        if (r_ != 0) return nccl_fail(#call, r_);       \
    } while (0)
// between GroupStart and GroupEnd: close the group before reporting, so the communicator is not left inside one
#define PMG_NCCL_G(call)                                \
    do {                                                \
        ncclResult_t r_ = (call);                       \
        if (r_ != 0) {                                  \
            g_nccl.GroupEnd();                          \
            return nccl_fail(#call, r_);                \
        }                                               \
    } while (0)

}  // namespace

bool comm_ready() { return g_comm != nullptr; }
int comm_rank() { return g_rank; }
int comm_size() { return g_nranks; }

// Halo exchange of `depth` rows of a slab array with `ny` owned rows (logical origin `p`, row pitch
// `pitch`): my first/last owned rows go to the neighbours' halo rows and theirs come into mine.
// `ny_up`: owned rows of rank-1 is not needed -- each side addresses only its own array.
pmg_status comm_halo_exchange(double *p, int ny, int pitch, int depth, cudaStream_t st)
{
    if (!g_comm) return PMG_OK;
    const size_t cnt = (size_t)depth * pitch;
    double *row0 = p - PADX;  // whole padded rows: keeps every transfer one contiguous block
    PMG_NCCL(g_nccl.GroupStart());
    if (g_rank > 0) {  // upper neighbour: send my rows [0, depth), receive my halo rows [-depth, 0)
        PMG_NCCL_G(g_nccl.Send(row0, cnt, ncclFloat64, g_rank - 1, g_comm, st));
        PMG_NCCL_G(g_nccl.Recv(row0 - (ptrdiff_t)depth * pitch, cnt, ncclFloat64, g_rank - 1, g_comm, st));
    }
    if (g_rank < g_nranks - 1) {  // lower neighbour: send rows [ny-depth, ny), receive halo rows [ny, ny+depth)
        PMG_NCCL_G(g_nccl.Send(row0 + (ptrdiff_t)(ny - depth) * pitch, cnt, ncclFloat64, g_rank + 1, g_comm, st));
        PMG_NCCL_G(g_nccl.Recv(row0 + (ptrdiff_t)ny * pitch, cnt, ncclFloat64, g_rank + 1, g_comm, st));
    }
    PMG_NCCL(g_nccl.GroupEnd());
    return PMG_OK;
}

// Gather the owned rows of every rank's slab into rank 0's whole-level array (same pitch).
// rows_of(r, &y0, &y1): slab of rank r on this level.
pmg_status comm_gather_rows(const double *slab, double *full, int pitch, const int *y0s, const int *y1s,
                            cudaStream_t st)
{
    if (!g_comm) return PMG_OK;
    PMG_NCCL(g_nccl.GroupStart());
    if (g_rank == 0) {
        for (int r = 1; r < g_nranks; ++r)
            PMG_NCCL_G(g_nccl.Recv(full - PADX + (ptrdiff_t)y0s[r] * pitch, (size_t)(y1s[r] - y0s[r]) * pitch,
                                 ncclFloat64, r, g_comm, st));
    } else {
        PMG_NCCL_G(g_nccl.Send(slab - PADX, (size_t)(y1s[g_rank] - y0s[g_rank]) * pitch, ncclFloat64, 0, g_comm, st));
    }
    PMG_NCCL(g_nccl.GroupEnd());
    if (g_rank == 0) {  // rank 0's own slab: device-to-device copy of its rows
        cudaError_t e = cudaMemcpyAsync(full - PADX + (ptrdiff_t)y0s[0] * pitch, slab - PADX,
                                        (size_t)(y1s[0] - y0s[0]) * pitch * sizeof(double),
                                        cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) {
            g_last_error = std::string("gather copy: ") + cudaGetErrorString(e);
            return PMG_ERR_CUDA;
        }
    }
    return PMG_OK;
}

// Scatter rows [y0 - halo, y1 + halo) (clipped to the level) of rank 0's whole-level array into every
// rank's slab, halo rows included.
pmg_status comm_scatter_rows(const double *full, double *slab, int n_rows, int pitch, const int *y0s,
                             const int *y1s, int halo, cudaStream_t st)
{
    if (!g_comm) return PMG_OK;
    auto range = [&](int r, int &a, int &b) {
        a = y0s[r] - halo < 0 ? 0 : y0s[r] - halo;
        b = y1s[r] + halo > n_rows ? n_rows : y1s[r] + halo;
    };
    PMG_NCCL(g_nccl.GroupStart());
    if (g_rank == 0) {
        for (int r = 1; r < g_nranks; ++r) {
            int a, b;
            range(r, a, b);
            PMG_NCCL_G(g_nccl.Send(full - PADX + (ptrdiff_t)a * pitch, (size_t)(b - a) * pitch, ncclFloat64, r, g_comm, st));
        }
    } else {
        int a, b;
        range(g_rank, a, b);
        PMG_NCCL_G(g_nccl.Recv(slab - PADX + (ptrdiff_t)(a - y0s[g_rank]) * pitch, (size_t)(b - a) * pitch,
                             ncclFloat64, 0, g_comm, st));
    }
    PMG_NCCL(g_nccl.GroupEnd());
    if (g_rank == 0) {
        int a, b;
        range(0, a, b);
        cudaError_t e = cudaMemcpyAsync(slab - PADX + (ptrdiff_t)(a - y0s[0]) * pitch, full - PADX + (ptrdiff_t)a * pitch,
                                        (size_t)(b - a) * pitch * sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) {
            g_last_error = std::string("scatter copy: ") + cudaGetErrorString(e);
            return PMG_ERR_CUDA;
        }
    }
    return PMG_OK;
}

pmg_status comm_allgather_rows(const double *slab, double *full, int rows, int pitch, cudaStream_t st)
{
    if (!g_comm) return PMG_OK;
    PMG_NCCL(g_nccl.AllGather(slab - PADX, full - PADX, (size_t)rows * pitch, ncclFloat64, g_comm, st));
    return PMG_OK;
}

// ---- NVLink peer access (CUDA IPC) ----------------------------------------------------------------------
// Phase 1 (collective): every rank exports one cudaMalloc'ed allocation; `handles` (n_ranks entries of
// IPC_HANDLE_BYTES) receives everybody's handle.  Nothing is opened here, so a rank cannot drop out half way.
pmg_status comm_ipc_exchange(void *base, unsigned char *handles, cudaStream_t st)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == IPC_HANDLE_BYTES, "IPC handle size");
    if (!g_comm) return PMG_OK;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    cudaError_t e = cudaIpcGetMemHandle(&mine, base);
    bool have = (e == cudaSuccess);
    if (!have) cudaGetLastError();
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    // one allocation for both buffers; if even that fails this process cannot take part in the collective at all
    // (the other ranks then see the NCCL error / abort of this one rather than a silent hang)
    unsigned char *d_buf = nullptr;
    if (cudaMalloc((void **)&d_buf, hb * (g_nranks + 1)) != cudaSuccess) {
        cudaGetLastError();
        g_last_error = "cudaMalloc failed (ipc handle exchange)";
        return PMG_ERR_ALLOC;
    }
    unsigned char *d_mine = d_buf, *d_all = d_buf + hb;
    cudaMemcpyAsync(d_mine, &mine, hb, cudaMemcpyHostToDevice, st);
    ncclResult_t r = g_nccl.AllGather(d_mine, d_all, hb, ncclInt8, g_comm, st);
    cudaMemcpyAsync(handles, d_all, hb * g_nranks, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    cudaFree(d_buf);
    if (r != 0) return nccl_fail("ncclAllGather(ipc handles)", r);
    if (e != cudaSuccess) {
        g_last_error = std::string("ipc handle exchange: ") + cudaGetErrorString(e);
        return PMG_ERR_CUDA;
    }
    if (!have) {
        g_last_error = "cudaIpcGetMemHandle failed";
        return PMG_ERR_COMM;  // reported AFTER the collective so that the other ranks are not left waiting
    }
    return PMG_OK;
}

// Phase 2 (local): map one exchanged handle; nullptr on failure.
void *comm_ipc_open(const unsigned char *handle)
{
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

// Phase 3 (collective): true iff every rank says ok.
bool comm_all_agree(bool ok, double *d_scratch /* 1 + n_ranks doubles */, cudaStream_t st)
{
    if (!g_comm) return ok;
    if (d_scratch == nullptr) d_scratch = g_agree;
    if (d_scratch == nullptr) return false;
    double mine = ok ? 1.0 : 0.0;
    std::vector<double> all((size_t)g_nranks, 0.0);
    cudaMemcpyAsync(d_scratch, &mine, sizeof(double), cudaMemcpyHostToDevice, st);
    ncclResult_t r = g_nccl.AllGather(d_scratch, d_scratch + 1, 1, ncclFloat64, g_comm, st);
    cudaMemcpyAsync(all.data(), d_scratch + 1, sizeof(double) * g_nranks, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || r != 0) return false;
    for (double v : all)
        if (v != 1.0) return false;
    return true;
}

void comm_ipc_close(void *peer)
{
    if (peer) cudaIpcCloseMemHandle(peer);
}

// every rank receives every rank's double, in rank order
pmg_status comm_allgather_double(const double *d_mine, double *d_all, cudaStream_t st)
{
    if (!g_comm) return PMG_OK;
    PMG_NCCL(g_nccl.AllGather(d_mine, d_all, 1, ncclFloat64, g_comm, st));
    return PMG_OK;
}

}  // namespace pmg

using namespace pmg;

extern "C" {

pmg_status pmg_comm_unique_id(unsigned char id[PMG_COMM_ID_BYTES])
{
    if (!id) {
        g_last_error = "null argument";
        return PMG_ERR_INVALID;
    }
    std::string err;
    if (!load_nccl(err)) {
        g_last_error = err;
        return PMG_ERR_COMM;
    }
    ncclUniqueId uid;
    std::memset(&uid, 0, sizeof(uid));
    PMG_NCCL(g_nccl.GetUniqueId(&uid));
    static_assert(sizeof(uid) == PMG_COMM_ID_BYTES, "unique id size");
    std::memcpy(id, &uid, PMG_COMM_ID_BYTES);
    return PMG_OK;
}

pmg_status pmg_comm_init(const unsigned char id[PMG_COMM_ID_BYTES], int rank, int n_ranks, int device)
{
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
        g_last_error = "bad argument";
        return PMG_ERR_INVALID;
    }
    if (g_comm) {
        g_last_error = "communicator already initialised";
        return PMG_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        g_last_error = "no CUDA device visible; this library has no CPU fallback";
        return PMG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) {
        g_last_error = "device ordinal out of range";
        return PMG_ERR_INVALID;
    }
    std::string err;
    if (!load_nccl(err)) {
        g_last_error = err;
        return PMG_ERR_COMM;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        g_last_error = "cudaSetDevice failed";
        return PMG_ERR_CUDA;
    }
    ncclUniqueId uid;
    std::memcpy(&uid, id, PMG_COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    PMG_NCCL(g_nccl.CommInitRank(&c, n_ranks, uid, rank));
    g_comm = c;
    g_rank = rank;
    g_nranks = n_ranks;
    g_device = device;
    if (cudaMalloc((void **)&g_agree, sizeof(double) * (size_t)(n_ranks + 1)) != cudaSuccess) {
        cudaGetLastError();
        g_agree = nullptr;
    }
    return PMG_OK;
}

pmg_status pmg_comm_finalize(void)
{
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    if (g_agree) {
        cudaFree(g_agree);
        g_agree = nullptr;
    }
    g_rank = 0;
    g_nranks = 1;
    return PMG_OK;
}

}  // extern "C"
