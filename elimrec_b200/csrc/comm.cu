// comm - NCCL collectives of the multi-GPU modes behind the C-ABI (the reference has no
// inter-GPU code at all, so there is no reference interface to mirror: these replace what a torch.distributed call would do).
//
// One communicator per process (one process per GPU), created from a unique id that the caller distributes however it
// likes (rank 0 calls elimrec_comm_unique_id, torch.distributed / MPI / a file broadcasts the 128 bytes).  Every collective
// is enqueued on the CALLER's stream, in order with the kernels around it: no host synchronisation, no extra streams - and,
// unlike collectives issued through torch.distributed's process group, capturable into the training step's CUDA graph.
// NCCL is resolved at run time with dlsym from the libnccl.so.2 the process already has loaded (torch's), so the library
// has no link-time NCCL dependency and never brings a second NCCL into the process.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

NcclApi* nccl() {
    static NcclApi api = [] {
        NcclApi a{};
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy torch loaded, if any
        if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h == nullptr) return a;
#define NCCL_SYM(name) *(void**)(&a.name) = dlsym(h, "nccl" #name)
        NCCL_SYM(GetUniqueId); NCCL_SYM(CommInitRank); NCCL_SYM(CommDestroy); NCCL_SYM(AllReduce); NCCL_SYM(AllGather);
        NCCL_SYM(Send); NCCL_SYM(Recv); NCCL_SYM(GroupStart); NCCL_SYM(GroupEnd); NCCL_SYM(GetErrorString);
#undef NCCL_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.Send && a.Recv &&
               a.GroupStart && a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return &api;
}

struct Comm {
    ncclComm_t comm;
    int rank, world;
};

#define ER_NCCL(call)                                                                         \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            elimrec_set_error("%s: %s", __func__, nccl()->GetErrorString(r__));               \
            return -5;                                                                        \
        }                                                                                     \
    } while (0)

#define ER_NEED_NCCL()                                                                        \
    do {                                                                                      \
        if (!nccl()->ok) {                                                                    \
            elimrec_set_error("%s: libnccl.so.2 not found in this process", __func__);        \
            return -6;                                                                        \
        }                                                                                     \
    } while (0)

}  // namespace

ELIMREC_API int elimrec_comm_unique_id(void* unique_id_out_host /* ELIMREC_COMM_ID_BYTES */) {
    ER_CHECK_ARG(unique_id_out_host != nullptr, "NULL buffer");
    ER_NEED_NCCL();
    static_assert(sizeof(ncclUniqueId) == ELIMREC_COMM_ID_BYTES, "unique id size");
    ER_NCCL(nccl()->GetUniqueId(reinterpret_cast<ncclUniqueId*>(unique_id_out_host)));
    return 0;
}

ELIMREC_API int elimrec_comm_init(const void* unique_id_host, int rank, int world, void** comm_out) {
    ER_CHECK_ARG(unique_id_host != nullptr && comm_out != nullptr && world >= 1 && rank >= 0 && rank < world, "bad arguments");
    ER_NEED_NCCL();
    ncclUniqueId id;
    memcpy(&id, unique_id_host, sizeof(id));
    Comm* c = new Comm{nullptr, rank, world};
    ncclResult_t r = nccl()->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        elimrec_set_error("elimrec_comm_init: %s", nccl()->GetErrorString(r));
        delete c;
        return -5;
    }
    *comm_out = c;
    return 0;
}

ELIMREC_API int elimrec_comm_destroy(void* comm) {
    if (comm == nullptr) return 0;
    Comm* c = static_cast<Comm*>(comm);
    if (nccl()->ok && c->comm != nullptr) nccl()->CommDestroy(c->comm);
    delete c;
    return 0;
}

ELIMREC_API int elimrec_comm_allreduce(void* comm, const float* send, float* recv, int64_t n, int average,
                                       elimrec_stream_t stream) {
    ER_CHECK_ARG(comm != nullptr && n >= 0, "bad arguments");
    if (n == 0) return 0;
    Comm* c = static_cast<Comm*>(comm);
    ER_NCCL(nccl()->AllReduce(send, recv, (size_t)n, ncclFloat32, average ? ncclAvg : ncclSum, c->comm, er_stream(stream)));
    return 0;
}

ELIMREC_API int elimrec_comm_allgather(void* comm, const void* send, void* recv, int64_t bytes_per_rank, elimrec_stream_t stream) {
    ER_CHECK_ARG(comm != nullptr && bytes_per_rank >= 0, "bad arguments");
    if (bytes_per_rank == 0) return 0;
    Comm* c = static_cast<Comm*>(comm);
    ER_NCCL(nccl()->AllGather(send, recv, (size_t)bytes_per_rank, ncclInt8, c->comm, er_stream(stream)));
    return 0;
}

// recv[q * bytes : (q+1) * bytes] = rank q's send[me * bytes : (me+1) * bytes]   (one grouped batch of send / recv pairs)
ELIMREC_API int elimrec_comm_alltoall(void* comm, const void* send, void* recv, int64_t bytes_per_pair, elimrec_stream_t stream) {
    ER_CHECK_ARG(comm != nullptr && bytes_per_pair >= 0, "bad arguments");
    if (bytes_per_pair == 0) return 0;
    Comm* c = static_cast<Comm*>(comm);
    cudaStream_t st = er_stream(stream);
    ER_NCCL(nccl()->GroupStart());
    for (int q = 0; q < c->world; ++q) {
        ncclResult_t r1 = nccl()->Send(static_cast<const char*>(send) + (size_t)q * bytes_per_pair, (size_t)bytes_per_pair, ncclInt8, q,
                                       c->comm, st);
        ncclResult_t r2 = nccl()->Recv(static_cast<char*>(recv) + (size_t)q * bytes_per_pair, (size_t)bytes_per_pair, ncclInt8, q,
                                       c->comm, st);
        if (r1 != ncclSuccess || r2 != ncclSuccess) {
            nccl()->GroupEnd();
            elimrec_set_error("elimrec_comm_alltoall: %s", nccl()->GetErrorString(r1 != ncclSuccess ? r1 : r2));
            return -5;
        }
    }
    ER_NCCL(nccl()->GroupEnd());
    return 0;
}
