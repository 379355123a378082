// comm.cu — NCCL plumbing for the halo exchange and the Gram/norm reductions.
// Replaces the MPI calls of src/utils/MPICommunicatorP2P.t.cpp:89-222,288-420 (MPI_Isend/Irecv/Waitall) with
// one grouped ncclSend/ncclRecv per neighbour on the plan's stream, and the MPI_Allreduce of
// src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:803-809 / MultiVector.t.cpp:567-573 with ncclAllReduce in
// place on the device (no D2H staging).  NCCL is resolved at run time with dlopen so that single-GPU use has
// no NCCL dependency and, inside a PyTorch process, the already-loaded libnccl.so.2 is shared.
#include <dlfcn.h>

#include "hx_internal.h"

namespace hx
{
  typedef struct ncclComm *ncclComm_t;
  typedef struct
  {
    char internal[128];
  } ncclUniqueId;
  enum
  {
    ncclSuccess = 0
  };
  enum
  {
    ncclInt8   = 0,
    ncclFloat64 = 8
  };
  enum
  {
    ncclSum = 0
  };

  struct Nccl
  {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *)                                                      = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                               = nullptr;
    int (*CommDestroy)(ncclComm_t)                                                          = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t)                   = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t)                         = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t)      = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t)           = nullptr;
    int (*GroupStart)()                                                                     = nullptr;
    int (*GroupEnd)()                                                                       = nullptr;
    const char *(*GetErrorString)(int)                                                      = nullptr;
  };
  static Nccl g_nccl;

  static int
  load_nccl()
  {
    if (g_nccl.handle)
      return HX_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *      h       = nullptr;
    for (const char *n : names)
      {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h)
          break;
      }
    HX_CHECK(h, HX_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define HX_SYM(field, name)                                          \
  *(void **)(&g_nccl.field) = dlsym(h, name);                        \
  HX_CHECK(g_nccl.field, HX_ERR_COMM, "NCCL symbol %s missing", name)
    HX_SYM(GetUniqueId, "ncclGetUniqueId");
    HX_SYM(CommInitRank, "ncclCommInitRank");
    HX_SYM(CommDestroy, "ncclCommDestroy");
    HX_SYM(Send, "ncclSend");
    HX_SYM(Recv, "ncclRecv");
    HX_SYM(AllReduce, "ncclAllReduce");
    HX_SYM(AllGather, "ncclAllGather");
    HX_SYM(GroupStart, "ncclGroupStart");
    HX_SYM(GroupEnd, "ncclGroupEnd");
    HX_SYM(GetErrorString, "ncclGetErrorString");
#undef HX_SYM
    g_nccl.handle = h;
    return HX_OK;
  }

#define HX_NCCL(call)                                                                              \
  do                                                                                               \
    {                                                                                              \
      int r_ = (call);                                                                             \
      if (r_ != ncclSuccess)                                                                       \
        {                                                                                          \
          set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));      \
          return HX_ERR_COMM;                                                                      \
        }                                                                                          \
    }                                                                                              \
  while (0)

  struct Comm
  {
    ncclComm_t comm = nullptr;
    int        nranks = 1, rank = 0;
  };

  int
  comm_unique_id(char id[128])
  {
    HX_TRY(load_nccl());
    ncclUniqueId u;
    HX_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return HX_OK;
  }

  int
  comm_create(Comm **c, const char id[128], int nranks, int rank)
  {
    HX_TRY(load_nccl());
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    Comm *cc   = new Comm();
    cc->nranks = nranks;
    cc->rank   = rank;
    int r      = g_nccl.CommInitRank(&cc->comm, nranks, u, rank);
    if (r != ncclSuccess)
      {
        set_error("ncclCommInitRank(%d/%d) failed: %s", rank, nranks, g_nccl.GetErrorString(r));
        delete cc;
        return HX_ERR_COMM;
      }
    *c = cc;
    return HX_OK;
  }

  void
  comm_destroy(Comm *c)
  {
    if (c)
      {
        if (c->comm && g_nccl.CommDestroy)
          g_nccl.CommDestroy(c->comm);
        delete c;
      }
  }

  // one grouped exchange: segment i of `send` goes to send_procs[i], segment i of `recv` comes from recv_procs[i]
  int
  comm_exchange(Comm *c, cudaStream_t s, const double *send, const std::vector<uint32_t> &send_procs,
                const std::vector<size_t> &send_counts, double *recv, const std::vector<uint32_t> &recv_procs,
                const std::vector<size_t> &recv_counts)
  {
    HX_CHECK(c && c->comm, HX_ERR_COMM, "communicator not initialised");
    HX_NCCL(g_nccl.GroupStart());
    size_t off = 0;
    for (size_t i = 0; i < recv_procs.size(); ++i)
      {
        if (recv_counts[i])
          HX_NCCL(g_nccl.Recv(recv + off, recv_counts[i], ncclFloat64, (int)recv_procs[i], c->comm, s));
        off += recv_counts[i];
      }
    off = 0;
    for (size_t i = 0; i < send_procs.size(); ++i)
      {
        if (send_counts[i])
          HX_NCCL(g_nccl.Send(send + off, send_counts[i], ncclFloat64, (int)send_procs[i], c->comm, s));
        off += send_counts[i];
      }
    HX_NCCL(g_nccl.GroupEnd());
    return HX_OK;
  }

  // small host-to-host all-gather (setup metadata of the peer-memory halo): staged through device memory
  int
  comm_allgather_bytes(Comm *c, cudaStream_t s, const void *mine, void *all, size_t bytes_per_rank)
  {
    HX_CHECK(c && c->comm, HX_ERR_COMM, "communicator not initialised");
    DevBuf<unsigned char> ds, dr;
    HX_TRY(ds.alloc(bytes_per_rank));
    HX_TRY(dr.alloc(bytes_per_rank * (size_t)c->nranks));
    HX_CUDA(cudaMemcpyAsync(ds.p, mine, bytes_per_rank, cudaMemcpyHostToDevice, s));
    HX_NCCL(g_nccl.AllGather(ds.p, dr.p, bytes_per_rank, ncclInt8, c->comm, s));
    HX_CUDA(cudaMemcpyAsync(all, dr.p, bytes_per_rank * (size_t)c->nranks, cudaMemcpyDeviceToHost, s));
    HX_CUDA(cudaStreamSynchronize(s));
    return HX_OK;
  }

  int
  comm_allreduce_sum(Comm *c, cudaStream_t s, double *buf, size_t n)
  {
    HX_CHECK(c && c->comm, HX_ERR_COMM, "communicator not initialised");
    HX_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat64, ncclSum, c->comm, s));
    return HX_OK;
  }
} // namespace hx
