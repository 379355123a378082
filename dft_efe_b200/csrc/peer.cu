// peer.cu — halo exchange over NVLink peer memory, fused with the pack / unpack kernels.
//
// Replaces MPICommunicatorP2P::updateGhostValues / accumulateAddLocallyOwned
// (src/utils/MPICommunicatorP2P.t.cpp:77-273, 278-470): instead of pack -> MPI_Isend/Irecv -> unpack (or the
// NCCL send/recv group of comm.cu, kept as the fallback), the PACK kernel stores the rows straight into the
// neighbour's receive buffer through a peer mapping (cudaIpc) and raises a sequence flag there; the neighbour's
// UNPACK / ordered-add kernel waits for the flags of its sources, consumes the buffer and acknowledges.  One
// exchange = two small kernels and no library call: ~10 us instead of ~30 us per exchange, and a C2-sized apply
// makes four of them (X halo, Y halo, projector accumulate + update).
//
// Flow control: every (halo, direction) carries a sequence number that all ranks advance together (the calls are
// collective, like the reference's).  A sender writes message s only after the receiver acknowledged message
// s-1 (ack word in the SENDER's arena, written remotely by the receiver); a receiver consumes message s when the
// flag word in ITS arena reaches s.  All waits are bounded (clock64): on timeout a status word is raised and the
// kernel exits, so a lost peer turns into HX_ERR_COMM at the next synchronisation instead of a hung GPU.
//
// Memory model: data stores to peer memory -> __threadfence_system() -> (last block) flag store with
// st.release.sys; the consumer polls with ld.acquire.sys and reads the buffer with ld.cg (never a stale L1 line).
#include "hx_internal.h"

namespace hx
{
  struct PeerDir // one direction of one halo, as the kernels see it
  {
    // sender side
    double *const *  rbase;     // [nDst] receive buffer of destination d (peer mapping)
    const uint32_t * rrowoff;   // [nDst] first row of my segment inside it
    uint32_t *const *rflag;     // [nDst] flag word at destination d (peer mapping)
    const uint32_t * ack;       // [nDst] acknowledgements from destination d (my arena)
    const uint32_t * seg;       // [nRows] destination index of send row k
    const uint32_t * segbegin;  // [nDst+1] first send row of destination d
    uint32_t         nDst, nRows;
    // receiver side
    const double *   recv;      // my receive buffer
    const uint32_t * flag;      // [nSrc] flags raised by source s (my arena)
    uint32_t *const *rack;      // [nSrc] ack word at source s (peer mapping)
    uint32_t         nSrc;
    uint32_t *       counter;   // [2] block counters (push, consume)
    uint32_t *       status;    // raised on timeout
  };

  // rows[k] = local row of send row k; message layout at the destination: (rrowoff + k - segbegin) * B + v
  __global__ void __launch_bounds__(256)
  peer_push_kernel(PeerDir d, const double *x, const uint32_t *rows, uint32_t B, uint32_t seq)
  {
    __shared__ int ok;
    if (threadIdx.x == 0)
      {
        ok = wait_words(d.ack, d.nDst, seq - 1u, d.status) ? 1 : 0;
        if (!ok)
          atomicExch(d.status, 1u);
      }
    __syncthreads();
    if (ok)
      {
        const bool   vec = (B % 2 == 0);
        const size_t per = vec ? B / 2 : B;
        const size_t tot = (size_t)d.nRows * per;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
          {
            const uint32_t k = (uint32_t)(i / per), c = (uint32_t)(i % per);
            const uint32_t s = d.seg[k];
            double *       dst = d.rbase[s] + ((size_t)d.rrowoff[s] + (k - d.segbegin[s])) * B;
            const double * src = x + (size_t)rows[k] * B;
            if (vec)
              reinterpret_cast<double2 *>(dst)[c] = reinterpret_cast<const double2 *>(src)[c];
            else
              dst[c] = src[c];
          }
      }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      {
        const uint32_t done = atomicAdd(d.counter, 1u);
        if (done == gridDim.x - 1)
          {
            d.counter[0] = 0u;
            __threadfence_system();
            // a timed-out exchange raises no flag: the neighbours run into their own timeout instead of consuming stale rows
            if (*reinterpret_cast<const volatile uint32_t *>(d.status) == 0u)
              for (uint32_t s = 0; s < d.nDst; ++s)
                st_release_sys(d.rflag[s], seq);
          }
      }
  }

  // consumer prologue / epilogue shared by the two receive kernels
  __device__ __forceinline__ bool
  peer_consume_begin(const PeerDir &d, uint32_t seq)
  {
    __shared__ int ok;
    if (threadIdx.x == 0)
      {
        ok = wait_words(d.flag, d.nSrc, seq, d.status) ? 1 : 0;
        if (!ok)
          atomicExch(d.status, 1u);
      }
    __syncthreads();
    return ok != 0;
  }
  __device__ __forceinline__ void
  peer_consume_end(const PeerDir &d, uint32_t seq)
  {
    __syncthreads();
    if (threadIdx.x == 0)
      {
        __threadfence();
        const uint32_t done = atomicAdd(d.counter + 1, 1u);
        if (done == gridDim.x - 1)
          {
            d.counter[1] = 0u;
            __threadfence_system();
            if (*reinterpret_cast<const volatile uint32_t *>(d.status) == 0u) // no acknowledgement after a timeout
              for (uint32_t s = 0; s < d.nSrc; ++s)
                st_release_sys(d.rack[s], seq);
          }
      }
  }

  // updateGhostValues, receive side: x[ids[k]] = recv[k]   (DiscontiguousDataOperations.cpp:54-71)
  __global__ void __launch_bounds__(256)
  peer_unpack_kernel(PeerDir d, double *x, const uint32_t *ids, uint32_t n, uint32_t B, uint32_t seq)
  {
    if (peer_consume_begin(d, seq))
      {
        if (B % 2 == 0)
          {
            const uint32_t bv  = B / 2;
            const size_t   tot = (size_t)n * bv;
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
              reinterpret_cast<double2 *>(x + (size_t)ids[i / bv] * B)[i % bv] = __ldcg(reinterpret_cast<const double2 *>(d.recv) + i);
          }
        else
          {
            const size_t tot = (size_t)n * B;
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
              x[(size_t)ids[i / B] * B + (i % B)] = __ldcg(d.recv + i);
          }
      }
    peer_consume_end(d, seq);
  }

  // accumulateAddLocallyOwned, receive side: owner rows += buffer rows in buffer order (no atomics)
  __global__ void __launch_bounds__(256)
  peer_add_rows_kernel(PeerDir d, double *x, const uint32_t *rows, const uint32_t *off, const uint32_t *pos,
                       uint32_t nrows, uint32_t B, uint32_t seq)
  {
    if (peer_consume_begin(d, seq))
      {
        if (B % 2 == 0)
          {
            const uint32_t bv  = B / 2;
            const size_t   tot = (size_t)nrows * bv;
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
              {
                const uint32_t r = (uint32_t)(i / bv), v = (uint32_t)(i % bv);
                double2 *      y = reinterpret_cast<double2 *>(x + (size_t)rows[r] * B) + v;
                double2        s = *y;
                for (uint32_t e = off[r]; e < off[r + 1]; ++e)
                  {
                    const double2 t = __ldcg(reinterpret_cast<const double2 *>(d.recv + (size_t)pos[e] * B) + v);
                    s.x += t.x;
                    s.y += t.y;
                  }
                *y = s;
              }
          }
        else
          {
            const size_t tot = (size_t)nrows * B;
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
              {
                const uint32_t r = (uint32_t)(i / B), v = (uint32_t)(i % B);
                double *       y = x + (size_t)rows[r] * B + v;
                double         s = *y;
                for (uint32_t e = off[r]; e < off[r + 1]; ++e)
                  s += __ldcg(d.recv + (size_t)pos[e] * B + v);
                *y = s;
              }
          }
      }
    peer_consume_end(d, seq);
  }

  // ------------------------------------------------------------------------------------------------
  // host side
  // ------------------------------------------------------------------------------------------------
  struct PeerState
  {
    unsigned char *arena = nullptr;
    size_t         arena_bytes = 0;
    size_t         offU = 0, offA = 0, offW = 0; // recvU, recvA, words
    uint32_t       nGP = 0, nTP = 0;
    std::vector<void *> mapped; // per rank: opened arena (nullptr = not a neighbour)
    // words (u32) in my arena: flagU[nGP] | ackA[nGP] | flagA[nTP] | ackU[nTP] | counters[4] | status
    uint32_t *wFlagU = nullptr, *wAckA = nullptr, *wFlagA = nullptr, *wAckU = nullptr, *wCounter = nullptr, *wStatus = nullptr;
    DevBuf<double *>   d_rbaseU, d_rbaseA;
    DevBuf<uint32_t>   d_rrowoffU, d_rrowoffA, d_segU, d_segA, d_segbeginU, d_segbeginA;
    DevBuf<uint32_t *> d_rflagU, d_rflagA, d_rackU, d_rackA;
    uint32_t           seqU = 0, seqA = 0;
    PeerDir            dirU, dirA;
    DevBuf<double *>   d_push_base; // [n_ghost] owner's accumulate buffer for ghost row j (peer mapping)
    DevBuf<uint32_t>   d_push_row;  // [n_ghost] row of ghost j inside it
    ~PeerState()
    {
      for (void *m : mapped)
        if (m)
          cudaIpcCloseMemHandle(m);
      if (arena)
        cudaFree(arena);
    }
  };

  void
  peer_destroy(PeerState *s)
  {
    delete s;
  }

  struct PeerWire // what every rank publishes about one halo (all-gathered)
  {
    cudaIpcMemHandle_t handle; // 64 bytes
    unsigned long long offU, offA, offW;
    uint32_t           nGP, nTP, pad0, pad1;
    // followed by 4 * nranks u32: segU_rowoff[q], idxU[q], segA_rowoff[q], idxA[q]  (0xffffffff = not a neighbour)
  };

  static size_t
  align_up(size_t x, size_t a)
  {
    return (x + a - 1) / a * a;
  }

  // collective: every rank of the communicator calls this for the same halo at the same point
  int
  peer_setup(hx_plan *p, Halo &h)
  {
    if (h.peer || h.peer_failed)
      return HX_OK;
    const int nr = p->nranks, me = p->rank;
    PeerState *s = new PeerState();
    s->nGP = (uint32_t)h.ghost_procs.size();
    s->nTP = (uint32_t)h.target_procs.size();
    s->offU = 0;
    s->offA = align_up((size_t)h.n_ghost * p->max_block * sizeof(double), 256);
    s->offW = s->offA + align_up((size_t)h.n_send * p->max_block * sizeof(double), 256);
    const size_t nwords = 2 * (size_t)s->nGP + 2 * (size_t)s->nTP + 8;
    s->arena_bytes      = s->offW + align_up(nwords * sizeof(uint32_t), 256);
    int         fail = 0;
    cudaError_t e    = cudaMalloc((void **)&s->arena, s->arena_bytes);
    if (e != cudaSuccess)
      fail = 1;
    const size_t wire_bytes = sizeof(PeerWire) + 4 * sizeof(uint32_t) * (size_t)nr;
    std::vector<unsigned char> mine(wire_bytes, 0), all(wire_bytes * nr, 0);
    PeerWire *w = reinterpret_cast<PeerWire *>(mine.data());
    uint32_t *t = reinterpret_cast<uint32_t *>(mine.data() + sizeof(PeerWire));
    if (!fail)
      {
        cudaMemset(s->arena, 0, s->arena_bytes);
        if (cudaIpcGetMemHandle(&w->handle, s->arena) != cudaSuccess)
          fail = 1;
      }
    w->offU = s->offU, w->offA = s->offA, w->offW = s->offW, w->nGP = s->nGP, w->nTP = s->nTP;
    w->pad0 = (uint32_t)fail;
    for (int q = 0; q < 4 * nr; ++q)
      t[q] = 0xffffffffu;
    for (uint32_t i = 0; i < s->nGP; ++i)
      {
        t[4 * h.ghost_procs[i] + 0] = h.ghost_ranges[2 * i]; // rows of proc i start here in my recvU
        t[4 * h.ghost_procs[i] + 1] = i;
      }
    {
      uint32_t off = 0;
      for (uint32_t i = 0; i < s->nTP; ++i)
        {
          t[4 * h.target_procs[i] + 2] = off; // rows of target i start here in my recvA
          t[4 * h.target_procs[i] + 3] = i;
          off += h.target_counts[i];
        }
    }
    cudaGetLastError();
    HX_TRY(comm_allgather_bytes(p->comm, p->stream, mine.data(), all.data(), wire_bytes));
    auto wire  = [&](int q) { return reinterpret_cast<const PeerWire *>(all.data() + wire_bytes * q); };
    auto table = [&](int q) { return reinterpret_cast<const uint32_t *>(all.data() + wire_bytes * q + sizeof(PeerWire)); };
    for (int q = 0; q < nr; ++q)
      if (wire(q)->pad0)
        fail = 1;
    // open the neighbours' arenas
    s->mapped.assign(nr, nullptr);
    if (!fail)
      {
        std::vector<char> need(nr, 0);
        for (uint32_t q : h.ghost_procs)
          need[q] = 1;
        for (uint32_t q : h.target_procs)
          need[q] = 1;
        for (int q = 0; q < nr && !fail; ++q)
          if (need[q] && q != me)
            {
              void *m = nullptr;
              if (cudaIpcOpenMemHandle(&m, wire(q)->handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
                {
                  fail = 1;
                  cudaGetLastError();
                }
              s->mapped[q] = m;
            }
          else if (need[q])
            s->mapped[q] = s->arena; // a rank listed as its own neighbour (periodic single-rank direction)
      }
    // every rank must take the same decision
    {
      std::vector<unsigned char> f1(4, 0), fall(4 * (size_t)nr, 0);
      f1[0] = (unsigned char)fail;
      HX_TRY(comm_allgather_bytes(p->comm, p->stream, f1.data(), fall.data(), 4));
      for (int q = 0; q < nr; ++q)
        if (fall[4 * (size_t)q])
          fail = 1;
    }
    if (fail)
      {
        delete s;
        h.peer_failed = true;
        return HX_OK; // the NCCL path stays in charge
      }
    uint32_t *words = reinterpret_cast<uint32_t *>(s->arena + s->offW);
    s->wFlagU       = words;
    s->wAckA        = words + s->nGP;
    s->wFlagA       = words + 2 * s->nGP;
    s->wAckU        = words + 2 * s->nGP + s->nTP;
    s->wCounter     = words + 2 * s->nGP + 2 * s->nTP;
    s->wStatus      = s->wCounter + 4;
    auto words_of   = [&](int q) { return reinterpret_cast<uint32_t *>((unsigned char *)s->mapped[q] + wire(q)->offW); };
    // update direction: I send to my target procs; they hold me as a ghost proc
    std::vector<double *>   rbU(s->nTP), rbA(s->nGP);
    std::vector<uint32_t>   roU(s->nTP), roA(s->nGP);
    std::vector<uint32_t *> rfU(s->nTP), rfA(s->nGP), rackU(s->nGP), rackA(s->nTP);
    int                     mismatch = 0;
    for (uint32_t i = 0; i < s->nTP; ++i)
      {
        const int       q  = (int)h.target_procs[i];
        const uint32_t *tq = table(q);
        if (tq[4 * me + 1] == 0xffffffffu)
          {
            mismatch = 1; // halo patterns of the two ranks disagree: voted on below, so that every rank leaves together
            continue;
          }
        rbU[i]   = reinterpret_cast<double *>((unsigned char *)s->mapped[q] + wire(q)->offU);
        roU[i]   = tq[4 * me + 0];
        rfU[i]   = words_of(q) + tq[4 * me + 1];                                   // flagU[idx of me at q]
        rackA[i] = words_of(q) + wire(q)->nGP + tq[4 * me + 1];                    // ackA[idx of me at q]
      }
    // accumulate direction: I send to my ghost procs (the owners); they hold me as a target proc
    for (uint32_t i = 0; i < s->nGP; ++i)
      {
        const int       q  = (int)h.ghost_procs[i];
        const uint32_t *tq = table(q);
        if (tq[4 * me + 3] == 0xffffffffu)
          {
            mismatch = 1;
            continue;
          }
        rbA[i]   = reinterpret_cast<double *>((unsigned char *)s->mapped[q] + wire(q)->offA);
        roA[i]   = tq[4 * me + 2];
        rfA[i]   = words_of(q) + 2 * wire(q)->nGP + tq[4 * me + 3];                 // flagA[idx of me at q]
        rackU[i] = words_of(q) + 2 * wire(q)->nGP + wire(q)->nTP + tq[4 * me + 3]; // ackU[idx of me at q]
      }
    std::vector<uint32_t> segU(h.n_send), segbU(s->nTP + 1, 0), segA(h.n_ghost), segbA(s->nGP + 1, 0);
    for (uint32_t i = 0; i < s->nTP; ++i)
      {
        segbU[i + 1] = segbU[i] + h.target_counts[i];
        for (uint32_t k = segbU[i]; k < segbU[i + 1]; ++k)
          segU[k] = i;
      }
    for (uint32_t i = 0; i < s->nGP; ++i)
      {
        if (h.ghost_ranges[2 * i] != segbA[i])
          mismatch = 1; // ghost ranges must be contiguous per ghost proc
        segbA[i + 1] = h.ghost_ranges[2 * i + 1];
        for (uint32_t k = segbA[i]; k < segbA[i + 1]; ++k)
          segA[k] = i;
      }
    {
      // every rank leaves together when any pair of halo patterns is inconsistent (no rank is left in a collective)
      std::vector<unsigned char> f1(4, 0), fall(4 * (size_t)nr, 0);
      f1[0] = (unsigned char)mismatch;
      int rc = comm_allgather_bytes(p->comm, p->stream, f1.data(), fall.data(), 4);
      for (int q = 0; q < nr; ++q)
        if (fall[4 * (size_t)q])
          mismatch = 1;
      if (rc != HX_OK || mismatch)
        {
          delete s;
          if (rc != HX_OK)
            return rc;
          set_error("halo patterns of the ranks disagree (a rank lists a neighbour that does not list it back, or ghost ranges are not contiguous)");
          return HX_ERR_COMM;
        }
    }
    HX_TRY(s->d_rbaseU.upload(rbU));
    HX_TRY(s->d_rbaseA.upload(rbA));
    HX_TRY(s->d_rrowoffU.upload(roU));
    HX_TRY(s->d_rrowoffA.upload(roA));
    HX_TRY(s->d_rflagU.upload(rfU));
    HX_TRY(s->d_rflagA.upload(rfA));
    HX_TRY(s->d_rackU.upload(rackU));
    HX_TRY(s->d_rackA.upload(rackA));
    HX_TRY(s->d_segU.upload(segU));
    HX_TRY(s->d_segA.upload(segA));
    HX_TRY(s->d_segbeginU.upload(segbU));
    HX_TRY(s->d_segbeginA.upload(segbA));
    HX_CUDA(cudaDeviceSynchronize());
    PeerDir &u = s->dirU;
    u.rbase = s->d_rbaseU.p, u.rrowoff = s->d_rrowoffU.p, u.rflag = s->d_rflagU.p, u.ack = s->wAckU;
    u.seg = s->d_segU.p, u.segbegin = s->d_segbeginU.p, u.nDst = s->nTP, u.nRows = h.n_send;
    u.recv = reinterpret_cast<const double *>(s->arena + s->offU), u.flag = s->wFlagU, u.rack = s->d_rackU.p;
    u.nSrc = s->nGP, u.counter = s->wCounter, u.status = s->wStatus;
    PeerDir &a = s->dirA;
    a.rbase = s->d_rbaseA.p, a.rrowoff = s->d_rrowoffA.p, a.rflag = s->d_rflagA.p, a.ack = s->wAckA;
    a.seg = s->d_segA.p, a.segbegin = s->d_segbeginA.p, a.nDst = s->nGP, a.nRows = h.n_ghost;
    a.recv = reinterpret_cast<const double *>(s->arena + s->offA), a.flag = s->wFlagA, a.rack = s->d_rackA.p;
    a.nSrc = s->nTP, a.counter = s->wCounter + 2, a.status = s->wStatus;
    {
      // ghost row j = ghost_local_ids[k] travels as send row k of the accumulate direction
      std::vector<uint32_t> gids(h.n_ghost);
      if (h.n_ghost)
        HX_CUDA(cudaMemcpy(gids.data(), h.d_ghost_local_ids.p, gids.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      std::vector<double *> pb(h.n_ghost, nullptr);
      std::vector<uint32_t> pr(h.n_ghost, 0);
      for (uint32_t k = 0; k < h.n_ghost; ++k)
        {
          const uint32_t sg = segA[k];
          pb[gids[k]]       = rbA[sg];
          pr[gids[k]]       = roA[sg] + (k - segbA[sg]);
        }
      HX_TRY(s->d_push_base.upload(pb));
      HX_TRY(s->d_push_row.upload(pr));
    }
    h.peer = s;
    p->peer_halos.push_back(&h);
    // nobody may push before every rank has mapped and zeroed its arena
    {
      std::vector<unsigned char> f1(4, 0), fall(4 * (size_t)nr, 0);
      HX_TRY(comm_allgather_bytes(p->comm, p->stream, f1.data(), fall.data(), 4));
    }
    return HX_OK;
  }

  static unsigned
  grid_for(size_t work_items)
  {
    // one thread of every block waits for the neighbours' words (a bounded wait on a remote event: blocks that become
    // resident later simply find it satisfied), the rest is a grid-stride copy: two blocks per SM are plenty
    size_t g = (work_items + 255) / 256;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(g, 296));
  }

  int
  peer_halo_update(hx_plan *p, Halo &h, double *X, uint32_t B)
  {
    PeerState *s   = h.peer;
    const uint32_t seq = ++s->seqU;
    if (h.n_send || s->nTP)
      {
        peer_push_kernel<<<grid_for((size_t)h.n_send * B / 2), 256, 0, p->stream>>>(s->dirU, X, h.d_owned_ids_for_targets.p, B, seq);
        p->launches++;
      }
    if (h.n_ghost || s->nGP)
      {
        peer_unpack_kernel<<<grid_for((size_t)h.n_ghost * B / 2), 256, 0, p->stream>>>(s->dirU, X + (size_t)h.n_owned * B,
                                                                                  h.d_ghost_local_ids.p, h.n_ghost, B, seq);
        p->launches++;
      }
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  int
  peer_halo_accumulate(hx_plan *p, Halo &h, double *Y, uint32_t B)
  {
    PeerState *s   = h.peer;
    const uint32_t seq = ++s->seqA;
    if (h.n_ghost || s->nGP)
      {
        // send row k = ghost row ghost_local_ids[k] (the pack order of the reference)
        peer_push_kernel<<<grid_for((size_t)h.n_ghost * B / 2), 256, 0, p->stream>>>(s->dirA, Y + (size_t)h.n_owned * B,
                                                                                    h.d_ghost_local_ids.p, B, seq);
        p->launches++;
      }
    if (h.n_send || s->nTP)
      {
        peer_add_rows_kernel<<<grid_for((size_t)h.n_acc_rows * B / 2), 256, 0, p->stream>>>(s->dirA, Y, h.d_acc_rows.p, h.d_acc_off.p,
                                                                                       h.d_acc_pos.p, h.n_acc_rows, B, seq);
        p->launches++;
      }
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }


  // ---- halo overlap: the cell kernel unpacks the ghost rows of X and pushes the ghost-row sums of Y itself ----
  // closing kernel of the accumulate: pushes the ghost rows the cell kernel could not (rows that receive contributions
  // after it: constrained / parent / staged / untouched ones), then raises the flags of the message at the owners.  It
  // runs after the cell kernel in stream order, so the cell kernel's own peer stores are complete.
  __global__ void __launch_bounds__(256)
  peer_push_rest_kernel(PeerDir d, const double *yghost, const uint32_t *gids, const uint32_t *rest, uint32_t nrest, uint32_t B,
                        uint32_t seq)
  {
    const size_t tot = (size_t)nrest * B;
    const bool   ok  = *reinterpret_cast<const volatile uint32_t *>(d.status) == 0u;
    if (ok)
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
        {
          const uint32_t k = rest[i / B], c = (uint32_t)(i % B);
          const uint32_t s = d.seg[k];
          d.rbase[s][((size_t)d.rrowoff[s] + (k - d.segbegin[s])) * B + c] = yghost[(size_t)gids[k] * B + c];
        }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      {
        const uint32_t done = atomicAdd(d.counter, 1u);
        if (done == gridDim.x - 1)
          {
            d.counter[0] = 0u;
            __threadfence_system();
            if (ok) // a timed-out exchange raises no flag: the neighbours time out as well instead of reading garbage
              for (uint32_t s = 0; s < d.nDst; ++s)
                st_release_sys(d.rflag[s], seq);
          }
      }
  }

  bool
  peer_overlap_available(const Halo &h)
  {
    return h.peer != nullptr;
  }

  // updateGhostValues, send side only (the receive side runs inside the cell kernel)
  int
  peer_push_update(hx_plan *p, Halo &h, const double *X, uint32_t B, uint32_t *seq)
  {
    PeerState *s = h.peer;
    *seq         = ++s->seqU;
    if (h.n_send || s->nTP)
      {
        peer_push_kernel<<<grid_for((size_t)h.n_send * B / 2), 256, 0, p->stream>>>(s->dirU, X, h.d_owned_ids_for_targets.p, B, *seq);
        p->launches++;
      }
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  int
  peer_overlap_args(hx_plan *p, Halo &h, double *X, uint32_t B, bool do_unpack, uint32_t seqU, HaloK *k)
  {
    PeerState *s  = h.peer;
    k->x_ready    = p->d_x_ready.p;
    k->flagU      = s->wFlagU;
    k->rackU      = s->d_rackU.p;
    k->recvU      = reinterpret_cast<const double *>(s->arena + s->offU);
    k->unpack_ids = p->d_unpack_ids.p;
    k->xghost     = X + (size_t)h.n_owned * B;
    k->nSrcU      = do_unpack ? s->nGP : 0u;
    k->seqU       = seqU;
    k->n_ghost    = h.n_ghost;
    k->do_unpack  = do_unpack ? 1u : 0u;
    k->ackA       = s->wAckA;
    k->push_base  = s->d_push_base.p;
    k->push_row   = s->d_push_row.p;
    k->nDstA      = s->nGP;
    k->seqA       = ++s->seqA;
    k->n_owned    = h.n_owned;
    k->do_push    = 1u;
    k->counter    = s->wCounter + 5; // own word: [0..1] / [2..3] are the push / consume counters of the two directions, [4] status
    k->status     = s->wStatus;
    const size_t bytes = (size_t)h.n_ghost * B * sizeof(double);
    k->n_halo_ctas     = (uint32_t)std::max<size_t>(1, std::min<size_t>(32, bytes / (256 * 1024) + 1));
    return HX_OK;
  }

  // accumulateAddLocallyOwned after a cell kernel that pushed its ghost-row sums itself
  int
  peer_finish_accumulate(hx_plan *p, Halo &h, double *Y, uint32_t B, uint32_t seqA)
  {
    PeerState *s = h.peer;
    if (h.n_ghost || s->nGP)
      {
        peer_push_rest_kernel<<<grid_for((size_t)p->n_push_rest * B), 256, 0, p->stream>>>(
          s->dirA, Y + (size_t)h.n_owned * B, h.d_ghost_local_ids.p, p->d_push_rest.p, p->n_push_rest, B, seqA);
        p->launches++;
      }
    if (h.n_send || s->nTP)
      {
        peer_add_rows_kernel<<<grid_for((size_t)h.n_acc_rows * B / 2), 256, 0, p->stream>>>(s->dirA, Y, h.d_acc_rows.p, h.d_acc_off.p,
                                                                                       h.d_acc_pos.p, h.n_acc_rows, B, seqA);
        p->launches++;
      }
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // ---- accumulateAddLocallyOwned followed by updateGhostValues on the same small vector, in ONE single-block kernel ----
  // (the projector coefficients C^H X of an apply: AtomCenterNonLocalOpContextFE::applyCconjtransOnX's all-reduce over the
  // ranks sharing an atom, src/basis/AtomCenterNonLocalOpContextFE.t.cpp:943-997 - a few rows per rank.)  The two exchanges
  // cost four launches with a grid-wide last-block hand-over and a system fence each; here one block walks the same wire
  // protocol (same buffers, flags, acknowledgements, same summation order at the owner), so a neighbour may run either form.
  __global__ void __launch_bounds__(1024)
  peer_acc_update_small_kernel(PeerDir da, PeerDir du, double *x, uint32_t n_owned, const uint32_t *gids, uint32_t n_ghost,
                               const uint32_t *send_rows, const uint32_t *acc_rows, const uint32_t *acc_off, const uint32_t *acc_pos,
                               uint32_t n_acc_rows, uint32_t B, uint32_t seqA, uint32_t seqU)
  {
    __shared__ int ok;
    double *       xg = x + (size_t)n_owned * B;
    if (threadIdx.x == 0)
      {
        // flow control of both directions: the previous messages have been consumed
        bool o = wait_words(da.ack, da.nDst, seqA - 1u, da.status);
        o      = o && wait_words(du.ack, du.nDst, seqU - 1u, du.status);
        if (!o)
          atomicExch(da.status, 1u);
        ok = o ? 1 : 0;
      }
    __syncthreads();
    // (1) accumulate, send side: my ghost rows go into their owners' buffers
    if (ok)
      for (size_t i = threadIdx.x; i < (size_t)da.nRows * B; i += blockDim.x)
        {
          const uint32_t k = (uint32_t)(i / B), c = (uint32_t)(i % B);
          const uint32_t s = da.seg[k];
          da.rbase[s][((size_t)da.rrowoff[s] + (k - da.segbegin[s])) * B + c] = xg[(size_t)gids[k] * B + c];
        }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      {
        if (ok)
          for (uint32_t s = 0; s < da.nDst; ++s)
            st_release_sys(da.rflag[s], seqA);
        // (2) accumulate, receive side: owned rows += buffer rows in buffer order
        const bool o = ok && wait_words(da.flag, da.nSrc, seqA, da.status);
        if (ok && !o)
          atomicExch(da.status, 1u);
        ok = o ? 1 : 0;
      }
    __syncthreads();
    if (ok)
      for (size_t i = threadIdx.x; i < (size_t)n_acc_rows * B; i += blockDim.x)
        {
          const uint32_t r = (uint32_t)(i / B), v = (uint32_t)(i % B);
          double *       y = x + (size_t)acc_rows[r] * B + v;
          double         t = *y;
          for (uint32_t e = acc_off[r]; e < acc_off[r + 1]; ++e)
            t += __ldcg(da.recv + (size_t)acc_pos[e] * B + v);
          *y = t;
        }
    __syncthreads();
    // (3) update, send side: my owned rows (now complete) go into the sharers' buffers
    if (ok)
      for (size_t i = threadIdx.x; i < (size_t)du.nRows * B; i += blockDim.x)
        {
          const uint32_t k = (uint32_t)(i / B), c = (uint32_t)(i % B);
          const uint32_t s = du.seg[k];
          du.rbase[s][((size_t)du.rrowoff[s] + (k - du.segbegin[s])) * B + c] = x[(size_t)send_rows[k] * B + c];
        }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      {
        if (ok)
          {
            for (uint32_t s = 0; s < da.nSrc; ++s)
              st_release_sys(da.rack[s], seqA); // the accumulate buffers have been read
            for (uint32_t s = 0; s < du.nDst; ++s)
              st_release_sys(du.rflag[s], seqU);
          }
        // (4) update, receive side
        const bool o = ok && wait_words(du.flag, du.nSrc, seqU, du.status);
        if (ok && !o)
          atomicExch(du.status, 1u);
        ok = o ? 1 : 0;
      }
    __syncthreads();
    if (ok)
      for (size_t i = threadIdx.x; i < (size_t)n_ghost * B; i += blockDim.x)
        xg[(size_t)gids[i / B] * B + (i % B)] = __ldcg(du.recv + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && ok)
      for (uint32_t s = 0; s < du.nSrc; ++s)
        st_release_sys(du.rack[s], seqU);
  }

  // at most this many doubles through one block (more: the multi-block kernels of the separate exchanges)
  constexpr size_t PEER_SMALL_DOUBLES = 96 * 1024;

  bool
  peer_acc_update_is_small(const Halo &h, uint32_t B)
  {
    return h.peer != nullptr && ((size_t)h.n_ghost + h.n_send + h.n_acc_rows) * B <= PEER_SMALL_DOUBLES;
  }

  int
  peer_halo_accumulate_update_small(hx_plan *p, Halo &h, double *X, uint32_t B)
  {
    PeerState *    s    = h.peer;
    const uint32_t seqA = ++s->seqA, seqU = ++s->seqU;
    peer_acc_update_small_kernel<<<1, 1024, 0, p->stream>>>(s->dirA, s->dirU, X, h.n_owned, h.d_ghost_local_ids.p, h.n_ghost,
                                                         h.d_owned_ids_for_targets.p, h.d_acc_rows.p, h.d_acc_off.p, h.d_acc_pos.p,
                                                         h.n_acc_rows, B, seqA, seqU);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // raised by a kernel whose wait timed out
  int
  peer_check_status(Halo &h)
  {
    if (!h.peer)
      return HX_OK;
    uint32_t st = 0;
    HX_CUDA(cudaMemcpy(&st, h.peer->wStatus, sizeof(st), cudaMemcpyDeviceToHost));
    HX_CHECK(st == 0, HX_ERR_COMM, "peer halo exchange timed out waiting for a neighbour");
    return HX_OK;
  }
} // namespace hx
