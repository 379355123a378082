// cell_kernel.cu — the dominant kernel: fused cell gather -> FP64 DMMA cell contraction -> deterministic scatter.
//
// Replaces, per reference H.X apply (src/ksdft/KohnShamOperatorContextFE.t.cpp:951-1199):
//   FECellWiseDataOperations::copyFieldToCellWiseData   (src/basis/FECellWiseDataOperations.t.cpp:58-86)
//   blasLapack::gemmStridedVarBatched (one dgemm_/cell)  (src/linearAlgebra/BlasLapack.t.cpp:388-436)
//   AtomCenterNonLocalOpContextFE::applyCOnVCconjtransX  (src/basis/AtomCenterNonLocalOpContextFE.t.cpp:998-1036)
//   FECellWiseDataOperations::addCellWiseDataToFieldData (src/basis/FECellWiseDataOperations.t.cpp:87-153)
//
// Design (B200, sm_100a):
//   * tcgen05 has no FP64 kind; the FP64 tensor path on sm_100a is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
//   * y_c[j,v] = sum_k A_c[j,k] xk[k,v] with A_c = [H_c | C_c^T]: the nonlocal C.(V C^H X) term is a K-extension
//     (rows n..n+nProj of the B operand are the V-scaled projector coefficients), so projector cells cost no
//     extra pass.
//   * A_c is pre-tiled once per reinit (pack_kernel) into the exact order the kernel consumes it: per cell a
//     linear stream of "stages"; stage (chunk, kc) holds, for the <= 8*MTW m-tiles of the chunk and KC k-steps, the
//     32-double DMMA A-fragments.  One cp.async.bulk (TMA, 1-D) per stage moves it into a shared-memory ring
//     guarded by full/empty mbarriers; compute warps read their fragments with conflict-free 256-B LDS.
//   * Persistent CTAs (two per SM), warp-specialised into four warpgroups whose register budgets are re-split with
//     setmaxnreg: 4 DMMA warps (k loop only), 8 scatter warps (everything with global-memory latency), one A-stream
//     warp (claims work items from a global counter and issues the bulk copies, running ahead across items) and one
//     gather warp (cp.async 16-B zero-filling gathers of the cell's rows of X / VCX into a padded shared tile).  The
//     accumulators of a finished m-chunk change hands through a swizzled shared-memory tile, so the DMMA warps never
//     wait for a global load (cell_apply_pipe_kernel below).
//   * Deterministic scatter without colour launches: work items are claimed in processing order; a cell adds
//     its rows into Y only after the immediately preceding toucher of each row has published an epoch stamp
//     (release/acquire through L2).  Per row the summation order is therefore ascending cell order - the
//     reference's CPU order - and bitwise reproducible.  The first toucher of a row stores instead of adding, so
//     Y needs no memset.  Neighbouring cells run close in time, so their shared rows of X and Y hit L2.
//     Rows shared by many cells (enrichment DoFs) go to a staging slot and are reduced in fixed order afterwards.
//   * The one-launch-per-colour variant (cells of one launch share no row) is kept as scatter_mode 1
//     (HXB200_SCATTER=coloured) for comparison.
#include <map>

#include "hx_internal.h"

namespace hx
{
  // CWARPS (m-tiles of a chunk / mtw) lives in hx_internal.h; with op->kc (k-steps of 4 per stage) it defines the packed layout
  constexpr int QD            = 16;              // item queue depth (A-stream warp -> gather / DMMA warps)
  constexpr int MAX_STAGES    = 8;
  constexpr uint32_t ITEM_END = 0xffffffffu;

  struct CellArgs
  {
    const double *  X;
    double *        Y;
    const double *  VCX;
    double *        stage;
    const double *  packed;
    const CellMeta *meta;
    const uint32_t *ids;
    const uint32_t *dest;
    const uint32_t *pids;
    const uint32_t *cell_list; // coloured mode: cells of this launch
    const ItemDesc *items;     // ordered mode: descriptors in processing order
    const uint32_t *wait_off;
    const uint32_t *wait_list;
    uint32_t *      flags;
    uint32_t *      counters;
    uint32_t        epoch;
    uint32_t        nItems;
    uint32_t        B;
    uint32_t        nBt;
    uint32_t        nStages;
    // Chebyshev epilogue (FUSE kernels only): see FuseArgs in hx_internal.h
    const double *  f_dinv;
    const double *  f_xprev;
    double *        f_out;
    double          f_a, f_b, f_c;
    uint32_t        f_has_c;   // c != 0 (the first degree has no Xprev term)
    uint32_t        f_discard; // Y tiles are whole 128-B lines (B % 16 == 0, aligned): dead partial sums are discarded
    uint32_t        shared_a;  // many cells stream the same packed matrix (hx_cellop_set_matrix_sharing): keep it in L2
    HaloK           halo;      // halo exchange inside the kernel (multi-rank plans, peer-memory transport); x_ready == nullptr: off
    unsigned long long *clk;   // [4 x nSamples ring]: clock64 / globaltimer at the start and end of CTA 0 (SM clock under load)
    uint32_t        kc;        // k-steps per stage of the packed stream (layout parameter, see pack_kernel)
  };

  __device__ __forceinline__ void
  dmma884(double &d0, double &d1, const double a, const double b)
  {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
  }

  __device__ __forceinline__ double
  ld_stream(const double *p)
  {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
  }

  // ---- mbarrier / bulk-copy / cp.async PTX wrappers ------------------------------------------------
  __device__ __forceinline__ uint32_t
  smem_u32(const void *p)
  {
    return (uint32_t)__cvta_generic_to_shared(p);
  }
  __device__ __forceinline__ void
  mbar_init(uint32_t bar, uint32_t count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  }
  __device__ __forceinline__ void
  mbar_arrive(uint32_t bar)
  {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ void
  mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void
  mbar_wait(uint32_t bar, uint32_t parity)
  {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity), "r"(0x989680u) // suspend-time hint: the warp sleeps in hardware until the phase completes
                 : "memory");
  }
  __device__ __forceinline__ void
  bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
  {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
  }
  __device__ __forceinline__ void
  bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
  {
    asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
  }
  __device__ __forceinline__ void
  cp_async_zfill16(uint32_t dst, const void *src, uint32_t src_bytes)
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
  }
  __device__ __forceinline__ void
  cp_async_zfill8(uint32_t dst, const void *src, uint32_t src_bytes)
  {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
  }
  __device__ __forceinline__ void
  cp_async_mbar_arrive_noinc(uint32_t bar)
  {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ uint32_t
  ld_volatile_shared(uint32_t addr)
  {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
  }
  __device__ __forceinline__ void
  st_volatile_shared(uint32_t addr, uint32_t v)
  {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
  }
  __device__ __forceinline__ uint32_t
  ld_acquire_gpu(const uint32_t *p)
  {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ void
  st_release_gpu(uint32_t *p, uint32_t v)
  {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  }
  // offset (doubles, relative to the cell's packed base) of stage (chunk starting at m-tile mc, k-chunk kc)
  __device__ __host__ __forceinline__ size_t
  stage_offset(int mc, int mtc, int kc, int nKC, int kcv)
  {
    return ((size_t)mc * nKC + (size_t)mtc * kc) * (kcv * 32);
  }

  // =================================================================================================
  // Pipelined ordered kernel (default): the contraction and the scatter of a work item run in DIFFERENT warps.
  // =================================================================================================
  // In the round-1 kernel the DMMA warps also performed the scatter epilogue of every item - destination codes,
  // the wait for the predecessors, Y / X / Xprev loads, stores - four serialised memory round trips per item during
  // which the CTA feeds no DMMA (ncu: tensor pipe 65 % active).  Here a CTA (two per SM) has four roles in four
  // warpgroups, with the register file re-split between them by setmaxnreg:
  //   * warps 0-3, DMMA: k loop only, 4 m-tiles x 4 n-tiles per warp (16 independent accumulators; 128 bytes of
  //     shared-memory fragments per DMMA instead of 192 with 2 m-tiles).  At the end of an m-chunk they park the
  //     accumulators in a shared-memory tile (swizzled, conflict-free for both sides) and start the next chunk / item.
  //   * warps 4-11, scatter: everything with global-memory latency.  Per chunk, threads 0-127 own one row each
  //     (destination code -> record {byte offset, flags}, dinv, L2 prefetch of Xprev) and threads 128-255 one
  //     predecessor each (ld.acquire.gpu on its stamp); all of that index data is requested one chunk ahead.  Then the
  //     256 threads read-modify-write the rows (16 bytes of a row per thread, straight-line predicated code over the
  //     records), and thread 0 publishes the item's stamp (st.release.gpu).
  //   * warp 12, A stream: claims items from the global counter (two ahead), fetches their descriptors, queues them
  //     for the other roles and issues one cp.async.bulk (TMA) per pipeline stage.
  //   * warp 13 (and 14, 15 in the kernels for cells of <= 64 DoFs, where the gather paces the pipeline), gather:
  //     zero-filling cp.async of the stage's rows of X (or of V C^H X) into the padded B tile; the 64-bit source address
  //     of a row is computed once, lane-parallel, and fetched with a shuffle.
  //   (warps of the fourth warpgroup without a role only complete it: they give their registers away and exit.)
  // Ordering / determinism: exactly the scheme above (one chain per row in processing order, first toucher stores,
  // last toucher of a fusable row applies the recurrence), so results are bit-identical to it.
  // Memory-model chain of a stamp: scatter threads st.cg -> bar.sync(scatter warps) -> thread 0 st.release.gpu;
  // consumer: scatter threads ld.acquire.gpu -> bar.sync -> ld.cg.
  constexpr int DWARPS       = 4;                // DMMA warps: 4 (or 2) m-tiles x 4 n-tiles each, 16 independent accumulators
  constexpr int DTHREADS     = DWARPS * 32;
  constexpr int SWARPS       = 8;
  constexpr int STHREADS     = SWARPS * 32;
  constexpr int SROWS        = 128;              // rows of the largest chunk (16 m-tiles)
  constexpr int PIPE_THREADS = (DWARPS + SWARPS + 4) * 32; // 512: launched with 64 registers per thread
  // register split (setmaxnreg): 128 x 112 + 256 x 56 + 128 x 32 = 512 x 64.  The k loop needs ~100 registers (16 accumulator
  // pairs + fragments); 56 per scatter thread hold two rows of Y / X / Xprev in flight without spilling (with 144 / 40 and one
  // row per batch the DMMA warps spent 20 % of their time waiting for the accumulator tile: 0.71 -> 0.63 ms at C2)
  constexpr int REG_DMMA     = 112;
  constexpr int REG_SCATTER  = 56;
  constexpr int REG_PRODUCER = 32;
  // build-time tuning knobs of the pipelined kernel (tools/build_variants.py compiles A/B libraries with other values)
#ifndef HX_PIPE_REGD
#define HX_PIPE_REGD REG_DMMA
#endif
#ifndef HX_PIPE_REGS
#define HX_PIPE_REGS REG_SCATTER
#endif
#ifndef HX_PIPE_RBF
#define HX_PIPE_RBF 2 // rows per scatter batch (loads in flight per thread), Chebyshev-epilogue kernels
#endif
#ifndef HX_PIPE_RBP
#define HX_PIPE_RBP 2 // same, plain apply
#endif
#ifndef HX_PIPE_NG_SMALL
#define HX_PIPE_NG_SMALL 3 // gather warps (1..3) of the one-m-tile-per-warp kernels (cells of <= 64 DoFs), see launch_cell_apply
#endif
#ifndef HX_PIPE_NG_LARGE
#define HX_PIPE_NG_LARGE 1 // ... of the two-m-tile kernels
#endif
#ifndef HX_PIPE_NACC
#define HX_PIPE_NACC 1 // accumulator tiles between the DMMA and the scatter warps
#endif
  constexpr int SMP_FULL     = 0;                // MAX_STAGES x 8
  constexpr int SMP_EMPTY    = 64;               // MAX_STAGES x 8
  constexpr int SMP_ACCFULL  = 128;              // (+16 per tile) accumulators of a chunk are in the shared tile
  constexpr int SMP_ACCFREE  = 136;              // (+16 per tile) the scatter warps have read them
  constexpr int SMP_Q        = 256;              // QD x 32 item queue
  constexpr int SMP_REC      = SMP_Q + QD * 32;  // 2 x SROWS x (16 + 8): row records {byte offset, flags} + dinv, current / next chunk
  constexpr int SMP_HEADER   = SMP_REC + 2 * SROWS * 24; // 6912

  __host__ __device__ constexpr int
  pipe_stage_bytes(int nt, int mtw, int kc)
  {
    return CWARPS * mtw * kc * 256 + 4 * kc * (nt * 8 + 4) * 8;
  }
  __host__ __device__ constexpr int
  pipe_acc_bytes(int nt, int mtw)
  {
    return CWARPS * mtw * 8 * nt * 8 * 8;
  }
  __device__ __forceinline__ void
  bar_scatter()
  {
    asm volatile("bar.sync 1, %0;" ::"n"(STHREADS) : "memory");
  }
  template <int R>
  __device__ __forceinline__ void
  reg_inc()
  {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
  }
  template <int R>
  __device__ __forceinline__ void
  reg_dec()
  {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
  }
  // queue entry of the pipelined kernel: 8 words
  struct PItem
  {
    uint32_t tag, ids_off, n, nproj, wait_off, nwait, proj_off;
  };
  __device__ __forceinline__ void
  pitem_load(uint32_t addr, PItem &it)
  {
    __threadfence_block();
    uint32_t pad;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(pad), "=r"(it.ids_off), "=r"(it.n), "=r"(it.nproj)
                 : "r"(addr)
                 : "memory");
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(it.wait_off), "=r"(it.nwait), "=r"(it.proj_off), "=r"(pad)
                 : "r"(addr + 16)
                 : "memory");
  }

  template <int NT, int MTW, int KCT, bool VEC, int MINB, bool FUSE, int RB, int NACC = 1, int REGD = REG_DMMA, int REGS = REG_SCATTER, int NG = 1>
  __global__ void __launch_bounds__(PIPE_THREADS, MINB) cell_apply_pipe_kernel(const CellArgs a)
  {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int  BT      = NT * 8;
    constexpr int  LDX     = BT + 4;
    constexpr int  MPC     = CWARPS * MTW; // m-tiles per chunk (the packed layout's)
    constexpr int  MPW     = MPC / DWARPS; // m-tiles per DMMA warp
    constexpr int  KROWS   = 4 * KCT;      // rows of X per stage
    constexpr int  A_BYTES = CWARPS * MTW * KCT * 256;
    constexpr int  S_BYTES = pipe_stage_bytes(NT, MTW, KCT);
    constexpr int  ACC_B   = pipe_acc_bytes(NT, MTW);
    constexpr int  ROW_B   = BT * 8; // bytes of one row of the accumulator tile
    const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t accb  = sbase + SMP_HEADER;
    const uint32_t stgb  = accb + NACC * ACC_B;
    const uint32_t NS    = a.nStages;

    if (tid == 0)
      {
        for (uint32_t s = 0; s < NS; ++s)
          {
            mbar_init(sbase + SMP_FULL + 8 * s, 1 + 32 * NG); // the A warp's arrive.expect_tx + the gather lanes (X)
            mbar_init(sbase + SMP_EMPTY + 8 * s, DWARPS);
          }
        for (int b = 0; b < NACC; ++b)
          {
            mbar_init(sbase + SMP_ACCFULL + 16 * b, DWARPS);
            mbar_init(sbase + SMP_ACCFREE + 16 * b, SWARPS);
          }
        for (int q = 0; q < QD; ++q)
          st_volatile_shared(sbase + SMP_Q + 32 * q, 0u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
    pdl_wait();
    pdl_launch();
    __syncthreads();

    if (a.halo.x_ready != nullptr && blockIdx.x < a.halo.n_halo_ctas)
      {
        // ---------------- halo CTAs: receive side of updateGhostValues, before they join the contraction ----------------
        // (MPICommunicatorP2P::updateGhostValuesEnd, src/utils/MPICommunicatorP2P.t.cpp:226-273: wait, unpack.)  The other
        // CTAs are already contracting interior cells; the cells that read ghost rows wait for x_ready.
        __shared__ int halo_ok;
        const HaloK &  hk = a.halo;
        if (tid == 0)
          {
            bool ok = wait_words(hk.flagU, hk.nSrcU, hk.seqU, hk.status);
            ok      = ok && wait_words(hk.ackA, hk.nDstA, hk.seqA - 1u, hk.status);
            if (!ok)
              atomicExch(hk.status, 1u);
            halo_ok = ok ? 1 : 0;
          }
        __syncthreads();
        if (halo_ok && hk.do_unpack)
          {
            const uint32_t B   = a.B;
            const size_t   tot = (size_t)hk.n_ghost * B;
            for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < tot; i += (size_t)hk.n_halo_ctas * blockDim.x)
              {
                const uint32_t g = hk.unpack_ids[i / B];
                if (!(g & 0x80000000u)) // a constrained ghost row keeps the value its parents gave it
                  hk.xghost[(size_t)g * B + (i % B)] = __ldcg(hk.recvU + i);
              }
          }
        __threadfence();
        __syncthreads();
        if (tid == 0)
          {
            const uint32_t done = atomicAdd(hk.counter, 1u);
            if (done == hk.n_halo_ctas - 1)
              {
                hk.counter[0] = 0u;
                __threadfence_system();
                if (halo_ok && hk.do_unpack) // a timed-out exchange acknowledges nothing
                  for (uint32_t s = 0; s < hk.nSrcU; ++s)
                    st_release_sys(hk.rackU[s], hk.seqU);
                st_release_gpu(hk.x_ready, a.epoch);
              }
          }
      }

    if (warp >= DWARPS + SWARPS)
      {
        reg_dec<REG_PRODUCER>();
        if (warp == DWARPS + SWARPS)
          {
            // ---------------- A-stream warp (one lane): claims, descriptors, queue, one bulk copy per stage ----------------
            if (lane != 0)
              return;
            uint32_t stage = 0, ph = 0;
            uint64_t evict_first;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(evict_first));
            const bool once  = (a.nBt == 1) && !a.shared_a;
            auto       fetch = [&](uint32_t w_, uint4 &lo, uint4 &hi) {
              lo = hi = make_uint4(0u, 0u, 0u, 0u);
              if (w_ < a.nItems)
                {
                  const uint4 *q = reinterpret_cast<const uint4 *>(a.items + w_ / a.nBt);
                  lo = __ldg(q), hi = __ldg(q + 1);
                }
            };
            uint32_t w_cur = atomicAdd(a.counters, 1u);
            uint4    lo_cur, hi_cur;
            fetch(w_cur, lo_cur, hi_cur);
            uint32_t w_next = atomicAdd(a.counters, 1u);
            for (uint32_t it = 0;; ++it)
              {
                const uint32_t slot = sbase + SMP_Q + 32 * (it % QD);
                while (ld_volatile_shared(slot) != 0u)
                  {
                  }
                if (w_cur >= a.nItems)
                  {
                    st_volatile_shared(slot, ITEM_END);
                    break;
                  }
                // ItemDesc: lo = {h_off lo, h_off hi, ids_off, n}, hi = {nproj, proj_off, wait_off, nwait}
                const uint4    lo = lo_cur, hi = hi_cur;
                const uint32_t w  = w_cur;
                asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(slot + 16), "r"(hi.z), "r"(hi.w), "r"(hi.y),
                             "r"(0u)
                             : "memory");
                asm volatile("st.volatile.shared.v2.u32 [%0], {%1,%2};" ::"r"(slot + 8), "r"(lo.w), "r"(hi.x) : "memory");
                st_volatile_shared(slot + 4, lo.z);
                __threadfence_block();
                st_volatile_shared(slot, w + 1u);
                w_cur = w_next;
                fetch(w_cur, lo_cur, hi_cur);
                w_next = atomicAdd(a.counters, 1u);

                const int     n = (int)lo.w, ktot = n + (int)hi.x;
                const int     nKC  = (ktot + KROWS - 1) / KROWS;
                const int     nMt  = (n + 7) >> 3;
                const double *srcA = a.packed + (((unsigned long long)lo.y << 32) | lo.x);
                for (int mc = 0; mc < nMt; mc += MPC)
                  {
                    const uint32_t bytes = (uint32_t)min(MPC, nMt - mc) * KCT * 256u;
                    for (int kc = 0; kc < nKC; ++kc)
                      {
                        mbar_wait(sbase + SMP_EMPTY + 8 * stage, ph ^ 1u);
                        const uint32_t full = sbase + SMP_FULL + 8 * stage;
                        mbar_arrive_expect_tx(full, bytes);
                        if (once)
                          bulk_g2s_hint(stgb + stage * S_BYTES, srcA, bytes, full, evict_first);
                        else
                          bulk_g2s(stgb + stage * S_BYTES, srcA, bytes, full);
                        srcA += bytes / 8;
                        if (++stage == NS)
                          {
                            stage = 0;
                            ph ^= 1u;
                          }
                      }
                  }
              }
          }
        else if (warp <= DWARPS + SWARPS + NG)
          {
            const int gw = warp - (DWARPS + SWARPS + 1); // gather warp gw issues every NG-th row instruction of a stage
            // ---------------- gather warp: rows of X (and of V C^H X) into the B tile of every stage ----------------
            constexpr int CPR = VEC ? BT / 2 : BT; // copies per row (<= 32)
            constexpr int RPI = 32 / CPR;          // rows per warp instruction
            constexpr int SPB = 32 / KROWS;        // stages per block of 32 gathered rows
            static_assert(KROWS <= 32 && 32 % KROWS == 0 && KROWS % RPI == 0, "stage rows");
            const int cc = lane % CPR;
            const int rr = lane / CPR;
            uint32_t  stage = 0, ph = 0;
            for (uint32_t it = 0;; ++it)
              {
                const uint32_t slot = sbase + SMP_Q + 32 * (it % QD);
                uint32_t       tag;
                while ((tag = ld_volatile_shared(slot)) == 0u)
                  {
                  }
                if (tag == ITEM_END)
                  break;
                PItem d;
                pitem_load(slot, d);
                const uint32_t w = tag - 1u;
                const int      n = (int)d.n, ktot = n + (int)d.nproj;
                if ((d.nwait & HX_ITEM_BOUNDARY) && a.halo.x_ready != nullptr)
                  {
                    // this cell reads ghost rows of X: they are in place once the halo CTAs have published the stamp (which
                    // also means the owners have consumed the previous accumulate message: this cell may push into their buffers)
                    if (lane == 0)
                      while (ld_acquire_gpu(a.halo.x_ready) != a.epoch)
                        {
                        }
                    __syncwarp();
                  }
                // start address of row k of the B operand (0 = zero row): X rows of the cell, then its rows of V C^H X
                auto row_addr = [&](int k) -> unsigned long long {
                  if (k >= ktot)
                    return 0ull;
                  if (k < n)
                    return (unsigned long long)(a.X + (size_t)__ldg(a.ids + d.ids_off + k) * a.B);
                  return (unsigned long long)(a.VCX + (size_t)__ldg(a.pids + d.proj_off + (k - n)) * a.B);
                };
                const uint32_t b0    = (w % a.nBt) * BT;
                const int      nKC   = (ktot + KROWS - 1) / KROWS;
                const int      nMt   = (n + 7) >> 3;
                const uint32_t col   = b0 + (VEC ? cc * 2 : cc);
                const bool     colok = col < a.B;
                for (int mc = 0; mc < nMt; mc += MPC)
                  {
                    unsigned long long raddr = row_addr(lane), raddr_next = row_addr(32 + lane);
                    for (int kc = 0; kc < nKC; ++kc)
                      {
                        if (kc && (kc % SPB) == 0)
                          {
                            raddr      = raddr_next;
                            raddr_next = row_addr((kc / SPB + 1) * 32 + lane);
                          }
                        mbar_wait(sbase + SMP_EMPTY + 8 * stage, ph ^ 1u);
                        const uint32_t xs = stgb + stage * S_BYTES + A_BYTES;
#pragma unroll
                        for (int r = 0; r < KROWS; r += RPI)
                          {
                            if (NG > 1 && (r / RPI) % NG != gw)
                              continue;
                            const unsigned long long rp = __shfl_sync(0xffffffffu, raddr, (kc % SPB) * KROWS + r + rr);
                            const bool               ok = (rp != 0ull) && colok;
                            const double *           src = ok ? reinterpret_cast<const double *>(rp) + col : a.X;
                            const uint32_t           dst = xs + ((uint32_t)(r + rr) * LDX + (VEC ? cc * 2 : cc)) * 8u;
                            if (VEC)
                              cp_async_zfill16(dst, src, ok ? 16u : 0u);
                            else
                              cp_async_zfill8(dst, src, ok ? 8u : 0u);
                          }
                        cp_async_mbar_arrive_noinc(sbase + SMP_FULL + 8 * stage);
                        if (++stage == NS)
                          {
                            stage = 0;
                            ph ^= 1u;
                          }
                      }
                  }
              }
          }
        return;
      }
    else if (warp >= DWARPS)
      {
        reg_dec<REGS>();
        // ---------------- scatter warps ----------------
        constexpr int TPR   = BT / 2;         // threads per row (16 bytes each)
        constexpr int RPP   = STHREADS / TPR; // rows per pass of the 256 threads
        constexpr int NPASS = (MPC * 8 + RPP - 1) / RPP;
        constexpr int LPR   = (BT * 8 + 127) / 128; // 128-byte lines per row of a tile
        constexpr int RBE   = RB < NPASS ? RB : NPASS; // rows per batch
        static_assert(NPASS % RBE == 0, "rows per batch");
        static_assert(MPC * 8 <= SROWS, "one scatter thread per row of a chunk");
        const int      st   = tid - DTHREADS;
        const int      q    = st % TPR;
        const int      rrow = st / TPR;
        const uint32_t B    = a.B;
        const bool     fc   = FUSE && (a.f_has_c != 0u);
        uint32_t       g    = 0; // chunk counter (phase of the accumulator hand-over, record buffer)
        // Index data of a chunk - thread st < 128: destination code of row st; thread st >= 128: entry st - 128 of the
        // item's predecessor list - is requested one chunk ahead, while the rows of the current chunk are being
        // read-modify-written, so a chunk starts with it in a register.
        bool     have_next = false;
        uint32_t idx_next  = 0xffffffffu;
        for (uint32_t it = 0;; ++it)
          {
            const uint32_t slot = sbase + SMP_Q + 32 * (it % QD);
            uint32_t       tag;
            while ((tag = ld_volatile_shared(slot)) == 0u)
              {
              }
            if (tag == ITEM_END)
              break;
            PItem info;
            pitem_load(slot, info);
            const uint32_t w   = tag - 1u, bt = w % a.nBt;
            const uint32_t nwait = info.nwait & ~HX_ITEM_BOUNDARY;
            const int      n   = (int)info.n;
            const int      nMt = (n + 7) >> 3;
            const uint32_t b0  = bt * BT;
            const uint32_t col = b0 + q * 2;
            const bool     c0ok = col < B, c1ok = col + 1 < B;
            for (int mc = 0; mc < nMt; mc += MPC, ++g)
              {
                const int      rbase = mc * 8;
                const int      nrows = min(MPC * 8, n - rbase);
                const uint32_t rbuf  = sbase + SMP_REC + (g & 1u) * (SROWS * 24);
                uint32_t       idx;
                if (have_next)
                  idx = idx_next;
                else if (st < SROWS)
                  idx = (st < nrows) ? __ldg(a.dest + info.ids_off + rbase + st) : 0xffffffffu;
                else
                  idx = (mc == 0 && (uint32_t)(st - SROWS) < nwait) ? __ldg(a.wait_list + info.wait_off + (st - SROWS)) : 0xffffffffu;
                have_next = false;
                double dinv_r = 0.0;
                if (st < SROWS)
                  {
                    if (FUSE && idx != 0xffffffffu && (idx & (HX_DEST_LASTF | HX_DEST_STAGED)) == HX_DEST_LASTF)
                      {
                        dinv_r = __dmul_rn(a.f_a, __ldg(a.f_dinv + HX_DEST_ROW(idx))); // s = a*dinv of the row
                        if (fc)
                          {
                            // this row's last toucher will read xprev[row, tile]: pull the lines into L2 now
                            const double *xp = a.f_xprev + (size_t)HX_DEST_ROW(idx) * B + b0;
#pragma unroll
                            for (int l = 0; l < LPR; ++l)
                              if (b0 + l * 16 < B)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + l * 16));
                          }
                      }
                  }
                // request the index data of the next chunk (same item, or the next item if it is queued already)
                if (mc + MPC < nMt)
                  {
                    if (st < SROWS)
                      idx_next = (st < n - rbase - MPC * 8) ? __ldg(a.dest + info.ids_off + rbase + MPC * 8 + st) : 0xffffffffu;
                    else
                      idx_next = 0xffffffffu;
                    have_next = true;
                  }
                else
                  {
                    const uint32_t nslot = sbase + SMP_Q + 32 * ((it + 1) % QD);
                    const uint32_t ntag  = ld_volatile_shared(nslot);
                    if (ntag != 0u && ntag != ITEM_END)
                      {
                        PItem ni;
                        pitem_load(nslot, ni);
                        if (st < SROWS)
                          idx_next = ((uint32_t)st < ni.n) ? __ldg(a.dest + ni.ids_off + st) : 0xffffffffu;
                        else
                          idx_next = ((uint32_t)(st - SROWS) < (ni.nwait & ~HX_ITEM_BOUNDARY)) ? __ldg(a.wait_list + ni.wait_off + (st - SROWS)) : 0xffffffffu;
                        have_next = true;
                      }
                  }
                if (st >= SROWS)
                  {
                    // the immediately preceding toucher of every row of this cell must have stored (acquire)
                    if (mc == 0)
                      {
                        if (idx != 0xffffffffu)
                          {
                            const uint32_t *f = a.flags + (size_t)idx * a.nBt + bt;
                            while (ld_acquire_gpu(f) != a.epoch)
                              {
                              }
                          }
                        for (uint32_t i = (uint32_t)st; i < nwait; i += SROWS)
                          {
                            const uint32_t *f = a.flags + (size_t)__ldg(a.wait_list + info.wait_off + i) * a.nBt + bt;
                            while (ld_acquire_gpu(f) != a.epoch)
                              {
                              }
                          }
                      }
                  }
                else
                  {
                    // record of this thread's row: flags bit 0 valid, 1 add (read the partial sum first), 2 last toucher of
                    // a fusable row, 3 staged
                    const bool valid  = idx != 0xffffffffu;
                    const bool staged = valid && (idx & HX_DEST_STAGED);
                    const bool lastf  = FUSE && valid && (idx & (HX_DEST_LASTF | HX_DEST_STAGED)) == HX_DEST_LASTF;
                    // 4 push: last toucher of a ghost row whose sum goes straight into its owner's accumulate buffer
                    const bool push = valid && !staged && (idx & HX_DEST_PUSH) && a.halo.x_ready != nullptr;
                    const uint32_t fl = !valid ? 0u :
                                                 (1u | ((!staged && !(idx & HX_DEST_FIRST)) ? 2u : 0u) | (lastf ? 4u : 0u) | (staged ? 8u : 0u) |
                                                  (push ? 16u : 0u));
                    const unsigned long long ro =
                      (unsigned long long)(staged ? (idx & 0x7fffffffu) : HX_DEST_ROW(idx)) * B * 8ull;
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rbuf + 16 * st), "r"((uint32_t)ro),
                                 "r"((uint32_t)(ro >> 32)), "r"(fl), "r"(0u)
                                 : "memory");
                    // second slot of the record: a*dinv of a fusable row, or the owner-side address of a pushed ghost row
                    unsigned long long aux = (unsigned long long)__double_as_longlong(dinv_r);
                    if (push)
                      {
                        const uint32_t j = HX_DEST_ROW(idx) - a.halo.n_owned;
                        aux = (unsigned long long)a.halo.push_base[j] + (unsigned long long)a.halo.push_row[j] * B * 8ull;
                      }
                    asm volatile("st.shared.u64 [%0], %1;" ::"r"(rbuf + SROWS * 16 + 8 * st), "l"(aux) : "memory");
                  }
                bar_scatter(); // records of the chunk are in shared memory; the predecessors have stored
                const uint32_t ab = (NACC == 2) ? (g & 1u) : 0u, aph = (NACC == 2) ? ((g >> 1) & 1u) : (g & 1u);
                mbar_wait(sbase + SMP_ACCFULL + 16 * ab, aph);
                const uint32_t acct = accb + ab * ACC_B;
                const unsigned long long colb = (unsigned long long)col * 8ull;
#pragma unroll 1
                for (int p0 = 0; p0 < NPASS; p0 += RBE)
                  {
                    uint32_t           fl[RBE];
                    unsigned long long t[RBE];
                    double2            y[RBE], xc[RBE], xp[RBE], av[RBE];
#pragma unroll
                    for (int u = 0; u < RBE; ++u)
                      {
                        const int r = (p0 + u) * RPP + rrow;
                        uint32_t  lo, hi, pad;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(lo), "=r"(hi), "=r"(fl[u]), "=r"(pad)
                                     : "r"(rbuf + 16 * r)
                                     : "memory");
                        t[u] = (((unsigned long long)hi << 32) | lo) + colb;
                        if (!c0ok || (MPC * 8 % RPP != 0 && r >= MPC * 8))
                          fl[u] = 0u;
                      }
#pragma unroll
                    for (int u = 0; u < RBE; ++u)
                      {
                        y[u] = xc[u] = xp[u] = make_double2(0.0, 0.0);
                        if (fl[u] & 2u)
                          {
                            if (VEC)
                              y[u] = __ldcg(reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(a.Y) + t[u]));
                            else
                              {
                                y[u].x = __ldcg(reinterpret_cast<const double *>(reinterpret_cast<const char *>(a.Y) + t[u]));
                                if (c1ok)
                                  y[u].y = __ldcg(reinterpret_cast<const double *>(reinterpret_cast<const char *>(a.Y) + t[u]) + 1);
                              }
                          }
                        if (FUSE && (fl[u] & 4u))
                          {
                            xc[u] = __ldcg(reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(a.X) + t[u]));
                            if (fc)
                              xp[u] = __ldcg(reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(a.f_xprev) + t[u]));
                          }
                      }
                    // accumulators of these rows (swizzled tile, see the DMMA warps)
#pragma unroll
                    for (int u = 0; u < RBE; ++u)
                      {
                        const int      r    = (p0 + u) * RPP + rrow;
                        const uint32_t addr = acct + (uint32_t)r * ROW_B + (uint32_t)(((q >> 2) ^ (r & (NT - 1))) * 64 + (q & 3) * 16);
                        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(av[u].x), "=d"(av[u].y) : "r"(addr) : "memory");
                      }
#pragma unroll
                    for (int u = 0; u < RBE; ++u)
                      {
                        double2 v;
                        v.x = y[u].x + av[u].x;
                        v.y = y[u].y + av[u].y;
                        char *ptr = reinterpret_cast<char *>((fl[u] & 8u) ? a.stage : a.Y) + t[u];
                        if (FUSE)
                          {
                            if (fl[u] & 4u)
                              {
                                // last toucher of a fusable row: the final (H X)[row, tile] is here,
                                // out = (a*dinv)*(H X) + (b*X + c*Xprev)   [cheb_combine_diag]
                                const int r = (p0 + u) * RPP + rrow;
                                double    dv;
                                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(dv) : "r"(rbuf + SROWS * 16 + 8 * r) : "memory");
                                double2 z;
                                z.x = __dmul_rn(a.f_b, xc[u].x), z.y = __dmul_rn(a.f_b, xc[u].y);
                                if (fc) // cheb_z
                                  z.x = __fma_rn(a.f_c, xp[u].x, z.x), z.y = __fma_rn(a.f_c, xp[u].y, z.y);
                                // the partial sums of this row are dead: drop their dirty L2 lines (whole 128-B lines only)
                                if ((fl[u] & 2u) && a.f_discard && (q & 7) == 0)
                                  asm volatile("discard.global.L2 [%0], 128;" ::"l"(ptr) : "memory");
                                v.x = __fma_rn(dv, v.x, z.x);
                                v.y = __fma_rn(dv, v.y, z.y);
                                ptr = reinterpret_cast<char *>(a.f_out) + t[u];
                              }
                          }
                        if (fl[u] & 1u)
                          {
                            if (VEC)
                              __stcg(reinterpret_cast<double2 *>(ptr), v);
                            else
                              {
                                __stcg(reinterpret_cast<double *>(ptr), v.x);
                                if (c1ok)
                                  __stcg(reinterpret_cast<double *>(ptr) + 1, v.y);
                              }
                          }
                        if (fl[u] & 16u)
                          {
                            // accumulateAddLocallyOwned, send side (MPICommunicatorP2P.t.cpp:288-380): this cell was the last
                            // local toucher of a ghost row - its sum goes into the owner's receive buffer over NVLink now
                            const int          r = (p0 + u) * RPP + rrow;
                            unsigned long long ra;
                            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(ra) : "r"(rbuf + SROWS * 16 + 8 * r) : "memory");
                            char *rp = reinterpret_cast<char *>(ra + colb);
                            if (VEC)
                              *reinterpret_cast<double2 *>(rp) = v;
                            else
                              {
                                *reinterpret_cast<double *>(rp) = v.x;
                                if (c1ok)
                                  *(reinterpret_cast<double *>(rp) + 1) = v.y;
                              }
                          }
                      }
                  }
                // the tile may be overwritten by the next chunk
                __syncwarp();
                if (lane == 0)
                  mbar_arrive(sbase + SMP_ACCFREE + 16 * ab);
              }
            // every row of the item is stored: publish its stamp (release, cumulative at gpu scope)
            bar_scatter();
            if (st == 0)
              {
                st_release_gpu(a.flags + w, a.epoch);
                st_volatile_shared(slot, 0u); // queue slot free again
              }
          }
      }
    else
      {
        reg_inc<REGD>();
        // ---------------- DMMA warps ----------------
        unsigned long long c0 = 0, t0 = 0;
        if (tid == 0 && blockIdx.x == 0 && a.clk)
          {
            c0 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
          }
        uint32_t stage = 0, ph = 0, g = 0;
        for (uint32_t it = 0;; ++it)
          {
            const uint32_t slot = sbase + SMP_Q + 32 * (it % QD);
            uint32_t       tag;
            while ((tag = ld_volatile_shared(slot)) == 0u)
              {
              }
            if (tag == ITEM_END)
              break;
            __threadfence_block();
            uint32_t n_, np_;
            asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(n_), "=r"(np_) : "r"(slot + 8) : "memory");
            const int n    = (int)n_;
            const int nKC  = (n + (int)np_ + KROWS - 1) / KROWS;
            const int nMt  = (n + 7) >> 3;
            const int xoff = A_BYTES / 8 + (lane & 3) * LDX + (lane >> 2); // B fragment inside a stage
            for (int mc = 0; mc < nMt; mc += MPC, ++g)
              {
                const int  mtc    = min(MPC, nMt - mc);
                const int  mtl0   = warp * MPW;
                const bool active = mtl0 < mtc;
                int        aoff[MPW];
#pragma unroll
                for (int j = 0; j < MPW; ++j)
                  aoff[j] = min(mtl0 + j, mtc - 1) * (KCT * 32) + lane;
                double acc[MPW][NT][2];
#pragma unroll
                for (int j = 0; j < MPW; ++j)
#pragma unroll
                  for (int t = 0; t < NT; ++t)
                    acc[j][t][0] = acc[j][t][1] = 0.0;
                for (int kc = 0; kc < nKC; ++kc)
                  {
                    mbar_wait(sbase + SMP_FULL + 8 * stage, ph);
                    if (active)
                      {
                        const double *St = reinterpret_cast<const double *>(smem_raw + SMP_HEADER + NACC * ACC_B + (size_t)stage * S_BYTES);
                        const double *xr = St + xoff;
#pragma unroll
                        for (int ks = 0; ks < KCT; ++ks)
                          {
                            double af[MPW], bf[NT];
#pragma unroll
                            for (int j = 0; j < MPW; ++j)
                              af[j] = St[aoff[j] + ks * 32];
#pragma unroll
                            for (int t = 0; t < NT; ++t)
                              bf[t] = xr[ks * 4 * LDX + t * 8];
#pragma unroll
                            for (int j = 0; j < MPW; ++j)
#pragma unroll
                              for (int t = 0; t < NT; ++t)
                                dmma884(acc[j][t][0], acc[j][t][1], af[j], bf[t]);
                          }
                      }
                    __syncwarp();
                    if (lane == 0)
                      mbar_arrive(sbase + SMP_EMPTY + 8 * stage);
                    if (++stage == NS)
                      {
                        stage = 0;
                        ph ^= 1u;
                      }
                  }
                // park the accumulators for the scatter warps (they have released the tile of the previous chunk)
                const uint32_t ab = (NACC == 2) ? (g & 1u) : 0u, aph = (NACC == 2) ? ((g >> 1) & 1u) : (g & 1u);
                mbar_wait(sbase + SMP_ACCFREE + 16 * ab, aph ^ 1u);
                if (active)
                  {
#pragma unroll
                    for (int j = 0; j < MPW; ++j)
                      if (mtl0 + j < mtc)
                        {
                          const int r = (mtl0 + j) * 8 + (lane >> 2);
#pragma unroll
                          for (int t = 0; t < NT; ++t)
                            {
                              const uint32_t addr = accb + ab * ACC_B + (uint32_t)r * ROW_B + (uint32_t)((t ^ (r & (NT - 1))) * 64 + (lane & 3) * 16);
                              asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(acc[j][t][0]), "d"(acc[j][t][1]) : "memory");
                            }
                        }
                  }
                __syncwarp();
                if (lane == 0)
                  mbar_arrive(sbase + SMP_ACCFULL + 16 * ab);
              }
          }
        if (tid == 0 && blockIdx.x == 0 && a.clk)
          {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            a.clk[0] += clock64() - c0; // SM cycles and nanoseconds CTA 0 spent in the kernel, summed over launches
            a.clk[1] += t1 - t0;
          }
        // last CTA out resets the work counters for the next launch
        if (tid == 0)
          {
            __threadfence();
            const uint32_t done = atomicAdd(a.counters + 1, 1u);
            if (done == gridDim.x - 1)
              {
                a.counters[0] = 0u;
                a.counters[1] = 0u;
                __threadfence();
              }
          }
      }
  }

  // =================================================================================================
  // One launch per colour (scatter_mode 1): one CTA = one cell x one column tile, A streamed from HBM
  // straight into registers.
  // =================================================================================================
  constexpr int CELL_THREADS = 256;
  constexpr int CELL_WARPS   = CELL_THREADS / 32;
  constexpr int PD           = 4; // register prefetch depth (k-steps) of the A stream

  template <int NT, int MTW, bool VEC, int MINB>
  __global__ void __launch_bounds__(CELL_THREADS, MINB) cell_apply_coloured_kernel(const CellArgs a)
  {
    extern __shared__ __align__(16) double xs[];
    constexpr int BT  = NT * 8;
    constexpr int LDX = BT + 4;
    constexpr int MPC = CELL_WARPS * MTW;
    const int     tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const uint32_t bt   = blockIdx.x % a.nBt;
    const uint32_t ci   = blockIdx.x / a.nBt;
    const uint32_t cell = a.cell_list[ci];
    const CellMeta cm   = a.meta[cell];
    const int      n    = (int)cm.n;
    const int      ktot = n + (int)cm.nproj;
    const int      KCv  = (int)a.kc;
    const int      nKC  = (ktot + 4 * KCv - 1) / (4 * KCv);
    const int      Kp   = nKC * 4 * KCv;
    const int      nK   = Kp >> 2;
    const int      nMt  = (n + 7) >> 3;
    const uint32_t B    = a.B;
    const uint32_t b0   = bt * BT;

    // ---- gather the cell's rows of X (and of V C^H X) into shared memory ----
    {
      constexpr int  PAIRS = BT / 2;
      constexpr int  RPP   = CELL_THREADS / PAIRS; // rows per pass
      const int      pr    = tid % PAIRS;
      const int      r0    = tid / PAIRS;
      const uint32_t col   = b0 + pr * 2;
      constexpr int  U     = 4;
      for (int kb = r0; kb < Kp; kb += RPP * U)
        {
          const double *src[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
            {
              const int k = kb + u * RPP;
              src[u]      = nullptr;
              if (k < n)
                src[u] = a.X + (size_t)__ldg(a.ids + cm.ids_off + k) * B;
              else if (k < ktot)
                src[u] = a.VCX + (size_t)__ldg(a.pids + cm.proj_off + (k - n)) * B;
            }
          double2 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
            {
              v[u] = make_double2(0.0, 0.0);
              if (src[u] != nullptr)
                {
                  if (VEC)
                    {
                      if (col < B)
                        v[u] = *reinterpret_cast<const double2 *>(src[u] + col);
                    }
                  else
                    {
                      if (col < B)
                        v[u].x = src[u][col];
                      if (col + 1 < B)
                        v[u].y = src[u][col + 1];
                    }
                }
            }
#pragma unroll
          for (int u = 0; u < U; ++u)
            {
              const int k = kb + u * RPP;
              if (k < Kp)
                *reinterpret_cast<double2 *>(xs + (size_t)k * LDX + pr * 2) = v[u];
            }
        }
    }
    __syncthreads();

    const double *Abase = a.packed + cm.h_off + lane;
    const double *xrow  = xs + (size_t)(lane & 3) * LDX + (lane >> 2);
    for (int mc = 0; mc < nMt; mc += MPC)
      {
        const int mtc  = min(MPC, nMt - mc);
        const int mtl0 = warp * MTW;
        if (mtl0 >= mtc)
          break;
        // fragment (m-tile mtl, k-step k) of this chunk sits at stage_offset(mc,mtc,k/KC) + (mtl*KC + k%KC)*32
        const double *Ap[MTW];
#pragma unroll
        for (int j = 0; j < MTW; ++j)
          Ap[j] = Abase + stage_offset(mc, mtc, 0, nKC, KCv) + (size_t)min(mtl0 + j, mtc - 1) * (KCv * 32);
        const size_t kc_stride = (size_t)mtc * (KCv * 32);
        auto         frag      = [&](int j, int k) { return Ap[j] + (size_t)(k / KCv) * kc_stride + (size_t)(k % KCv) * 32; };

        double acc[MTW][NT][2];
#pragma unroll
        for (int j = 0; j < MTW; ++j)
#pragma unroll
          for (int t = 0; t < NT; ++t)
            acc[j][t][0] = acc[j][t][1] = 0.0;

        double af[PD][MTW];
#pragma unroll
        for (int i = 0; i < PD; ++i)
#pragma unroll
          for (int j = 0; j < MTW; ++j)
            af[i][j] = (i < nK) ? ld_stream(frag(j, i)) : 0.0;

        for (int ks = 0; ks < nK; ks += PD)
          {
#pragma unroll
            for (int i = 0; i < PD; ++i)
              {
                const int k = ks + i;
                if (k < nK)
                  {
                    double ac[MTW];
#pragma unroll
                    for (int j = 0; j < MTW; ++j)
                      ac[j] = af[i][j];
                    const int kn = k + PD;
                    if (kn < nK)
                      {
#pragma unroll
                        for (int j = 0; j < MTW; ++j)
                          af[i][j] = ld_stream(frag(j, kn));
                      }
                    const double *xr = xrow + (size_t)k * 4 * LDX;
                    double        b[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t)
                      b[t] = xr[t * 8];
#pragma unroll
                    for (int j = 0; j < MTW; ++j)
#pragma unroll
                      for (int t = 0; t < NT; ++t)
                        dmma884(acc[j][t][0], acc[j][t][1], ac[j], b[t]);
                  }
              }
          }

        // ---- scatter-add (colour-exclusive rows: plain RMW on the zeroed Y; shared rows: staging slot) ----
#pragma unroll
        for (int j = 0; j < MTW; ++j)
          {
            const int r = (mc + mtl0 + j) * 8 + (lane >> 2);
            if (mtl0 + j < mtc && r < n)
              {
                const uint32_t d      = __ldg(a.dest + cm.ids_off + r);
                const bool     staged = (d & HX_DEST_STAGED) != 0;
                double *       dst    = staged ? a.stage + (size_t)(d & 0x7fffffffu) * B : a.Y + (size_t)HX_DEST_ROW(d) * B;
                const bool     add    = !staged;
#pragma unroll
                for (int t = 0; t < NT; ++t)
                  {
                    const uint32_t col = b0 + t * 8 + (lane & 3) * 2;
                    if (VEC)
                      {
                        if (col < B)
                          {
                            double2 *p = reinterpret_cast<double2 *>(dst + col);
                            double2  y = add ? *p : make_double2(0.0, 0.0);
                            y.x += acc[j][t][0];
                            y.y += acc[j][t][1];
                            *p = y;
                          }
                      }
                    else
                      {
                        if (col < B)
                          dst[col] = (add ? dst[col] : 0.0) + acc[j][t][0];
                        if (col + 1 < B)
                          dst[col + 1] = (add ? dst[col + 1] : 0.0) + acc[j][t][1];
                      }
                  }
              }
          }
      }
  }

  // -------------------------------------------------------------------------------------------------
  // pack: raw row-major n x n cell matrices (+ column-major nProj x n projector matrices) -> the stage stream.
  // For chunk mc (mtc m-tiles), k-chunk kc, local m-tile mtl, k-step ks, lane:
  //   packed[stage_offset(mc,mtc,kc) + ((mtl*KC + ks)*32 + lane)] = A[(mc+mtl)*8 + lane/4][(kc*KC + ks)*4 + lane%4]
  // with KC = op->kc k-steps per stage (2 or 4, chosen per operator in pack_cell_matrices)
  __global__ void
  pack_kernel(const double *            raw,
              unsigned long long        raw_base,
              const unsigned long long *raw_off,
              const double *            cellC,
              const unsigned long long *c_off,
              const CellMeta *          meta,
              double *                  packed,
              uint32_t                  cell_begin,
              int                       mpc,
              int                       KCv)
  {
    const uint32_t cell = cell_begin + blockIdx.x;
    const CellMeta cm   = meta[cell];
    const int      n = (int)cm.n, np = (int)cm.nproj;
    const int      nKC = (n + np + 4 * KCv - 1) / (4 * KCv), nMt = (n + 7) >> 3;
    const double * H   = raw ? raw + (raw_off[cell] - raw_base) : nullptr; // null: structure only (H part zero)
    const double * Cc  = (np > 0) ? cellC + c_off[cell] : nullptr;
    double *       out = packed + cm.h_off;
    const size_t   tot = (size_t)nMt * nKC * KCv * 32;
    for (size_t idx = threadIdx.x; idx < tot; idx += blockDim.x)
      {
        // decode idx in stream order
        const size_t per_chunk = (size_t)mpc * nKC * KCv * 32;
        const int    ch        = (int)(idx / per_chunk);
        const int    mc        = ch * mpc;
        const int    mtc       = min(mpc, nMt - mc);
        size_t       rem       = idx - (size_t)ch * per_chunk;
        const int    kc        = (int)(rem / ((size_t)mtc * KCv * 32));
        rem -= (size_t)kc * mtc * KCv * 32;
        const int mtl  = (int)(rem / (KCv * 32));
        const int ks   = (int)((rem / 32) % KCv);
        const int lane = (int)(rem & 31);
        const int r = (mc + mtl) * 8 + (lane >> 2), k = (kc * KCv + ks) * 4 + (lane & 3);
        double    v = 0.0;
        if (r < n)
          {
            if (k < n)
              v = H ? H[(size_t)r * n + k] : 0.0;
            else if (k < n + np)
              v = Cc[(size_t)(k - n) + (size_t)r * np];
          }
        out[idx] = v;
      }
  }

  // ---- identical cell matrices share one packed copy (hx_cellop_set_matrix_sharing) ----
  // order-independent 2 x 64-bit fingerprint of each cell's packed stream
  __global__ void
  hash_cells_kernel(const double *packed, const CellMeta *meta, const unsigned long long *len, unsigned long long *hash)
  {
    __shared__ unsigned long long s1[256], s2[256];
    const uint32_t                cell = blockIdx.x;
    const unsigned long long *    src  = reinterpret_cast<const unsigned long long *>(packed + meta[cell].h_off);
    const unsigned long long      n    = len[cell];
    unsigned long long            h1 = 0, h2 = 0;
    for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x)
      {
        const unsigned long long v = src[i];
        h1 += v * ((0x9E3779B97F4A7C15ull * (i + 1)) | 1ull);
        unsigned long long t = v + i * 0xBF58476D1CE4E5B9ull;
        t ^= t >> 31;
        t *= 0x94D049BB133111EBull;
        h2 ^= t ^ (t >> 29);
      }
    s1[threadIdx.x] = h1, s2[threadIdx.x] = h2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
      {
        if ((int)threadIdx.x < o)
          {
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] ^= s2[threadIdx.x + o];
          }
        __syncthreads();
      }
    if (threadIdx.x == 0)
      hash[2 * (size_t)cell] = s1[0], hash[2 * (size_t)cell + 1] = s2[0];
  }
  // bitwise comparison of every candidate with its representative
  __global__ void
  verify_shared_kernel(const double *packed, const unsigned long long *off_self, const unsigned long long *off_rep,
                       const unsigned long long *len, const uint32_t *cells, uint32_t *mismatch)
  {
    const uint32_t            c = cells[blockIdx.x];
    const unsigned long long *a = reinterpret_cast<const unsigned long long *>(packed + off_self[blockIdx.x]);
    const unsigned long long *b = reinterpret_cast<const unsigned long long *>(packed + off_rep[blockIdx.x]);
    const unsigned long long  n = len[c];
    bool                      bad = false;
    for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x)
      bad = bad || (a[i] != b[i]);
    if (bad)
      atomicAdd(mismatch, 1u);
  }

  static int
  share_identical_matrices(hx_op *op)
  {
    hx_plan *p = op->plan;
    op->n_unique = p->C;
    if (p->C < 2)
      return HX_OK;
    std::vector<unsigned long long> len(p->C);
    for (uint32_t c = 0; c < p->C; ++c)
      {
        const CellMeta &m  = op->h_meta[c];
        const uint32_t  kr = 4u * (uint32_t)op->kc;
        const uint32_t  Kp = (m.n + m.nproj + kr - 1) / kr * kr, Mp = (m.n + 7) & ~7u;
        len[c]             = (unsigned long long)Kp * Mp;
      }
    DevBuf<unsigned long long> d_len, d_hash;
    HX_TRY(d_len.upload(len));
    HX_TRY(d_hash.alloc(2 * (size_t)p->C));
    hash_cells_kernel<<<p->C, 256, 0, p->stream>>>(op->d_packed.p, op->d_meta.p, d_len.p, d_hash.p);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    std::vector<unsigned long long> hash(2 * (size_t)p->C);
    HX_CUDA(cudaMemcpyAsync(hash.data(), d_hash.p, hash.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    HX_CUDA(cudaStreamSynchronize(p->stream));
    struct Key
    {
      unsigned long long a, b, l;
      bool
      operator<(const Key &o) const
      {
        return a != o.a ? a < o.a : (b != o.b ? b < o.b : l < o.l);
      }
    };
    std::map<Key, uint32_t>         rep;
    std::vector<uint32_t>           cand;
    std::vector<unsigned long long> off_self, off_rep;
    std::vector<uint32_t>           rep_of(p->C);
    for (uint32_t c = 0; c < p->C; ++c)
      {
        const Key k{hash[2 * (size_t)c], hash[2 * (size_t)c + 1], len[c]};
        auto      it = rep.find(k);
        if (it == rep.end())
          {
            rep[k]    = c;
            rep_of[c] = c;
          }
        else
          {
            rep_of[c] = it->second;
            cand.push_back(c);
            off_self.push_back(op->h_meta[c].h_off);
            off_rep.push_back(op->h_meta[it->second].h_off);
          }
      }
    if (cand.empty())
      return HX_OK;
    DevBuf<unsigned long long> d_self, d_rep;
    DevBuf<uint32_t>           d_cand, d_mis;
    HX_TRY(d_self.upload(off_self));
    HX_TRY(d_rep.upload(off_rep));
    HX_TRY(d_cand.upload(cand));
    HX_TRY(d_mis.alloc(1));
    HX_CUDA(cudaMemsetAsync(d_mis.p, 0, sizeof(uint32_t), p->stream));
    verify_shared_kernel<<<(unsigned)cand.size(), 256, 0, p->stream>>>(op->d_packed.p, d_self.p, d_rep.p, d_len.p, d_cand.p,
                                                                       d_mis.p);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    uint32_t mis = 0;
    HX_CUDA(cudaMemcpyAsync(&mis, d_mis.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    HX_CUDA(cudaStreamSynchronize(p->stream));
    if (mis != 0)
      return HX_OK; // a fingerprint collision: keep every cell's own copy
    for (uint32_t c : cand)
      op->h_meta[c].h_off = op->h_meta[rep_of[c]].h_off;
    op->n_unique = (uint32_t)rep.size();
    HX_TRY(op->d_meta.upload(op->h_meta));
    std::vector<ItemDesc> items(p->C);
    for (uint32_t w = 0; w < p->C; ++w)
      {
        const CellMeta &m = op->h_meta[p->h_order[w]];
        ItemDesc &      d = items[w];
        d.h_off = m.h_off, d.ids_off = m.ids_off, d.n = m.n, d.nproj = m.nproj, d.proj_off = m.proj_off;
        d.wait_off = p->h_wait_off[w], d.nwait = (p->h_wait_off[w + 1] - p->h_wait_off[w]) | (p->h_boundary[p->h_order[w]] ? HX_ITEM_BOUNDARY : 0u);
      }
    HX_TRY(op->d_items.upload(items));
    return HX_OK;
  }

  int
  pack_cell_matrices(hx_op *op, const double *raw, int on_device)
  {
    hx_plan *p = op->plan;
    // m-tiles per warp: small cells (n <= 64) keep all 8 DMMA warps busy with one m-tile each
    op->mtw       = (p->max_n <= 64) ? 1 : 2;
    // k-steps per pipeline stage: seven stages either way (8 KB of A + 8 rows of X with 16 m-tiles per chunk, 8 KB + 16
    // rows with 8)
    op->kc = (op->mtw == 2) ? 2 : 4;
    const int mpc = CWARPS * op->mtw;
    size_t    tot = 0;
    op->h_meta.resize(p->C);
    uint32_t poff = 0;
    op->max_kp = op->max_mp = 0;
    for (uint32_t c = 0; c < p->C; ++c)
      {
        CellMeta &m = op->h_meta[c];
        m.n         = p->h_ncd[c];
        m.ids_off   = p->h_cell_off[c];
        m.nproj     = op->has_nl ? op->h_ncp[c] : 0;
        m.proj_off  = poff;
        poff += m.nproj;
        const uint32_t kr = 4u * (uint32_t)op->kc;
        const uint32_t Kp = (m.n + m.nproj + kr - 1) / kr * kr, Mp = (m.n + 7) & ~7u;
        m.h_off = tot;
        tot += (size_t)Kp * Mp;
        op->max_kp = Kp > op->max_kp ? Kp : op->max_kp;
        op->max_mp = Mp > op->max_mp ? Mp : op->max_mp;
      }
    HX_TRY(op->d_meta.upload(op->h_meta));
    {
      std::vector<ItemDesc> items(p->C);
      for (uint32_t w = 0; w < p->C; ++w)
        {
          const CellMeta &m = op->h_meta[p->h_order[w]];
          ItemDesc &      d = items[w];
          d.h_off = m.h_off, d.ids_off = m.ids_off, d.n = m.n, d.nproj = m.nproj, d.proj_off = m.proj_off;
          d.wait_off = p->h_wait_off[w], d.nwait = (p->h_wait_off[w + 1] - p->h_wait_off[w]) | (p->h_boundary[p->h_order[w]] ? HX_ITEM_BOUNDARY : 0u);
        }
      HX_TRY(op->d_items.upload(items));
    }
    if (op->packed_doubles != tot || op->d_packed.p == nullptr)
      {
        HX_TRY(op->d_packed.alloc(tot));
        op->packed_doubles = tot;
      }
    // raw offsets
    std::vector<unsigned long long> raw_off(p->C + 1, 0);
    for (uint32_t c = 0; c < p->C; ++c)
      raw_off[c + 1] = raw_off[c] + (unsigned long long)p->h_ncd[c] * p->h_ncd[c];
    DevBuf<unsigned long long> d_raw_off;
    HX_TRY(d_raw_off.upload(raw_off.data(), raw_off.size()));
    HX_CUDA(cudaDeviceSynchronize());

    if (on_device)
      {
        if (p->C)
          {
            pack_kernel<<<p->C, 256, 0, p->stream>>>(raw, 0ull, d_raw_off.p, op->d_cell_c.p, op->d_c_off.p,
                                                     op->d_meta.p, op->d_packed.p, 0, mpc, op->kc);
            p->launches++;
          }
        HX_CUDA(cudaGetLastError());
        HX_CUDA(cudaStreamSynchronize(p->stream));
      }
    else
      {
        // upload in chunks of <= 256 MB to bound the temporary
        const unsigned long long chunk_max = 32ull << 20; // doubles
        DevBuf<double>           tmp;
        uint32_t                 c0 = 0;
        while (c0 < p->C)
          {
            uint32_t c1 = c0;
            while (c1 < p->C && (raw_off[c1 + 1] - raw_off[c0] <= chunk_max || c1 == c0))
              ++c1;
            const unsigned long long cnt = raw_off[c1] - raw_off[c0];
            if (tmp.n < cnt)
              HX_TRY(tmp.alloc(cnt));
            HX_CUDA(cudaMemcpyAsync(tmp.p, raw + raw_off[c0], cnt * sizeof(double), cudaMemcpyHostToDevice, p->stream));
            pack_kernel<<<c1 - c0, 256, 0, p->stream>>>(tmp.p, raw_off[c0], d_raw_off.p, op->d_cell_c.p,
                                                        op->d_c_off.p, op->d_meta.p, op->d_packed.p, c0, mpc, op->kc);
            p->launches++;
            HX_CUDA(cudaGetLastError());
            HX_CUDA(cudaStreamSynchronize(p->stream));
            c0 = c1;
          }
      }
    op->have_matrices = true;
    op->n_unique      = p->C;
    if (op->share_identical)
      HX_TRY(share_identical_matrices(op));
    return HX_OK;
  }

  // ---- event bracket around the cell-kernel launches of one apply (read back in hx_plan_cell_kernel_time_ms) ----
  static int
  timing_begin(hx_plan *p, cudaEvent_t *e1)
  {
    *e1 = nullptr;
    if (!p->timing)
      return HX_OK;
    if (p->ev_used + 2 > p->ev_pool.size())
      for (int i = 0; i < 64; ++i)
        {
          cudaEvent_t e;
          HX_CUDA(cudaEventCreate(&e));
          p->ev_pool.push_back(e);
        }
    cudaEvent_t e0 = p->ev_pool[p->ev_used++];
    *e1            = p->ev_pool[p->ev_used++];
    HX_CUDA(cudaEventRecord(e0, p->stream));
    return HX_OK;
  }

  template <int NT, int MTW, bool VEC, int MINB>
  static int
  launch_colours(hx_op *op, const CellArgs &base, size_t smem)
  {
    hx_plan *p = op->plan;
    auto     k = cell_apply_coloured_kernel<NT, MTW, VEC, MINB>;
    HX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e1;
    HX_TRY(timing_begin(p, &e1));
    for (uint32_t col = 0; col < p->n_colours; ++col)
      {
        const uint32_t nc = p->h_colour_off[col + 1] - p->h_colour_off[col];
        if (nc == 0)
          continue;
        CellArgs a  = base;
        a.cell_list = p->d_colour_cells.p + p->h_colour_off[col];
        k<<<nc * a.nBt, CELL_THREADS, smem, p->stream>>>(a);
        p->launches++;
        p->cell_launches++;
      }
    if (e1)
      HX_CUDA(cudaEventRecord(e1, p->stream));
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  template <int NT, int MTW, int KCT, bool VEC, int MINB, bool FUSE, int RB, int NACC = 1, int REGD = REG_DMMA, int REGS = REG_SCATTER, int NG = 1>
  static int
  launch_pipe(hx_op *op, CellArgs a)
  {
    hx_plan *    p      = op->plan;
    auto         k      = cell_apply_pipe_kernel<NT, MTW, KCT, VEC, MINB, FUSE, RB, NACC, REGD, REGS, NG>;
    const size_t budget = 227 * 1024 / MINB - 1024; // per CTA (228 KB per SM, 1 KB reserved by the runtime per CTA)
    const size_t fixed  = SMP_HEADER + (size_t)NACC * pipe_acc_bytes(NT, MTW);
    size_t       ns     = (budget - fixed) / (size_t)pipe_stage_bytes(NT, MTW, KCT);
    if (ns > MAX_STAGES)
      ns = MAX_STAGES;
    const size_t smem = fixed + ns * (size_t)pipe_stage_bytes(NT, MTW, KCT);
    a.nStages         = (uint32_t)ns;
    HX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    HX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, PIPE_THREADS, smem));
    HX_CHECK(occ >= 1, HX_ERR_UNSUPPORTED, "pipelined cell kernel does not fit on an SM (smem %zu)", smem);
    uint32_t grid = (uint32_t)(p->sm_count * std::min(occ, MINB));
    if (grid > a.nItems)
      grid = a.nItems;
    cudaEvent_t e1;
    HX_TRY(timing_begin(p, &e1));
    HX_CUDA(launch_pdl(k, grid, PIPE_THREADS, smem, p->stream, a));
    p->launches++;
    p->cell_launches++;
    if (e1)
      HX_CUDA(cudaEventRecord(e1, p->stream));
    return HX_OK;
  }

  int
  launch_cell_apply(hx_op *op, const double *X, double *Y, uint32_t B, const FuseArgs *fuse, bool *fused_applied, const HaloK *halo)
  {
    if (fused_applied)
      *fused_applied = false;
    hx_plan *p = op->plan;
    HX_CHECK(op->have_matrices, HX_ERR_INVALID, "cell operator has no matrices (call hx_cellop_set_matrices)");
    if (p->C == 0)
      return HX_OK;
    CellArgs a;
    memset(&a, 0, sizeof(a));
    a.X         = X;
    a.Y         = Y;
    a.VCX       = op->d_cx.p;
    a.stage     = p->d_stage.p;
    a.packed    = op->d_packed.p;
    a.meta      = op->d_meta.p;
    a.ids       = p->d_ids.p;
    a.dest      = p->d_dest.p;
    a.pids      = op->d_pids.p;
    a.cell_list = p->d_order.p;
    a.items     = op->d_items.p;
    a.wait_off  = p->d_wait_off.p;
    a.wait_list = p->d_wait_list.p;
    a.flags     = p->d_flags.p;
    a.counters  = p->d_counters.p;
    a.B         = B;
    a.shared_a  = (op->n_unique * 2u < p->C) ? 1u : 0u;
    a.kc        = (uint32_t)op->kc;
    a.clk       = p->timing ? p->d_clk.p : nullptr;
    if (halo)
      a.halo = *halo;
    // column tile: widest of {8,16,32} columns that B needs and shared memory allows
    int  nt      = B > 16 ? 4 : (B > 8 ? 2 : 1);
    auto xtile_of = [&](int nt_) { return (size_t)op->max_kp * (nt_ * 8 + 4) * sizeof(double); };
    const bool vec = (B % 2 == 0) && ((((uintptr_t)X | (uintptr_t)Y | (uintptr_t)a.VCX | (uintptr_t)a.stage) & 15) == 0);
    const bool ordered = (p->scatter_mode == 0);
    if (ordered)
      {
        a.nBt    = (B + nt * 8 - 1) / (nt * 8);
        a.nItems = p->C * a.nBt;
        a.epoch  = ++p->epoch;
        if (p->epoch == 0xfffffff0u)
          {
            // epoch wrap: clear the stamps (never reached in practice)
            HX_CUDA(cudaMemsetAsync(p->d_flags.p, 0, p->d_flags.n * sizeof(uint32_t), p->stream));
            p->epoch = 0;
            a.epoch  = ++p->epoch;
          }
        HX_CHECK((size_t)a.nItems <= p->d_flags.n, HX_ERR_INVALID, "flag array too small");
        // the Chebyshev epilogue exists for the vectorised variants (B even, 16-B aligned operands)
        const bool fz = fuse != nullptr && vec && fuse->dinv && fuse->out && (fuse->c == 0.0 || fuse->xprev) &&
                        ((((uintptr_t)fuse->out | (uintptr_t)fuse->xprev) & 15) == 0);
        if (fz)
          {
            a.f_dinv = fuse->dinv, a.f_xprev = fuse->xprev ? fuse->xprev : fuse->out, a.f_out = fuse->out;
            a.f_a = fuse->a, a.f_b = fuse->b, a.f_c = fuse->c;
            a.f_has_c = (fuse->c != 0.0) ? 1u : 0u;
            a.f_discard = (B % (uint32_t)(nt * 8) == 0 && nt >= 2 && (((uintptr_t)Y) & 127) == 0) ? 1u : 0u;
            if (fused_applied)
              *fused_applied = true;
          }
        // rows per scatter batch (loads in flight per thread): what the registers of a scatter thread hold
        constexpr int RBF = HX_PIPE_RBF, RBP = HX_PIPE_RBP;
        // small cells (<= 64 DoFs: 4 stages of 16 rows per item) are paced by the gather warp's cp.async issue rate: three
        // gather warps share the rows of a stage there (order 3, B = 32: 0.505 -> 0.539 of the roofline); with 125-DoF
        // cells one is enough
#define HX_NG(MTW_) ((MTW_) == 1 ? HX_PIPE_NG_SMALL : HX_PIPE_NG_LARGE)
#define HX_PIPE(NT_, MTW_, KC_)                                                                     \
  (fz ? launch_pipe<NT_, MTW_, KC_, true, 2, true, RBF, HX_PIPE_NACC, HX_PIPE_REGD, HX_PIPE_REGS, HX_NG(MTW_)>(op, a) :                 \
        (vec ? launch_pipe<NT_, MTW_, KC_, true, 2, false, RBP, HX_PIPE_NACC, HX_PIPE_REGD, HX_PIPE_REGS, HX_NG(MTW_)>(op, a) :         \
               launch_pipe<NT_, MTW_, KC_, false, 2, false, RBP, HX_PIPE_NACC, HX_PIPE_REGD, HX_PIPE_REGS, HX_NG(MTW_)>(op, a)))
#define HX_PIPE_NT(MTW_, KC_) (nt == 4 ? HX_PIPE(4, MTW_, KC_) : (nt == 2 ? HX_PIPE(2, MTW_, KC_) : HX_PIPE(1, MTW_, KC_)))
        if (op->mtw == 1)
          return HX_PIPE_NT(1, 4);
        return HX_PIPE_NT(2, 2);
#undef HX_PIPE_NT
#undef HX_PIPE
#undef HX_NG
      }
    while (nt > 1 && xtile_of(nt) > 200 * 1024)
      nt >>= 1;
    HX_CHECK(xtile_of(nt) <= 220 * 1024, HX_ERR_UNSUPPORTED, "cell with %u DoFs does not fit shared memory", op->max_kp);
    a.nBt             = (B + nt * 8 - 1) / (nt * 8);
    const size_t smem = xtile_of(nt);
#define HX_COL(NT_, MTW_) \
  (vec ? launch_colours<NT_, MTW_, true, 2>(op, a, smem) : launch_colours<NT_, MTW_, false, 2>(op, a, smem))
    if (op->mtw == 1)
      switch (nt)
        {
          case 4:
            return HX_COL(4, 1);
          case 2:
            return HX_COL(2, 1);
          default:
            return HX_COL(1, 1);
        }
    switch (nt)
      {
        case 4:
          return HX_COL(4, 2);
        case 2:
          return HX_COL(2, 2);
        default:
          return HX_COL(1, 2);
      }
#undef HX_COL
  }

  // -------------------------------------------------------------------------------------------------
  // Nonlocal phase A: per projector cell, CXcell[p,v] = sum_k C_c[p + k*nP] x_c[k,v]
  // (AtomCenterNonLocalOpContextFE::applyCconjtransOnX, src/basis/AtomCenterNonLocalOpContextFE.t.cpp:889-942),
  // written to a per-(cell,projector) staging row; reduced per projector row in ascending cell order afterwards.
  constexpr int NLA_KCH  = 256; // rows of the cell gathered per pass
  constexpr int NLA_MAXS = 8;   // projector groups of 8 per cell held in registers: nProj_c <= 64
  __global__ void __launch_bounds__(256)
  nl_phase_a_kernel(const double *X, const uint32_t *ids, const uint32_t *nl_cells, const CellMeta *meta,
                    const double *cellC, const unsigned long long *c_off, double *cx_stage, uint32_t B)
  {
    extern __shared__ __align__(16) double xs[]; // [NLA_KCH][32] rows of X, then [8][NLA_KCH] projector coefficients
    pdl_wait();
    pdl_launch();
    const uint32_t cell = nl_cells[blockIdx.x];
    const uint32_t b0   = blockIdx.y * 32;
    const CellMeta cm   = meta[cell];
    const int      n = (int)cm.n, np = (int)cm.nproj;
    const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t col  = b0 + lane;
    const double * Cc   = cellC + c_off[cell];
    double *       cs   = xs + (size_t)NLA_KCH * 32;
    double         acc[NLA_MAXS];
#pragma unroll
    for (int g = 0; g < NLA_MAXS; ++g)
      acc[g] = 0.0;
    // the rows of the cell pass through shared memory NLA_KCH at a time (any cell size fits); every projector's sum runs
    // over k in ascending order across the passes, so the result does not depend on the chunking
    for (int k0 = 0; k0 < n; k0 += NLA_KCH)
      {
        const int kc = min(NLA_KCH, n - k0);
        __syncthreads();
        // gather: 4 rows per warp in flight (row id -> row is a dependent pair of loads)
        for (int kk = warp; kk < kc; kk += 32)
          {
            uint32_t r[4];
            double   v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              r[u] = (kk + 8 * u < kc) ? __ldg(ids + cm.ids_off + k0 + kk + 8 * u) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              v[u] = (kk + 8 * u < kc && col < B) ? __ldg(X + (size_t)r[u] * B + col) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (kk + 8 * u < kc)
                xs[(kk + 8 * u) * 32 + lane] = v[u];
          }
        // C_c (np x n, projector index fastest) is staged 8 projectors at a time, so the dot products read it as
        // broadcasts instead of a chain of dependent global loads
#pragma unroll
        for (int g = 0; g < NLA_MAXS; ++g)
          {
            const int p0 = g * 8;
            if (p0 >= np)
              break;
            const int pc = min(8, np - p0);
            __syncthreads();
            for (int i = threadIdx.x; i < pc * kc; i += blockDim.x)
              {
                const int k = i / pc, q = i % pc;
                cs[q * NLA_KCH + k] = __ldg(Cc + (size_t)(p0 + q) + (size_t)(k0 + k) * np);
              }
            __syncthreads();
            if (warp < pc)
              {
                const double *cr = cs + warp * NLA_KCH;
                double        s_ = acc[g];
                for (int k = 0; k < kc; ++k)
                  s_ += cr[k] * xs[k * 32 + lane];
                acc[g] = s_;
              }
          }
      }
#pragma unroll
    for (int g = 0; g < NLA_MAXS; ++g)
      if (g * 8 + warp < np && col < B)
        cx_stage[(size_t)(cm.proj_off + g * 8 + warp) * B + col] = acc[g];
  }

  // CX[row,:] = V[row] * sum over the row's staging slots (fixed order)   [reduce + (single rank) V scale]
  __global__ void
  nl_reduce_kernel(const double *cx_stage, const uint32_t *pr_off, const uint32_t *pr_slots, const double *V,
                   double *CX, uint32_t n_rows, uint32_t B, int scale)
  {
    pdl_wait();
    pdl_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_rows * B)
      return;
    const uint32_t row = (uint32_t)(i / B), v = (uint32_t)(i % B);
    double         s = 0.0;
    // fixed ascending order; the loads of NL_ILP consecutive slots are in flight together (only the adds are serial)
    constexpr int  NL_ILP = 16;
    const uint32_t end    = pr_off[row + 1];
    for (uint32_t e0 = pr_off[row]; e0 < end; e0 += NL_ILP)
      {
        uint32_t sl[NL_ILP];
        double   t[NL_ILP];
#pragma unroll
        for (int u = 0; u < NL_ILP; ++u)
          sl[u] = pr_slots[min(e0 + u, end - 1)]; // clamped: every load is unconditional and in range
#pragma unroll
        for (int u = 0; u < NL_ILP; ++u)
          t[u] = cx_stage[(size_t)sl[u] * B + v];
#pragma unroll
        for (int u = 0; u < NL_ILP; ++u)
          if (e0 + u < end)
            s += t[u];
      }
    CX[i] = scale ? V[row] * s : s;
  }

  int
  launch_nl_phase_a(hx_op *op, const double *X, uint32_t B)
  {
    hx_plan *p = op->plan;
    if (!op->has_nl)
      return HX_OK;
    const uint32_t ncell = (uint32_t)op->h_nl_cells.size();
    if (ncell)
      {
        const size_t smem = (size_t)NLA_KCH * (32 + 8) * sizeof(double);
        HX_CUDA(cudaFuncSetAttribute(nl_phase_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(ncell, (B + 31) / 32);
        HX_CUDA(launch_pdl(nl_phase_a_kernel, grid, 256, smem, p->stream, X, p->d_ids.p, op->d_nl_cells.p, op->d_meta.p,
                           op->d_cell_c.p, op->d_c_off.p, op->d_cx_stage.p, B));
        p->launches++;
      }
    const bool   single = (p->nranks == 1);
    const size_t tot    = (size_t)op->n_proj_local * B;
    if (tot)
      {
        HX_CUDA(launch_pdl(nl_reduce_kernel, (unsigned)((tot + 255) / 256), 256, 0, p->stream, op->d_cx_stage.p, op->d_pr_off.p,
                           op->d_pr_slots.p, op->d_v.p, op->d_cx.p, op->n_proj_local, B, single ? 1 : 0));
        p->launches++;
      }
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }
} // namespace hx
