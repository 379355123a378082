// hx_internal.h — internal structures of libhxb200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/hxb200.h"

namespace hx
{
  void set_error(const char *fmt, ...);

#define HX_CUDA(call)                                                                    \
  do                                                                                     \
    {                                                                                    \
      cudaError_t e_ = (call);                                                           \
      if (e_ != cudaSuccess)                                                             \
        {                                                                                \
          hx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
          return HX_ERR_CUDA;                                                            \
        }                                                                                \
    }                                                                                    \
  while (0)

#define HX_CHECK(cond, code, ...)   \
  do                                \
    {                               \
      if (!(cond))                  \
        {                           \
          hx::set_error(__VA_ARGS__); \
          return (code);            \
        }                           \
    }                               \
  while (0)

#define HX_CHECK_B(plan, B)                                                                        \
  HX_CHECK((plan) != nullptr, HX_ERR_INVALID, "null plan");                                        \
  HX_CHECK((B) >= 1 && (B) <= (plan)->max_block, HX_ERR_INVALID, "B = %u outside [1, max_block = %u]", (B), \
           (plan)->max_block)

#define HX_TRY(call)      \
  do                      \
    {                     \
      int r_ = (call);    \
      if (r_ != HX_OK)    \
        return r_;        \
    }                     \
  while (0)

  // Programmatic dependent launch (default for single-rank plans; HXB200_PDL=0 / 1 forces it off / on): the kernels of one H.X apply / filter degree are launched with
  // cudaLaunchAttributeProgrammaticStreamSerialization, so the launch and block scheduling of kernel k+1 overlap the
  // tail of kernel k.  Every such kernel executes pdl_wait() before its first global-memory access (it returns once
  // the preceding grid has completed and its writes are visible: the stream-order semantics are unchanged) and
  // pdl_launch() right after it, which lets the next kernel in the stream be scheduled behind this one.  Launched
  // without the attribute both instructions are no-ops.
  bool pdl_enabled();
#ifdef __CUDACC__
  __device__ __forceinline__ void
  pdl_wait()
  {
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  __device__ __forceinline__ void
  pdl_launch()
  {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  template <typename... KArgs, typename... Args>
  inline cudaError_t
  launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = grid;
    cfg.blockDim           = block;
    cfg.dynamicSmemBytes   = smem;
    cfg.stream             = stream;
    cudaLaunchAttribute at[1];
    at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                        = at;
    cfg.numAttrs                                     = pdl_enabled() ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  }
#endif

  template <typename T>
  struct DevBuf
  {
    T *    p = nullptr;
    size_t n = 0;
    DevBuf()               = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &
    operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void
    release()
    {
      if (p)
        cudaFree(p);
      p = nullptr;
      n = 0;
    }
    int
    alloc(size_t count)
    {
      release();
      n = count;
      if (count == 0)
        return HX_OK;
      cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
      if (e != cudaSuccess)
        {
          p = nullptr;
          n = 0;
          set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
          return HX_ERR_NOMEM;
        }
      return HX_OK;
    }
    int
    upload(const T *h, size_t count)
    {
      int r = alloc(count);
      if (r != HX_OK)
        return r;
      if (count)
        HX_CUDA(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
      return HX_OK;
    }
    int
    upload(const std::vector<T> &v)
    {
      return upload(v.data(), v.size());
    }
  };

  struct PeerState; // peer.cu: NVLink peer-memory transport of one halo
  void peer_destroy(PeerState *s);

  // device-side copy of one MPIPatternP2P + the communicator buffers of MPICommunicatorP2P
  struct Halo
  {
    PeerState *peer        = nullptr; // set once the peer-memory transport is up (lazy, collective)
    bool       peer_failed = false;   // IPC mapping unavailable: NCCL send/recv stays in charge
    Halo() = default;
    Halo(const Halo &) = delete;
    Halo &
    operator=(const Halo &) = delete;
    ~Halo()
    {
      if (peer)
        peer_destroy(peer);
    }
    uint32_t              n_owned = 0, n_ghost = 0;
    std::vector<uint32_t> ghost_procs, ghost_ranges, target_procs, target_counts;
    uint32_t              n_send = 0; // total owned indices for targets
    DevBuf<uint32_t>      d_ghost_local_ids, d_owned_ids_for_targets;
    // accumulate side: unique owned rows receiving halo contributions -> ordered buffer positions
    uint32_t              n_acc_rows = 0;
    DevBuf<uint32_t>      d_acc_rows, d_acc_off, d_acc_pos;
    DevBuf<double>        d_send, d_recv; // NCCL transport only; sized for max_block on first use
    size_t                buf_doubles = 0;
    int
    ensure_staging()
    {
      if (d_send.n < buf_doubles)
        HX_TRY(d_send.alloc(buf_doubles));
      if (d_recv.n < buf_doubles)
        HX_TRY(d_recv.alloc(buf_doubles));
      return HX_OK;
    }
    int
    init(const hx_halo_desc &h, uint32_t max_block);
  };

  struct CellMeta
  {
    unsigned long long h_off;    // offset (doubles) of this cell's packed matrix
    uint32_t           ids_off;  // offset into cell_local_ids / dest
    uint32_t           n;        // DoFs of the cell
    uint32_t           nproj;    // projectors of the cell (0 without nonlocal part)
    uint32_t           proj_off; // offset into cell_proj_local_ids
  };

  // one work descriptor per processing position (ordered kernel): everything a CTA needs to know about a cell,
  // fetched with a single 32-byte load after the work counter returns
  struct ItemDesc
  {
    unsigned long long h_off;
    uint32_t           ids_off, n, nproj, proj_off, wait_off, nwait;
  };

  struct Comm;  // NCCL communicator wrapper (comm.cu)
  struct Dense; // cuSOLVER handle + workspace (dense.cu)

  // one ConstraintsLocal object on the device: the reference CSR + the parent-side transpose used by the
  // deterministic child->parent pass.  Set 0 lives in the plan's own fields (the mesh's constraints); further sets
  // (e.g. the inhomogeneous-Dirichlet constraints of the Poisson problem's X basis manager) are added with
  // hx_plan_add_constraints.
  struct ConstraintView
  {
    uint32_t        nR = 0, nPar = 0;
    uint32_t        max_row = 0, max_child = 0; // longest constraint row / longest child list of a parent
    const uint32_t *row_ids = nullptr, *row_sizes = nullptr, *row_offsets = nullptr, *col_ids = nullptr;
    const double *  col_vals = nullptr, *inhom = nullptr;
    const uint32_t *par_ids = nullptr, *par_off = nullptr, *par_child = nullptr;
    const double *  par_w = nullptr;
  };
  struct ConstraintSet
  {
    uint32_t         nR = 0, nnz = 0, nPar = 0, max_row = 0, max_child = 0;
    DevBuf<uint32_t> d_row_ids, d_row_sizes, d_row_offsets, d_col_ids, d_par_ids, d_par_off, d_par_child;
    DevBuf<double>   d_col_vals, d_inhom, d_par_w;
    ConstraintView
    view() const
    {
      ConstraintView v;
      v.nR = nR, v.nPar = nPar, v.max_row = max_row, v.max_child = max_child;
      v.row_ids = d_row_ids.p, v.row_sizes = d_row_sizes.p, v.row_offsets = d_row_offsets.p, v.col_ids = d_col_ids.p;
      v.col_vals = d_col_vals.p, v.inhom = d_inhom.p;
      v.par_ids = d_par_ids.p, v.par_off = d_par_off.p, v.par_child = d_par_child.p, v.par_w = d_par_w.p;
      return v;
    }
  };
} // namespace hx

namespace hx
{
  // the packed (fragment-major) cell-matrix stream of the cell kernel: stages of op->kc k-steps (4 columns each) for the
  // CWARPS * mtw m-tiles (8 rows each) of a chunk; see pack_kernel in cell_kernel.cu
  constexpr int CWARPS = 8;
} // namespace hx

#define HX_DEST_STAGED 0x80000000u
#define HX_DEST_FIRST 0x40000000u
#define HX_DEST_LASTF 0x20000000u /* last toucher (processing order) of a row whose Chebyshev update can be fused */
#define HX_DEST_PUSH 0x10000000u  /* last toucher of a ghost row whose partial sum goes straight to its owner (halo overlap) */
#define HX_DEST_ROW(d) ((d)&0x0fffffffu)
#define HX_ITEM_BOUNDARY 0x80000000u /* ItemDesc::nwait: the cell reads ghost rows of X (waits for the halo inside the kernel) */

namespace hx
{
  // Chebyshev recurrence epilogue fused into the cell kernel's scatter: the last toucher of a "fusable" row
  // (owned classical row, unconstrained, no hanging-node children, no halo contribution, not staged) holds the
  // final (H X)[r,:] in registers and writes out[r,:] = a*dinv[r]*(H X)[r,:] + b*X[r,:] + c*xprev[r,:] directly,
  // so H X never goes to memory for those rows.  All other owned rows go through cheb_fused_kernel on a row list.
  struct FuseArgs
  {
    const double *dinv  = nullptr; // M^-1 diagonal over local rows
    const double *xprev = nullptr; // may alias out; ignored when c == 0
    double *      out   = nullptr;
    double        a = 0.0, b = 0.0, c = 0.0;
  };

#ifdef __CUDACC__
  // one definition of the update so the fused epilogue and the row-list kernel agree bit for bit:
  //   z = b*xcur + c*xprev (available before H X is final), out = a*t + z
  // (ChebyshevFilter.t.cpp:105-124 computes (a*t + b*xcur) + c*xprev: same terms, association differs in the last ulp)
  __device__ __forceinline__ double
  cheb_z(double b, double xc, double c, double xp)
  {
    const double z = __dmul_rn(b, xc);
    return (c != 0.0) ? __fma_rn(c, xp, z) : z;
  }
  __device__ __forceinline__ double
  cheb_combine(double a, double t, double b, double xc, double c, double xp)
  {
    return __fma_rn(a, t, cheb_z(b, xc, c, xp));
  }
  // free row of a diagonal M^-1 (t = dinv * hx): the scale s = a*dinv is formed once per row, out = s*hx + z
  __device__ __forceinline__ double
  cheb_combine_diag(double a, double dinv, double hx, double b, double xc, double c, double xp)
  {
    return __fma_rn(__dmul_rn(a, dinv), hx, cheb_z(b, xc, c, xp));
  }
#endif
} // namespace hx

namespace hx
{
  // what the cell kernel needs to take part in a halo exchange (halo overlap, peer-memory transport): all device
  // pointers; x_ready == nullptr switches the whole thing off
  struct HaloK
  {
    uint32_t *       x_ready = nullptr; // local stamp word: epoch once ghosts are unpacked and the accumulate buffers are free
    // update direction (X): wait for the sources' flags, copy my receive buffer into the ghost rows, acknowledge
    const uint32_t * flagU   = nullptr; // [nSrcU] in my arena
    uint32_t *const *rackU   = nullptr; // [nSrcU] ack words at the sources (peer mappings)
    const double *   recvU   = nullptr;
    const uint32_t * unpack_ids = nullptr; // [n_ghost] ghost index of buffer row k; bit 31: constrained, keep as filled
    double *         xghost  = nullptr; // X + n_owned * B
    uint32_t         nSrcU = 0, seqU = 0, n_ghost = 0, do_unpack = 0;
    // accumulate direction (Y): the owners must have consumed the previous message before anyone writes their buffers
    const uint32_t * ackA    = nullptr; // [nDstA] in my arena
    double *const *  push_base = nullptr; // [n_ghost] owner's buffer of ghost row j (peer mapping)
    const uint32_t * push_row  = nullptr; // [n_ghost] row inside it
    uint32_t         nDstA = 0, seqA = 0, n_owned = 0, do_push = 0;
    uint32_t *       counter = nullptr; // CTA counter of the unpack phase
    uint32_t *       status  = nullptr; // raised on timeout
    uint32_t         n_halo_ctas = 0;
  };
} // namespace hx

#ifdef __CUDACC__
namespace hx
{
  // ---- peer-memory halo: device-side primitives shared by peer.cu and the cell kernel ----
  constexpr unsigned long long PEER_TIMEOUT_CYCLES = 57000000000ull; // ~30 s at 1.9 GHz: ranks may arrive late
  __device__ __forceinline__ uint32_t
  ld_acquire_sys(const uint32_t *p)
  {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ void
  st_release_sys(uint32_t *p, uint32_t v)
  {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  }
  // wait until words[i] >= seq for all i < n (sequence numbers only grow); returns false on timeout, and at once
  // when an earlier exchange already timed out (the status word stays raised: no cascade of 30-s waits)
  __device__ __forceinline__ bool
  wait_words(const uint32_t *words, uint32_t n, uint32_t seq, const uint32_t *status)
  {
    if (*reinterpret_cast<const volatile uint32_t *>(status) != 0u)
      return false;
    const unsigned long long t0 = clock64();
    for (uint32_t i = 0; i < n; ++i)
      while ((int32_t)(ld_acquire_sys(words + i) - seq) < 0)
        if (clock64() - t0 > PEER_TIMEOUT_CYCLES)
          return false;
    return true;
  }
} // namespace hx
#endif

struct hx_plan
{
  int          rank = 0, nranks = 1;
  cudaStream_t stream     = nullptr;
  bool         own_stream = false;
  cudaStream_t copy_in = nullptr, copy_out = nullptr; // host-batch pipeline of hx_chebyshev_filter_host_batches
  cudaEvent_t  pipe_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t     n_owned = 0, n_ghost = 0, n_local = 0, n_owned_classical = 0;
  uint64_t     n_global = 0; // sum of n_owned over the ranks (hx_plan_global_size, computed on first use)
  uint32_t     C = 0, S = 0, max_n = 0, max_block = 0;
  size_t       S2 = 0;

  std::vector<uint32_t> h_ncd, h_ids, h_cell_off; // h_cell_off[C+1]
  hx::DevBuf<uint32_t>  d_ids, d_cell_off, d_ncd;

  // constraints (reference CSR) + parent-side transpose for the deterministic child->parent
  uint32_t              nR = 0, nnz = 0;
  hx::DevBuf<uint32_t>  d_row_ids, d_row_sizes, d_row_offsets, d_col_ids;
  hx::DevBuf<double>    d_col_vals, d_inhom;
  uint32_t              nPar = 0, max_row = 0, max_child = 0; // longest constraint row / child list (chain depth of the row kernels)
  std::vector<uint32_t> h_par_ids, h_par_off, h_par_child;
  std::vector<double>   h_par_w;
  hx::DevBuf<uint32_t>  d_par_ids, d_par_off, d_par_child;
  hx::DevBuf<double>    d_par_w;
  std::vector<uint32_t> h_row_ids, h_modrows; // constrained rows; rows an apply may modify in X (+ ghosts)
  hx::DevBuf<uint32_t>  d_modrows;
  hx::DevBuf<uint32_t>  d_rowinfo; // [n_local]: 0xFFFFFFFF free, 0xFFFFFFFE constrained, else parent index

  // colouring of cells over non-shared DoFs; shared (high-incidence, e.g. enrichment) rows go
  // through a staging buffer + ordered reduction
  uint32_t              n_colours = 0;
  std::vector<uint32_t> h_colour, h_colour_off, h_colour_cells;
  hx::DevBuf<uint32_t>  d_colour_cells;
  hx::DevBuf<uint32_t>  d_dest; // [S]: local row id (| HX_DEST_FIRST when the cell is the row's first toucher in
                                // processing order), or HX_DEST_STAGED | staging slot
  uint32_t              n_shared = 0, n_slots = 0;
  hx::DevBuf<uint32_t>  d_sh_rows, d_sh_off, d_sh_slots;
  hx::DevBuf<double>    d_stage; // n_slots x max_block
  // two-stage reduction of heavily shared rows (used when some row has > 2*SH_CHUNK slots)
  uint32_t              n_sh_chunks = 0;
  hx::DevBuf<uint32_t>  d_sh_ch_begin, d_sh_ch_end, d_sh_ch_off;
  hx::DevBuf<double>    d_sh_partial; // n_sh_chunks x max_block

  // ordered (persistent) scatter: cells are processed in `order`; a cell adds into Y only after the
  // immediately preceding toucher of each of its rows has signalled completion -> fixed summation order
  // per row (ascending processing order, the reference's CPU order) without colour launches or atomics
  int                   scatter_mode = 0; // 0 = ordered persistent kernel, 1 = one launch per colour
  std::vector<uint32_t> h_order, h_wait_off, h_wait_list;
  std::vector<char>     h_boundary; // per cell: reads ghost rows of X
  hx::DevBuf<uint32_t>  d_order, d_wait_off, d_wait_list;
  hx::DevBuf<uint32_t>  d_flags;    // [C * ceil(max_block/8)] epoch stamps
  hx::DevBuf<uint32_t>  d_counters; // [0] work counter, [1] finished CTAs
  hx::DevBuf<unsigned long long> d_clk; // [0] SM cycles, [1] ns that CTA 0 of the cell kernel ran (while kernel timing is on)
  uint32_t              epoch = 0;
  uint32_t              n_untouched = 0;
  hx::DevBuf<uint32_t>  d_untouched; // rows no cell writes (zeroed explicitly each apply)
  // Chebyshev epilogue fusion: owned rows NOT updated inside the cell kernel (row-list pass afterwards)
  uint32_t              n_nonfuse = 0, n_fusable = 0;
  hx::DevBuf<uint32_t>  d_nonfuse_rows;
  // the same rows split by kind (experiment HXB200_SPLIT_ROWLIST=1): [0, n_nonfuse_plain) rows without a child list,
  // then the parent rows - so that only the parents pay for the deep-chain kernel variant
  uint32_t              n_nonfuse_plain = 0;
  hx::DevBuf<uint32_t>  d_nonfuse_split;
  bool                  cheb_fill_dead = false;        // no row of the M^-1 step reads a constrained row of its input
                                                       // (no parents, no constrained enrichment row): the fused filter
                                                       // skips the hanging-node fill of its scratch H.X
  // halo exchange overlapped with the cell kernel (peer-memory transport, constraint set 0 on both sides):
  //   X side: the ghost rows are unpacked by the first CTAs of the cell kernel itself while the others contract interior
  //           cells; only the cells that read ghost rows (ItemDesc::nwait & HX_ITEM_BOUNDARY) wait for them.  Needs: no
  //           constraint row with a ghost parent (its fill would have to follow the unpack).
  //   Y side: the last toucher of a ghost row stores its final partial sum straight into the owner's accumulate buffer
  //           (HX_DEST_PUSH); rows that get contributions after the kernel (constrained / parent / staged / untouched ghost
  //           rows) are pushed by the short kernel that also raises the flags.
  bool                  overlap_x_ok = false, overlap_y_ok = false;
  uint32_t              n_boundary_cells = 0, n_push_direct = 0;
  hx::DevBuf<uint32_t>  d_unpack_ids;   // [n_ghost] ghost_local_ids, bit 31 set for constrained ghost rows (kept as filled)
  hx::DevBuf<uint32_t>  d_push_rest;    // ghost rows (buffer positions) NOT pushed by the cell kernel
  uint32_t              n_push_rest = 0;
  hx::DevBuf<uint32_t>  d_x_ready;      // [1] epoch stamp: ghosts of X are in place, accumulate buffers may be written
  bool                  cheb_fusable_multirank = true; // no constrained ghost row has parents (see api.cu)
  bool                  cheb_fusable_agreed    = false; // ... on every rank (AND-ed across the communicator once)
  int                   sm_count = 0;

  std::vector<hx::ConstraintSet *> extra_constraints; // sets 1.. (set 0 = the fields above)
  hx::ConstraintView
  constraint_view(uint32_t set) const;

  hx::Halo  halo;
  hx::Comm *comm = nullptr;
  int       halo_transport = 0;          // 0 = undecided / single rank, 1 = NCCL send/recv, 2 = NVLink peer memory
  std::vector<hx::Halo *> peer_halos;    // halos with a live peer transport (status checked at synchronisation)

  // scratch block vectors (n_local x max_block), allocated on demand
  std::vector<hx::DevBuf<double> *> scratch;
  hx::DevBuf<double>                d_small; // small device scratch (norms, gram blocks, per-column scalars)
  hx::DevBuf<double>                d_dense_s, d_dense_q, d_dense_w; // B x B projected matrix, rotation matrix, eigenvalues
  hx::Dense *                       dense = nullptr;               // cuSOLVER handle + workspace (dense.cu)
  double *                          h_pinned = nullptr;
  size_t                            h_pinned_bytes = 0;

  // phase trace (HXB200_TRACE=1 or hx_plan_trace): CUDA events at the phase boundaries of an apply / filter degree
  bool                                     trace = false;
  std::vector<std::pair<const char *, cudaEvent_t>> trace_marks;
  std::vector<cudaEvent_t>                 trace_pool;
  void
  mark(const char *name); // no-op unless tracing

  uint64_t    launches = 0;
  bool        timing   = false;
  double      cell_ms  = 0.0;
  uint64_t    cell_launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_pool; // (start, stop) pairs recorded around the cell-kernel launches
  size_t                   ev_used = 0;

  ~hx_plan();
  int
  get_scratch(size_t idx, double **p, uint32_t cols = 0); // n_local * cols doubles (0: max_block), grow-only
  int
  ensure_small(size_t doubles);
  int
  ensure_pinned(size_t bytes);
};

enum hx_op_kind
{
  HX_OP_CELL = 1,
  HX_OP_DIAG = 2
};

struct hx_op
{
  hx_plan *plan = nullptr;
  int      kind = 0;
  // --- cell operator ---
  std::vector<hx::CellMeta> h_meta;
  hx::DevBuf<hx::CellMeta>  d_meta;
  hx::DevBuf<hx::ItemDesc>  d_items; // [C] in processing order
  hx::DevBuf<double>        d_packed; // fragment-major packed cell matrices (+ projector columns)
  size_t                    packed_doubles = 0;
  bool                      have_matrices  = false;
  uint32_t                  max_kp = 0, max_mp = 0;
  int                       mtw = 2; // m-tiles (8 rows) per compute warp per chunk: packed layout depends on it
  int                       kc  = 4; // k-steps (4 columns) per pipeline stage: packed layout depends on it
  // nonlocal
  bool                      has_nl = false;
  int                       nl_reads_ghost = -1; // a projector cell reads ghost rows of X (-1: not determined yet)
  hx::Halo                  phalo;
  uint32_t                  n_proj_local = 0, sum_proj = 0;
  std::vector<uint32_t>     h_ncp, h_pids, h_nl_cells; // cells with projectors
  hx::DevBuf<uint32_t>      d_pids, d_nl_cells;
  hx::DevBuf<double>        d_cell_c, d_v, d_cx, d_cx_stage;
  std::vector<size_t>       h_c_off;                   // per cell offset into cell_c
  hx::DevBuf<unsigned long long> d_c_off;
  hx::DevBuf<uint32_t>      d_pr_off, d_pr_slots;      // projector row -> staging slots (ordered)
  uint32_t                  x_set = 0, y_set = 0; // constraint sets applied to X (parent->child) and Y (child->parent)
  bool                      share_identical = false; // hx_cellop_set_matrix_sharing
  uint32_t                  n_unique = 0;            // distinct packed matrices after the last set_matrices
  // --- diagonal operator ---
  int                variant = 0;
  hx::DevBuf<double> d_diag, d_enr_block;
  uint32_t           nE = 0;
  uint32_t           nE_global = 0, enr_offset = 0; // HX_DIAG_OEFE_GLOBAL: all enrichment functions of the system, first owned one
};

namespace hx
{
  // kernels.cu / cell_kernel.cu entry points (host launchers)
  int launch_p2c(hx_plan *p, double *X, uint32_t B, uint32_t set = 0);
  // zero_rows = false leaves the constrained rows of Y as they are (for a scratch Y whose constrained rows nobody reads)
  int launch_c2p(hx_plan *p, double *Y, uint32_t B, uint32_t set = 0, bool zero_rows = true);
  int launch_coldot(hx_plan *p, const double *x, const double *y, uint32_t B, size_t nrows, double *out_dev);
  int launch_col_divide(hx_plan *p, const double *num, const double *den, double *out, double *out_neg, uint32_t B);
  int launch_cg_dots2(hx_plan *p, const double *z, const double *r, const double *pd, const double *w, uint32_t B,
                      double *zdotr, double *pdotw, double *alpha, double *nalpha);
  int launch_cg_update(hx_plan *p, double *x, const double *pd, double *r, const double *w, double *z, const double *dinv,
                       const double *alpha, uint32_t B, double *zdotr_new, double *rr, const double *zdotr_old, double *beta);
  int launch_zero_constrained(hx_plan *p, double *Y, uint32_t B, uint32_t set = 0);
  int launch_pack(hx_plan *p, const double *x, uint32_t B, const uint32_t *ids, uint32_t n, double *buf);
  int launch_unpack(hx_plan *p, const double *buf, uint32_t B, const uint32_t *ids, uint32_t n, double *x);
  int launch_add_rows(hx_plan *p, const double *buf, uint32_t B, const uint32_t *rows, const uint32_t *off,
                      const uint32_t *pos, uint32_t nrows, double *x);
  int launch_row_scale(hx_plan *p, const double *d, const double *x, double *y, uint32_t B, size_t nrows);
  int launch_axpby(hx_plan *p, size_t n, double a, const double *x, double b, const double *y, double *z);
  int launch_axpby_blocked(hx_plan *p, size_t nrows, uint32_t B, double a1, const double *a, const double *x,
                           double b1, const double *b, const double *y, double *z);
  int launch_colsumsq(hx_plan *p, const double *x, uint32_t B, size_t nrows, double *out_dev);
  int launch_shared_reduce(hx_plan *p, double *Y, uint32_t B);
  int launch_zero_rows(hx_plan *p, double *Y, uint32_t B, const uint32_t *rows, uint32_t n);
  int launch_enr_block(hx_plan *p, const double *blk, uint32_t nE, const double *Xenr, double *Yenr, uint32_t B);
  int launch_enr_block_global(hx_plan *p, const double *blk, uint32_t nEg, uint32_t off, uint32_t nE, const double *Xg, double *Yenr,
                              uint32_t B);
  // fuse != nullptr asks for the Chebyshev epilogue; *fused_applied tells whether the launched variant did it
  int launch_cell_apply(hx_op *op, const double *X, double *Y, uint32_t B, const FuseArgs *fuse = nullptr,
                        bool *fused_applied = nullptr, const HaloK *halo = nullptr);
  int launch_nl_phase_a(hx_op *op, const double *X, uint32_t B);
  int pack_cell_matrices(hx_op *op, const double *raw_dev_or_host, int on_device); // raw == nullptr: structure only
  // use_row_list false: all owned rows; true: only the n_rows listed rows
  // force_depth: 0 = by the plan's longest child list and the launch size, 1 = the rows are known to have no child list
  int launch_cheb_fused(hx_plan *p, hx_op *binv, const double *s1, const double *xcur, const double *xprev,
                        double *out, uint32_t B, double a, double b, double c, bool use_row_list = false,
                        const uint32_t *rows = nullptr, uint32_t n_rows = 0, int force_depth = 0);
  int gram_block(hx_plan *p, const double *X, uint32_t B, uint32_t j0, const double *OpXb, uint32_t b,
                 size_t nOwned, double *S_dev);
  int rotate(hx_plan *p, double *X, uint32_t B, size_t nOwned, const double *Q_dev, int transpose, int lowerTri,
             double *tmp, size_t tmp_rows = 0); // tmp holds tmp_rows x B doubles (0: nOwned rows)
  // api.cu helpers shared with eigen.cu
  int op_apply(hx_op *op, double *X, double *Y, uint32_t B, int ugx, int ugy);
  int copy_cols(hx_plan *p, const double *src, uint32_t ldsrc, uint32_t c0s, double *dst, uint32_t lddst, uint32_t c0d,
                uint32_t ncols, size_t nrows);
  int halo_update(hx_plan *p, Halo &h, double *X, uint32_t B);
  int plan_sync(hx_plan *p); // stream synchronisation + status of the peer-memory halos
  size_t gram_workspace_doubles(const hx_plan *p, uint32_t B, uint32_t batch, size_t nOwned); // all column batches of a B-wide block
  // dense.cu: B x B subspace problems on the device (cuSOLVER, resolved with dlopen)
  struct Dense;
  void dense_destroy(Dense *d);
  int  dense_cholesky_inverse(hx_plan *p, double *S_dev, uint32_t B, int *info_host);
  int  dense_sym_eig(hx_plan *p, double *S_dev, uint32_t B, double *evals_dev, int *info_host);
  int  dense_place_gram_block(hx_plan *p, const double *Sd, uint32_t M, uint32_t b, uint32_t j0, double *S, uint32_t B);
  int  dense_transpose(hx_plan *p, const double *A, double *At, uint32_t B);
  // comm.cu
  int comm_unique_id(char id[128]);
  int comm_create(Comm **c, const char id[128], int nranks, int rank);
  void comm_destroy(Comm *c);
  int comm_exchange(Comm *c, cudaStream_t s, const double *send, const std::vector<uint32_t> &send_procs,
                    const std::vector<size_t> &send_counts, double *recv, const std::vector<uint32_t> &recv_procs,
                    const std::vector<size_t> &recv_counts);
  int comm_allreduce_sum(Comm *c, cudaStream_t s, double *buf, size_t n);
  int comm_allgather_bytes(Comm *c, cudaStream_t s, const void *mine, void *all, size_t bytes_per_rank);
  // peer.cu
  int peer_setup(hx_plan *p, Halo &h);
  int peer_halo_update(hx_plan *p, Halo &h, double *X, uint32_t B);
  int peer_halo_accumulate(hx_plan *p, Halo &h, double *Y, uint32_t B);
  bool peer_acc_update_is_small(const Halo &h, uint32_t B);
  int peer_halo_accumulate_update_small(hx_plan *p, Halo &h, double *X, uint32_t B);
  int peer_check_status(Halo &h);
  // halo overlap: push only (the cell kernel unpacks), kernel arguments, closing kernels of the accumulate
  int peer_push_update(hx_plan *p, Halo &h, const double *X, uint32_t B, uint32_t *seq);
  int peer_overlap_args(hx_plan *p, Halo &h, double *X, uint32_t B, bool do_unpack, uint32_t seqU, HaloK *k);
  int peer_finish_accumulate(hx_plan *p, Halo &h, double *Y, uint32_t B, uint32_t seqA);
  bool peer_overlap_available(const Halo &h);
} // namespace hx
