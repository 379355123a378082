// cell_kernel.cu — the dominant kernel: fused cell gather -> FP64 DMMA cell contraction -> coloured scatter.
//
// Replaces, per reference H.X apply (src/ksdft/KohnShamOperatorContextFE.t.cpp:951-1199):
//   FECellWiseDataOperations::copyFieldToCellWiseData   (src/basis/FECellWiseDataOperations.t.cpp:58-86)
//   blasLapack::gemmStridedVarBatched (one dgemm_/cell)  (src/linearAlgebra/BlasLapack.t.cpp:388-436)
//   AtomCenterNonLocalOpContextFE::applyCOnVCconjtransX  (src/basis/AtomCenterNonLocalOpContextFE.t.cpp:998-1036)
//   FECellWiseDataOperations::addCellWiseDataToFieldData (src/basis/FECellWiseDataOperations.t.cpp:87-153)
//
// Layout / algorithm (B200, sm_100a):
//   * tcgen05 has no FP64 kind; the FP64 tensor path on sm_100a is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
//   * One CTA = one cell x one tile of BT = 8*NT wavefunction columns.  y_c[j,v] = sum_k A_c[j,k] xk[k,v] with
//     A_c = [H_c | C_c^T] (the nonlocal C.(V C^H X) term is a K-extension: rows n..n+nProj of the B operand are
//     the V-scaled projector coefficients), so projector cells cost no extra pass.
//   * A_c is pre-tiled once per reinit into DMMA-fragment-major order (pack_kernel): the 32 doubles of one 8x4
//     A fragment are contiguous, so every warp streams its row panel from HBM with perfectly coalesced 256-B
//     loads straight into registers (register prefetch ring, depth PD), no shared-memory staging of A:
//     each A element is used by exactly one warp of one CTA.
//   * The gathered x_c tile lives in shared memory (row stride BT+4 doubles -> conflict-free B-fragment reads).
//   * Accumulators stay in registers and are scattered with 16-B read-modify-writes; cells of one launch have
//     the same colour (share no DoF), so there are no atomics and the result is bitwise reproducible.  Rows
//     shared by many cells (enrichment DoFs) are written to a staging slot and reduced in fixed order afterwards.
#include "hx_internal.h"

namespace hx
{
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
    const uint32_t *cell_list;
    uint32_t        B;
    uint32_t        nBt;
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

  constexpr int CELL_THREADS = 256;
  constexpr int CELL_WARPS   = CELL_THREADS / 32;
  constexpr int MTW          = 2; // m-tiles (8 rows each) per warp per chunk
  constexpr int PD           = 4; // register prefetch depth (k-steps) of the A stream

  template <int NT, bool VEC, int MINB>
  __global__ void __launch_bounds__(CELL_THREADS, MINB) cell_apply_kernel(const CellArgs a)
  {
    extern __shared__ __align__(16) double xs[];
    constexpr int BT  = NT * 8;
    constexpr int LDX = BT + 4;
    const int     tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const uint32_t bt   = blockIdx.x % a.nBt;
    const uint32_t ci   = blockIdx.x / a.nBt;
    const uint32_t cell = a.cell_list[ci];
    const CellMeta cm   = a.meta[cell];
    const int      n    = (int)cm.n;
    const int      ktot = n + (int)cm.nproj;
    const int      Kp   = (ktot + 3) & ~3;
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

    // ---- contraction: each warp owns MTW m-tiles per chunk of CELL_WARPS*MTW tiles ----
    const double *Abase = a.packed + cm.h_off + lane;
    const double *xrow  = xs + (size_t)(lane & 3) * LDX + (lane >> 2);
    for (int mc = 0; mc < nMt; mc += CELL_WARPS * MTW)
      {
        const int mt0 = mc + warp * MTW;
        if (mt0 >= nMt)
          break;
        const double *Ap[MTW];
#pragma unroll
        for (int j = 0; j < MTW; ++j)
          {
            const int mt = min(mt0 + j, nMt - 1);
            Ap[j]        = Abase + (size_t)mt * nK * 32;
          }
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
            af[i][j] = (i < nK) ? ld_stream(Ap[j] + (size_t)i * 32) : 0.0;

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
                          af[i][j] = ld_stream(Ap[j] + (size_t)kn * 32);
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

        // ---- scatter-add (colour-exclusive rows: plain RMW; shared rows: staging slot) ----
#pragma unroll
        for (int j = 0; j < MTW; ++j)
          {
            const int r = (mt0 + j) * 8 + (lane >> 2);
            if (mt0 + j < nMt && r < n)
              {
                const uint32_t d   = __ldg(a.dest + cm.ids_off + r);
                double *       dst = (d & 0x80000000u) ? a.stage + (size_t)(d & 0x7fffffffu) * B : a.Y + (size_t)d * B;
                const bool     add = !(d & 0x80000000u);
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
  // pack: raw row-major n x n cell matrices (+ column-major nProj x n projector matrices) -> fragment-major
  // tiles.  packed[((mt*nK + ks)*32 + lane)] = A[mt*8 + lane/4][ks*4 + lane%4].
  __global__ void
  pack_kernel(const double *            raw,
              unsigned long long        raw_base,
              const unsigned long long *raw_off,
              const double *            cellC,
              const unsigned long long *c_off,
              const CellMeta *          meta,
              double *                  packed,
              uint32_t                  cell_begin)
  {
    const uint32_t cell = cell_begin + blockIdx.x;
    const CellMeta cm   = meta[cell];
    const int      n = (int)cm.n, np = (int)cm.nproj;
    const int      Kp = (n + np + 3) & ~3, nK = Kp >> 2, Mp = (n + 7) & ~7;
    const double * H  = raw + (raw_off[cell] - raw_base);
    const double * Cc = (np > 0) ? cellC + c_off[cell] : nullptr;
    double *       out = packed + cm.h_off;
    const size_t   tot = (size_t)Mp * Kp;
    for (size_t idx = threadIdx.x; idx < tot; idx += blockDim.x)
      {
        const int lane = (int)(idx & 31);
        const size_t f = idx >> 5;
        const int ks = (int)(f % nK), mt = (int)(f / nK);
        const int r = mt * 8 + (lane >> 2), k = ks * 4 + (lane & 3);
        double    v = 0.0;
        if (r < n)
          {
            if (k < n)
              v = H[(size_t)r * n + k];
            else if (k < n + np)
              v = Cc[(size_t)(k - n) + (size_t)r * np];
          }
        out[idx] = v;
      }
  }

  int
  pack_cell_matrices(hx_op *op, const double *raw, int on_device)
  {
    hx_plan *p = op->plan;
    // packed offsets
    size_t tot = 0;
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
        const uint32_t Kp = (m.n + m.nproj + 3) & ~3u, Mp = (m.n + 7) & ~7u;
        m.h_off = tot;
        tot += (size_t)Kp * Mp;
        op->max_kp = Kp > op->max_kp ? Kp : op->max_kp;
        op->max_mp = Mp > op->max_mp ? Mp : op->max_mp;
      }
    HX_TRY(op->d_meta.upload(op->h_meta));
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

    if (on_device)
      {
        if (p->C)
          {
            pack_kernel<<<p->C, 256, 0, p->stream>>>(raw, 0ull, d_raw_off.p, op->d_cell_c.p, op->d_c_off.p,
                                                     op->d_meta.p, op->d_packed.p, 0);
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
                                                        op->d_c_off.p, op->d_meta.p, op->d_packed.p, c0);
            p->launches++;
            HX_CUDA(cudaGetLastError());
            HX_CUDA(cudaStreamSynchronize(p->stream));
            c0 = c1;
          }
      }
    op->have_matrices = true;
    return HX_OK;
  }

  template <int NT, bool VEC, int MINB>
  static int
  launch_colours(hx_op *op, const CellArgs &base, size_t smem)
  {
    hx_plan *p = op->plan;
    auto     k = cell_apply_kernel<NT, VEC, MINB>;
    HX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (p->timing)
      {
        // events are only recorded here (no host sync inside the timed region); they are read back in
        // hx_plan_cell_kernel_time_ms
        if (p->ev_used + 2 > p->ev_pool.size())
          for (int i = 0; i < 64; ++i)
            {
              cudaEvent_t e;
              HX_CUDA(cudaEventCreate(&e));
              p->ev_pool.push_back(e);
            }
        e0 = p->ev_pool[p->ev_used++];
        e1 = p->ev_pool[p->ev_used++];
        HX_CUDA(cudaEventRecord(e0, p->stream));
      }
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
    if (p->timing)
      HX_CUDA(cudaEventRecord(e1, p->stream));
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  int
  launch_cell_apply(hx_op *op, const double *X, double *Y, uint32_t B)
  {
    hx_plan *p = op->plan;
    HX_CHECK(op->have_matrices, HX_ERR_INVALID, "cell operator has no matrices (call hx_cellop_set_matrices)");
    if (p->C == 0)
      return HX_OK;
    CellArgs a;
    a.X      = X;
    a.Y      = Y;
    a.VCX    = op->d_cx.p;
    a.stage  = p->d_stage.p;
    a.packed = op->d_packed.p;
    a.meta   = op->d_meta.p;
    a.ids    = p->d_ids.p;
    a.dest   = p->d_dest.p;
    a.pids   = op->d_pids.p;
    a.B      = B;
    // tile width: widest of {8,16,32,64} columns that B needs and shared memory allows
    int nt = B > 32 ? 8 : (B > 16 ? 4 : (B > 8 ? 2 : 1));
    auto smem_of = [&](int nt_) { return (size_t)op->max_kp * (nt_ * 8 + 4) * sizeof(double); };
    while (nt > 1 && smem_of(nt) > 200 * 1024)
      nt >>= 1;
    HX_CHECK(smem_of(nt) <= 220 * 1024, HX_ERR_UNSUPPORTED, "cell with %u DoFs does not fit shared memory", op->max_kp);
    a.nBt          = (B + nt * 8 - 1) / (nt * 8);
    const bool vec = (B % 2 == 0);
    const size_t smem = smem_of(nt);
#define HX_DISPATCH(NT_, MINB_)                                  \
  (vec ? launch_colours<NT_, true, MINB_>(op, a, smem) : launch_colours<NT_, false, MINB_>(op, a, smem))
    switch (nt)
      {
        case 8:
          return HX_DISPATCH(8, 1);
        case 4:
          return HX_DISPATCH(4, 2);
        case 2:
          return HX_DISPATCH(2, 2);
        default:
          return HX_DISPATCH(1, 2);
      }
#undef HX_DISPATCH
  }

  // -------------------------------------------------------------------------------------------------
  // Nonlocal phase A: per projector cell, CXcell[p,v] = sum_k C_c[p + k*nP] x_c[k,v]
  // (AtomCenterNonLocalOpContextFE::applyCconjtransOnX, src/basis/AtomCenterNonLocalOpContextFE.t.cpp:889-942),
  // written to a per-(cell,projector) staging row; reduced per projector row in ascending cell order afterwards.
  __global__ void __launch_bounds__(256)
  nl_phase_a_kernel(const double *X, const uint32_t *ids, const uint32_t *nl_cells, const CellMeta *meta,
                    const double *cellC, const unsigned long long *c_off, double *cx_stage, uint32_t B)
  {
    extern __shared__ __align__(16) double xs[]; // [n][32]
    const uint32_t cell = nl_cells[blockIdx.x];
    const uint32_t b0   = blockIdx.y * 32;
    const CellMeta cm   = meta[cell];
    const int      n = (int)cm.n, np = (int)cm.nproj;
    const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t col  = b0 + lane;
    for (int k = warp; k < n; k += 8)
      xs[k * 32 + lane] = (col < B) ? X[(size_t)ids[cm.ids_off + k] * B + col] : 0.0;
    __syncthreads();
    const double *Cc = cellC + c_off[cell];
    for (int pj = warp; pj < np; pj += 8)
      {
        double s = 0.0;
        for (int k = 0; k < n; ++k)
          s += Cc[(size_t)pj + (size_t)k * np] * xs[k * 32 + lane];
        if (col < B)
          cx_stage[(size_t)(cm.proj_off + pj) * B + col] = s;
      }
  }

  // CX[row,:] = V[row] * sum over the row's staging slots (fixed order)   [reduce + (single rank) V scale]
  __global__ void
  nl_reduce_kernel(const double *cx_stage, const uint32_t *pr_off, const uint32_t *pr_slots, const double *V,
                   double *CX, uint32_t n_rows, uint32_t B, int scale)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_rows * B)
      return;
    const uint32_t row = (uint32_t)(i / B), v = (uint32_t)(i % B);
    double         s = 0.0;
    for (uint32_t e = pr_off[row]; e < pr_off[row + 1]; ++e)
      s += cx_stage[(size_t)pr_slots[e] * B + v];
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
        const size_t smem = (size_t)p->max_n * 32 * sizeof(double);
        HX_CUDA(cudaFuncSetAttribute(nl_phase_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(ncell, (B + 31) / 32);
        nl_phase_a_kernel<<<grid, 256, smem, p->stream>>>(X, p->d_ids.p, op->d_nl_cells.p, op->d_meta.p,
                                                          op->d_cell_c.p, op->d_c_off.p, op->d_cx_stage.p, B);
        p->launches++;
      }
    const bool   single = (p->nranks == 1);
    const size_t tot    = (size_t)op->n_proj_local * B;
    if (tot)
      {
        nl_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, p->stream>>>(
          op->d_cx_stage.p, op->d_pr_off.p, op->d_pr_slots.p, op->d_v.p, op->d_cx.p, op->n_proj_local, B, single ? 1 : 0);
        p->launches++;
      }
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }
} // namespace hx
