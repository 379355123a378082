// assemble.cu — the step BEFORE the H.X path each SCF iteration: the cell matrices of a local
// potential, FEBasisOperations::computeFEMatrices(IDENTITY, MULT, MULT, IDENTITY, f)
// (src/basis/FEBasisOperations.t.cpp:2210-2243 -> BasisWeakFormKernelWithField, :41-427):
//
//     C_c[i, j] = sum_q N_c[q, i] * f[q] * JxW[q] * N_c[q, j]          per cell c, n_c x n_c, k = nq_c
//
// The reference forms f x JxW (hadamardProduct), scales a copy of the basis values with it (scaleStridedVarBatched)
// and calls one dgemm('N','C') per cell.  Here one CTA owns one 64 x 64 tile of one cell's matrix (tiles on and below
// the diagonal only; the mirror image is written by the same CTA), streams the basis values of the cell in chunks of
// 16 quadrature points through a cp.async double buffer, applies f x JxW to the A fragments in registers and feeds
// mma.sync.m8n8k4.f64.  An optional `add_to` array (e.g. the kinetic cell matrices) is added in the epilogue, which
// is KohnShamOperatorContextFE::reinit's component sum (src/ksdft/KohnShamOperatorContextFE.t.cpp:1259-1282) fused.
// Output: the flat S2 array hx_cellop_set_matrices(.., on_device = 1) consumes - the potential never leaves the GPU.
#include <algorithm>
#include <memory>

#include "hx_internal.h"

struct hx_fe_basis
{
  hx_plan *                   plan = nullptr;
  bool                        same_basis = false;
  uint32_t                    C = 0, max_n = 0;
  size_t                      n_quad_total = 0, S2 = 0;
  std::vector<uint32_t>       h_nq;
  hx::DevBuf<double>          d_basis, d_jxw, d_w, d_f, d_occ, d_rho;
  struct Cell
  {
    unsigned long long basis_off, out_off;
    uint32_t           q_off, n, nq, ids_off;
  };
  hx::DevBuf<Cell> d_cells;
};

namespace hx
{
  constexpr int FT  = 64; // tile edge
  constexpr int FKC = 16; // quadrature points per stage
  constexpr int FLD = FT + 4;

  __device__ __forceinline__ void
  fe_dmma(double &d0, double &d1, const double a, const double b)
  {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
  }
  __device__ __forceinline__ void
  fe_cp16(void *smem, const void *gmem)
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
  }
  __device__ __forceinline__ void
  fe_cp8(void *smem, const void *gmem)
  {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
  }

  // [FKC x 64] tile of the cell's basis values (row q, columns c0..c0+63 of an nq x n matrix, DoF index fastest)
  __device__ __forceinline__ void
  fe_load_tile(double *sm, const double *g, uint32_t n, uint32_t q0, uint32_t nq, uint32_t c0, bool aligned16, int tid)
  {
#pragma unroll
    for (int it = 0; it < 2; ++it)
      {
        const int      ch = tid + it * 256;
        const int      r  = ch >> 5;
        const int      cc = (ch & 31) * 2;
        double *       d  = sm + r * FLD + cc;
        const uint32_t q = q0 + r, c = c0 + cc;
        const double * s = g + (size_t)q * n + c;
        if (q < nq && c + 1 < n)
          {
            if (aligned16)
              fe_cp16(d, s);
            else
              {
                fe_cp8(d, s);
                fe_cp8(d + 1, s + 1);
              }
          }
        else
          { // last odd column / rows beyond the rule: still asynchronous (a synchronous load here would stall the
            // whole stage on the global latency)
            if (q < nq && c < n)
              fe_cp8(d, s);
            else
              d[0] = 0.0;
            d[1] = 0.0;
          }
      }
  }

  // grid: cells x lower tiles of the largest cell, the tiles of one cell adjacent (they read the same basis values:
  // launched together they share them through L2)
  // PACKED: the result goes straight into the cell kernel's fragment-major stream (pack_kernel's layout): an 8 x 4
  // block of the matrix is one 256-B fragment, and the two accumulator values a thread holds for (row, col..col+1)
  // are neighbours inside it
  template <bool PACKED>
  __global__ void __launch_bounds__(256, 4)
  fe_matrices_kernel(const hx_fe_basis::Cell *cells, const double *basis, const double *w, const double *add_to, double *out,
                     uint32_t tilesPerCell, const CellMeta *meta, int mpc, int KCv)
  {
    __shared__ __align__(16) double As[2][FKC * FLD];
    __shared__ __align__(16) double Bs[2][FKC * FLD];
    __shared__ double               Ws[2][FKC];
    const hx_fe_basis::Cell         cell = cells[blockIdx.x / tilesPerCell];
    const uint32_t                  n = cell.n, nq = cell.nq;
    const uint32_t                  tilesM = (n + FT - 1) / FT;
    // tile index -> (tm, tn) with tn <= tm
    uint32_t tm = 0, t = blockIdx.x % tilesPerCell;
    while (t > tm)
      {
        t -= tm + 1;
        ++tm;
      }
    const uint32_t tn = t;
    if (tm >= tilesM)
      return;
    const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t m0 = tm * FT, n0 = tn * FT;
    const double * N  = basis + cell.basis_off;
    const double * wc = w + cell.q_off;
    const bool     aligned16 = ((n & 1u) == 0) && ((cell.basis_off & 1ull) == 0);
    double         acc[2][4][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u)
        acc[j][u][0] = acc[j][u][1] = 0.0;
    const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
    const int nchunks = (int)((nq + FKC - 1) / FKC);
    const bool diag    = (tm == tn); // both operands are the same columns of N: load them once
    auto       load    = [&](int buf, uint32_t q0) {
      fe_load_tile(As[buf], N, n, q0, nq, m0, aligned16, tid);
      if (!diag)
        fe_load_tile(Bs[buf], N, n, q0, nq, n0, aligned16, tid);
      if (tid < FKC)
        Ws[buf][tid] = (q0 + tid < nq) ? wc[q0 + tid] : 0.0;
    };
    // the summand of the epilogue (reinit's component sum): pull its lines into L2 now, behind the k loop
    const double *add = add_to ? add_to + cell.out_off : nullptr;
    if (add)
      {
        const uint32_t r = tid >> 2, seg = (tid & 3) * 16;
        if (m0 + r < n && n0 + seg < n)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(add + (size_t)(m0 + r) * n + n0 + seg));
        if (!diag && n0 + r < n && m0 + seg < n)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(add + (size_t)(n0 + r) * n + m0 + seg));
      }
    if (nchunks)
      {
        load(0, 0);
        asm volatile("cp.async.commit_group;");
      }
    for (int c = 0; c < nchunks; ++c)
      {
        const int cur = c & 1;
        if (c + 1 < nchunks)
          {
            load(cur ^ 1, (uint32_t)(c + 1) * FKC);
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 1;");
          }
        else
          asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        const double *as = As[cur] + (lane & 3) * FLD + wm + (lane >> 2);
        const double *bs = (diag ? As[cur] : Bs[cur]) + (lane & 3) * FLD + wn + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < FKC / 4; ++k4)
          {
            const double wk = Ws[cur][k4 * 4 + (lane & 3)];
            double       a[2], bb[4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
              a[j] = as[k4 * 4 * FLD + j * 8] * wk; // fxJxWxN of the reference, formed in registers
#pragma unroll
            for (int u = 0; u < 4; ++u)
              bb[u] = bs[k4 * 4 * FLD + u * 8];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int u = 0; u < 4; ++u)
                fe_dmma(acc[j][u][0], acc[j][u][1], a[j], bb[u]);
          }
        __syncthreads();
      }
    double *o = out + cell.out_off;
    // packed layout of this cell (PACKED only)
    size_t per_chunk = 0;
    int    nKC = 0, nMt = 0;
    if (PACKED)
      {
        const CellMeta cm = meta[blockIdx.x / tilesPerCell];
        o                 = out + cm.h_off;
        nKC               = ((int)cm.n + (int)cm.nproj + 4 * KCv - 1) / (4 * KCv);
        nMt               = ((int)cm.n + 7) >> 3;
        per_chunk         = (size_t)mpc * nKC * KCv * 32;
      }
    auto pidx = [&](uint32_t r, uint32_t c) -> size_t {
      const int mt = (int)(r >> 3), ch = mt / mpc, mc = ch * mpc, mtc = min(mpc, nMt - mc), mtl = mt - mc;
      const int kc = (int)(c >> 2) / KCv, ks = (int)(c >> 2) % KCv;
      return (size_t)ch * per_chunk + (size_t)kc * mtc * KCv * 32 + (size_t)(mtl * KCv + ks) * 32 + (r & 7) * 4 + (c & 3);
    };
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          {
            const uint32_t r = m0 + wm + j * 8 + (lane >> 2);
            const uint32_t cidx = n0 + wn + u * 8 + (lane & 3) * 2 + e;
            if (r < n && cidx < n)
              {
                const double v  = acc[j][u][e];
                const size_t i1 = (size_t)r * n + cidx;
                o[PACKED ? pidx(r, cidx) : i1] = add ? v + add[i1] : v;
                if (!diag)
                  { // mirror image of an off-diagonal tile (the matrix is symmetric)
                    const size_t i2 = (size_t)cidx * n + r;
                    o[PACKED ? pidx(cidx, r) : i2] = add ? v + add[i2] : v;
                  }
              }
          }
  }

  // ---------------------------------------------------------------------------------------------------------------
  // 8f rank 3: DensityCalculator::computeRho (src/ksdft/DensityCalculator.t.cpp:283-437) =
  //   FEBasisOperations::interpolate (src/basis/FEBasisOperations.t.cpp:996-1275: per cell psiQuad (B x nq) =
  //   xCell (B x n_c) . N (n_c x nq)) followed by computeRhoInBatch (:37-70: rho[q] = sum_i 2 |psi_i(q)|^2 occ_i).
  // One CTA = up to 256 quadrature points of one cell: for every 32-wide pass over the wavefunctions it gathers the cell's rows of X
  // (cell -> DoF map) in chunks of 16 DoFs next to the matching basis values, runs the DMMA contraction, squares the
  // accumulators, weights them with 2 occ_i and keeps one partial sum per quadrature point; psi at the quadrature
  // points never goes to memory.  Fixed summation order (no atomics).
  constexpr int RLD = FKC + 4;
  constexpr int RQ  = 256;    // quadrature points per CTA (8 warps x 32 rows)
  constexpr int RV  = 32;     // wavefunctions per pass
  constexpr int RBL = RV + 4; // padded row of the gathered X tile
  constexpr size_t RHO_SMEM = (size_t)2 * (RQ * RLD + FKC * RBL) * sizeof(double);
  __global__ void __launch_bounds__(256, 2)
  rho_kernel(const hx_fe_basis::Cell *cells, const double *basis, const uint32_t *ids, const double *X, uint32_t B,
             const double *occ2, double *rho)
  {
    extern __shared__ __align__(16) double rsm[];
    double *                               As0 = rsm;                 // [2][RQ * RLD]  basis values [q][dof chunk]
    double *                               Bs0 = rsm + 2 * RQ * RLD;  // [2][FKC * RBL] gathered X rows [dof chunk][vectors]
    const hx_fe_basis::Cell                cell = cells[blockIdx.x];
    const uint32_t                         n = cell.n, nq = cell.nq;
    const uint32_t                         q0 = blockIdx.y * RQ;
    if (q0 >= nq)
      return;
    const int       tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *  N   = basis + cell.basis_off;
    const uint32_t *cid = ids + cell.ids_off;
    const bool      alignedX = ((B & 1u) == 0) && ((((uintptr_t)X) & 15) == 0);
    const bool      alignedN = ((n & 1u) == 0) && ((cell.basis_off & 1ull) == 0);
    const int       wm       = warp * 32;
    // a warp whose 32 rows lie beyond the cell's quadrature points still helps loading, but skips the math
    const bool      active = q0 + wm < nq;
    double          part[4] = {0.0, 0.0, 0.0, 0.0}; // rows wm + j*8 + (lane >> 2)
    const int       nchunks = (int)((n + FKC - 1) / FKC);
    for (uint32_t v0 = 0; v0 < B; v0 += RV)
      {
        double acc[4][4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int u = 0; u < 4; ++u)
            acc[j][u][0] = acc[j][u][1] = 0.0;
        auto load = [&](int buf, uint32_t k0) {
          double *Ab = As0 + buf * RQ * RLD, *Bb = Bs0 + buf * FKC * RBL;
          // A: RQ quadrature points x 16 DoFs; element (q, j) at N[q*n + j]: 2048 pairs, 8 per thread
#pragma unroll
          for (int it = 0; it < 8; ++it)
            {
              const int      ch = tid + it * 256;
              const int      r = ch >> 3, kk = (ch & 7) * 2;
              const uint32_t q = q0 + r, j = k0 + kk;
              double *       d = Ab + r * RLD + kk;
              const double * s = N + (size_t)q * n + j;
              if (q < nq && j + 1 < n)
                {
                  if (alignedN)
                    fe_cp16(d, s);
                  else
                    {
                      fe_cp8(d, s);
                      fe_cp8(d + 1, s + 1);
                    }
                }
              else
                {
                  if (q < nq && j < n)
                    fe_cp8(d, s);
                  else
                    d[0] = 0.0;
                  d[1] = 0.0;
                }
            }
          // B: 16 DoFs x 32 vectors gathered through the cell -> DoF map: 256 pairs, one per thread
          {
            const int      r = tid >> 4, cc = (tid & 15) * 2;
            const uint32_t j = k0 + r, v = v0 + cc;
            double *       d = Bb + r * RBL + cc;
            if (j < n && v + 1 < B && alignedX)
              fe_cp16(d, X + (size_t)cid[j] * B + v);
            else
              {
                const double *s = (j < n) ? X + (size_t)cid[j] * B : nullptr;
                if (s && v < B)
                  fe_cp8(d, s + v);
                else
                  d[0] = 0.0;
                if (s && v + 1 < B)
                  fe_cp8(d + 1, s + v + 1);
                else
                  d[1] = 0.0;
              }
          }
        };
        load(0, 0);
        asm volatile("cp.async.commit_group;");
        for (int c = 0; c < nchunks; ++c)
          {
            const int cur = c & 1;
            if (c + 1 < nchunks)
              {
                load(cur ^ 1, (uint32_t)(c + 1) * FKC);
                asm volatile("cp.async.commit_group;");
                asm volatile("cp.async.wait_group 1;");
              }
            else
              asm volatile("cp.async.wait_group 0;");
            __syncthreads();
            if (active)
              {
                const double *as = As0 + cur * RQ * RLD + (wm + (lane >> 2)) * RLD + (lane & 3);
                const double *bs = Bs0 + cur * FKC * RBL + (lane & 3) * RBL + (lane >> 2);
#pragma unroll
                for (int k4 = 0; k4 < FKC / 4; ++k4)
                  {
                    double a[4], bb[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      a[j] = as[j * 8 * RLD + k4 * 4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                      bb[u] = bs[k4 * 4 * RBL + u * 8];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                      for (int u = 0; u < 4; ++u)
                        fe_dmma(acc[j][u][0], acc[j][u][1], a[j], bb[u]);
                  }
              }
            __syncthreads();
          }
        // b += 2 |psi|^2 occ over this pass's vectors (columns v0 + u*8 + (lane&3)*2 + e)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            {
              const uint32_t v = v0 + u * 8 + (lane & 3) * 2 + e;
              const double   o = (v < B) ? occ2[v] : 0.0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                part[j] = fma(acc[j][u][e] * acc[j][u][e], o, part[j]);
            }
      }
    // the 4 lanes of a quad share a row
#pragma unroll
    for (int j = 0; j < 4; ++j)
      {
        part[j] += __shfl_xor_sync(0xffffffffu, part[j], 1);
        part[j] += __shfl_xor_sync(0xffffffffu, part[j], 2);
        const uint32_t q = q0 + wm + j * 8 + (lane >> 2);
        if ((lane & 3) == 0 && q < nq)
          rho[cell.q_off + q] = part[j];
      }
  }

  __global__ void
  fe_weights_kernel(const double *jxw, const double *f, double *w, size_t n)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      w[i] = jxw[i] * f[i]; // hadamardProduct(jxwStorage, f) of FEBasisOperations.t.cpp:170-176
  }
} // namespace hx

using namespace hx;

extern "C"
{
  int
  hx_fe_basis_create(hx_plan *plan, const hx_fe_basis_desc *d, hx_fe_basis **out)
  {
    HX_CHECK(plan && d && out, HX_ERR_INVALID, "null argument");
    HX_CHECK(d->struct_size == sizeof(hx_fe_basis_desc), HX_ERR_INVALID, "hx_fe_basis_desc size mismatch (ABI)");
    HX_CHECK(d->num_cell_quad && d->basis_data && d->jxw, HX_ERR_INVALID, "null array in hx_fe_basis_desc");
    std::unique_ptr<hx_fe_basis> b(new hx_fe_basis());
    b->plan       = plan;
    b->same_basis = d->same_basis_in_all_cells != 0;
    b->C          = plan->C;
    b->h_nq.assign(d->num_cell_quad, d->num_cell_quad + plan->C);
    std::vector<hx_fe_basis::Cell> cells(plan->C);
    size_t                         boff = 0, ooff = 0, qoff = 0;
    for (uint32_t c = 0; c < plan->C; ++c)
      {
        const uint32_t n = plan->h_ncd[c], nq = b->h_nq[c];
        HX_CHECK(!b->same_basis || (n == plan->h_ncd[0] && nq == b->h_nq[0]), HX_ERR_INVALID,
                 "same_basis_in_all_cells needs identical DoF and quadrature counts in every cell");
        HX_CHECK(qoff + nq <= 0xffffffffull, HX_ERR_UNSUPPORTED, "more than 2^32 quadrature points on one rank");
        cells[c].basis_off = b->same_basis ? 0 : boff;
        cells[c].out_off   = ooff;
        cells[c].q_off     = (uint32_t)qoff;
        cells[c].n         = n;
        cells[c].nq        = nq;
        cells[c].ids_off   = plan->h_cell_off[c];
        boff += (size_t)n * nq;
        ooff += (size_t)n * n;
        qoff += nq;
        b->max_n = std::max(b->max_n, n);
      }
    b->n_quad_total = qoff;
    b->S2           = ooff;
    const size_t nbasis = b->same_basis ? (plan->C ? (size_t)plan->h_ncd[0] * b->h_nq[0] : 0) : boff;
    HX_TRY(b->d_basis.upload(d->basis_data, nbasis));
    HX_TRY(b->d_jxw.upload(d->jxw, qoff));
    HX_TRY(b->d_w.alloc(qoff));
    HX_TRY(b->d_cells.upload(cells));
    *out = b.release();
    return HX_OK;
  }

  int
  hx_compute_rho(hx_fe_basis *b, const double *X_dev, uint32_t B, const double *occupation_host, double *rho,
                 int rho_on_device)
  {
    HX_CHECK(b && X_dev && occupation_host && rho, HX_ERR_INVALID, "null argument");
    HX_CHECK(B >= 1, HX_ERR_INVALID, "B must be >= 1");
    hx_plan *p = b->plan;
    if (b->C == 0 || b->n_quad_total == 0)
      return HX_OK;
    std::vector<double> occ2(B);
    for (uint32_t i = 0; i < B; ++i)
      occ2[i] = 2.0 * occupation_host[i]; // b += 2.0 * absSq(psi) * occupation (DensityCalculator.t.cpp:63-64)
    if (b->d_occ.n < B)
      HX_TRY(b->d_occ.alloc(B));
    HX_CUDA(cudaMemcpyAsync(b->d_occ.p, occ2.data(), B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    double *out = rho;
    if (!rho_on_device)
      {
        if (b->d_rho.n < b->n_quad_total)
          HX_TRY(b->d_rho.alloc(b->n_quad_total));
        out = b->d_rho.p;
      }
    uint32_t max_nq = 0;
    for (uint32_t q : b->h_nq)
      max_nq = std::max(max_nq, q);
    const uint32_t qt = (max_nq + RQ - 1) / RQ;
    HX_CHECK(qt <= 65535, HX_ERR_UNSUPPORTED, "more than 16M quadrature points per cell are not supported");
    // per launch (cheap): a second device in the same process needs the opt-in as well
    HX_CUDA(cudaFuncSetAttribute(rho_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RHO_SMEM));
    p->mark("rho:begin");
    dim3 grid(b->C, qt);
    rho_kernel<<<grid, 256, RHO_SMEM, p->stream>>>(b->d_cells.p, b->d_basis.p, p->d_ids.p, X_dev, B, b->d_occ.p, out);
    p->mark("rho");
    p->launches++;
    HX_CUDA(cudaGetLastError());
    if (!rho_on_device)
      HX_CUDA(cudaMemcpyAsync(rho, out, b->n_quad_total * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    HX_CUDA(cudaStreamSynchronize(p->stream)); // occ2 is a host temporary
    return HX_OK;
  }

  int
  hx_fe_basis_destroy(hx_fe_basis *b)
  {
    if (b)
      {
        cudaStreamSynchronize(b->plan->stream);
        delete b;
      }
    return HX_OK;
  }

  static int
  compute_fe_matrices(hx_fe_basis *b, const double *f, int f_on_device, const double *add_to_dev, double *cell_matrices_dev,
                      hx_op *packed_op)
  {
    hx_plan *p = b->plan;
    if (b->C == 0 || b->n_quad_total == 0)
      return HX_OK;
    const double *fd = f;
    if (!f_on_device)
      {
        if (b->d_f.n < b->n_quad_total)
          HX_TRY(b->d_f.alloc(b->n_quad_total));
        HX_CUDA(cudaMemcpyAsync(b->d_f.p, f, b->n_quad_total * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        fd = b->d_f.p;
      }
    fe_weights_kernel<<<(unsigned)((b->n_quad_total + 255) / 256), 256, 0, p->stream>>>(b->d_jxw.p, fd, b->d_w.p, b->n_quad_total);
    const uint32_t tilesM = (b->max_n + FT - 1) / FT;
    const uint32_t tiles  = tilesM * (tilesM + 1) / 2;
    HX_CHECK(tiles <= 65535, HX_ERR_UNSUPPORTED, "cell matrices larger than 23000 x 23000 are not supported");
    p->mark("fe-matrices:begin");
    HX_CHECK((unsigned long long)b->C * tiles < 0x7fffffffull, HX_ERR_UNSUPPORTED, "too many cell-matrix tiles for one launch");
    if (packed_op)
      fe_matrices_kernel<true><<<b->C * tiles, 256, 0, p->stream>>>(b->d_cells.p, b->d_basis.p, b->d_w.p, add_to_dev,
                                                                   packed_op->d_packed.p, tiles, packed_op->d_meta.p,
                                                                   CWARPS * packed_op->mtw, packed_op->kc);
    else
      fe_matrices_kernel<false><<<b->C * tiles, 256, 0, p->stream>>>(b->d_cells.p, b->d_basis.p, b->d_w.p, add_to_dev,
                                                                    cell_matrices_dev, tiles, nullptr, 0, 4);
    p->mark("fe-matrices");
    p->launches += 2;
    HX_CUDA(cudaGetLastError());
    if (!f_on_device)
      HX_CUDA(cudaStreamSynchronize(p->stream)); // the caller may reuse its host array
    return HX_OK;
  }

  int
  hx_compute_fe_matrices(hx_fe_basis *b, const double *f, int f_on_device, const double *add_to_dev, double *cell_matrices_dev)
  {
    HX_CHECK(b && f && cell_matrices_dev, HX_ERR_INVALID, "null argument");
    HX_CHECK(add_to_dev != cell_matrices_dev, HX_ERR_INVALID, "add_to and the output must not alias (mirrored tile writes)");
    return compute_fe_matrices(b, f, f_on_device, add_to_dev, cell_matrices_dev, nullptr);
  }

  int
  hx_cellop_assemble_matrices(hx_op *op, hx_fe_basis *b, const double *f, int f_on_device, const double *add_to_dev)
  {
    HX_CHECK(op && b && f, HX_ERR_INVALID, "null argument");
    HX_CHECK(op->kind == HX_OP_CELL && op->plan == b->plan, HX_ERR_INVALID, "operator and basis data belong to different plans");
    HX_CHECK(!op->share_identical, HX_ERR_INVALID, "matrix sharing is on for this operator");
    // the stream layout, its zero padding and the projector columns are laid down once
    if (!op->have_matrices || op->n_unique != op->plan->C)
      HX_TRY(pack_cell_matrices(op, nullptr, 1));
    HX_TRY(compute_fe_matrices(b, f, f_on_device, add_to_dev, nullptr, op));
    op->have_matrices = true;
    return HX_OK;
  }
}
