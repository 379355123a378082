// kernels.cu — the HBM-bound helpers around the cell contraction: hanging-node constraints, halo
// pack/unpack/accumulate, row scaling, fused Chebyshev recurrence, column norms, shared-row reduction.
// All are bandwidth kernels: coalesced row-contiguous accesses (a row of the block vector is B contiguous
// doubles), 16-B vector accesses when B is even, grids sized from the data.
#include "hx_internal.h"

namespace hx
{
  static inline unsigned
  nblk(size_t n, unsigned t = 256)
  {
    return (unsigned)((n + t - 1) / t);
  }

  // The fixed-order sums below (constraint rows, parent-side transposes, shared-row slots) are chains of
  // index -> value loads.  Only the additions have to be serial: the loads of ILP consecutive terms are issued
  // together, so a chain of m terms costs m / ILP memory latencies instead of m (summation order unchanged).
  constexpr int ILP    = 8;
  constexpr int ILP_SH = 16;

  // ---- constraints -------------------------------------------------------------------------------
  // distributeParentToChild (src/basis/ConstraintsInternal.cpp:35-108): X[r,:] = inh_r + sum_j w_rj X[col_rj,:].
  // Rows are independent once the constraints are closed (checked at plan creation), so one thread per
  // (row, vector); the j-order of the reference is kept.
  __global__ void
  p2c_kernel(double *X, uint32_t B, uint32_t nR, const uint32_t *rowIds, const uint32_t *rowSizes,
             const uint32_t *rowOffsets, const uint32_t *colIds, const double *colVals, const double *inhom)
  {
    pdl_wait();
    pdl_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nR * B)
      return;
    const uint32_t r = (uint32_t)(i / B), v = (uint32_t)(i % B);
    double         s = inhom[r];
    const uint32_t o = rowOffsets[r], m = rowSizes[r];
    for (uint32_t j0 = 0; j0 < m; j0 += ILP)
      {
        uint32_t cid[ILP];
        double   w[ILP], x[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          {
            const uint32_t j = min(j0 + u, m - 1); // clamped: every load is unconditional and in range
            cid[u]           = colIds[o + j];
            w[u]             = colVals[o + j];
          }
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          x[u] = X[(size_t)cid[u] * B + v];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          if (j0 + u < m)
            s += w[u] * x[u];
      }
    X[(size_t)rowIds[r] * B + v] = s;
  }

  // distributeChildToParent (src/basis/ConstraintsInternal.cpp:110-170) without atomics: one thread per
  // (parent, vector) walks the parent-side transpose in the reference's (row, entry) order.
  __global__ void
  c2p_kernel(double *Y, uint32_t B, uint32_t nPar, const uint32_t *parIds, const uint32_t *parOff,
             const uint32_t *parChild, const double *parW)
  {
    pdl_wait();
    pdl_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nPar * B)
      return;
    const uint32_t q = (uint32_t)(i / B), v = (uint32_t)(i % B);
    double *       y = Y + (size_t)parIds[q] * B + v;
    double         s = *y;
    const uint32_t end = parOff[q + 1];
    for (uint32_t e0 = parOff[q]; e0 < end; e0 += ILP)
      {
        uint32_t ch[ILP];
        double   w[ILP], x[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          {
            const uint32_t e = min(e0 + u, end - 1);
            ch[u]            = parChild[e];
            w[u]             = parW[e];
          }
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          x[u] = Y[(size_t)ch[u] * B + v];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          if (e0 + u < end)
            s += w[u] * x[u];
      }
    *y = s;
  }

  __global__ void
  zero_rows_kernel(double *Y, uint32_t B, uint32_t nR, const uint32_t *rowIds)
  {
    pdl_wait();
    pdl_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nR * B)
      return;
    Y[(size_t)rowIds[i / B] * B + (i % B)] = 0.0;
  }

  int
  launch_p2c(hx_plan *p, double *X, uint32_t B, uint32_t set)
  {
    const ConstraintView c = p->constraint_view(set);
    if (c.nR == 0)
      return HX_OK;
    HX_CUDA(launch_pdl(p2c_kernel, nblk((size_t)c.nR * B), 256, 0, p->stream, X, B, c.nR, c.row_ids, c.row_sizes, c.row_offsets,
                       c.col_ids, c.col_vals, c.inhom));
    p->launches++;
    return HX_OK;
  }

  int
  launch_zero_constrained(hx_plan *p, double *Y, uint32_t B, uint32_t set)
  {
    const ConstraintView c = p->constraint_view(set);
    if (c.nR == 0)
      return HX_OK;
    HX_CUDA(launch_pdl(zero_rows_kernel, nblk((size_t)c.nR * B), 256, 0, p->stream, Y, B, c.nR, c.row_ids));
    p->launches++;
    return HX_OK;
  }

  int
  launch_zero_rows(hx_plan *p, double *Y, uint32_t B, const uint32_t *rows, uint32_t n)
  {
    if (n == 0)
      return HX_OK;
    HX_CUDA(launch_pdl(zero_rows_kernel, nblk((size_t)n * B), 256, 0, p->stream, Y, B, n, rows));
    p->launches++;
    return HX_OK;
  }

  int
  launch_c2p(hx_plan *p, double *Y, uint32_t B, uint32_t set)
  {
    const ConstraintView c = p->constraint_view(set);
    if (c.nR == 0)
      return HX_OK;
    if (c.nPar)
      {
        HX_CUDA(launch_pdl(c2p_kernel, nblk((size_t)c.nPar * B), 256, 0, p->stream, Y, B, c.nPar, c.par_ids, c.par_off,
                           c.par_child, c.par_w));
        p->launches++;
      }
    HX_CUDA(cudaGetLastError());
    return launch_zero_constrained(p, Y, B, set);
  }

  // ---- halo pack / unpack / accumulate (src/utils/DiscontiguousDataOperations.cpp:36-92) ------------
  __global__ void
  pack_rows_kernel(const double *x, uint32_t B, const uint32_t *ids, uint32_t n, double *buf)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * B)
      return;
    buf[i] = x[(size_t)ids[i / B] * B + (i % B)];
  }
  __global__ void
  unpack_rows_kernel(const double *buf, uint32_t B, const uint32_t *ids, uint32_t n, double *x)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * B)
      return;
    x[(size_t)ids[i / B] * B + (i % B)] = buf[i];
  }
  // accumulate: ids may repeat (one owned row wanted by several ranks); a CSR row -> buffer positions
  // built at plan creation keeps the reference's buffer order and needs no atomics.
  __global__ void
  add_rows_kernel(const double *buf, uint32_t B, const uint32_t *rows, const uint32_t *off, const uint32_t *pos,
                  uint32_t nrows, double *x)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * B)
      return;
    const uint32_t r = (uint32_t)(i / B), v = (uint32_t)(i % B);
    double *       d = x + (size_t)rows[r] * B + v;
    double         s = *d;
    const uint32_t end = off[r + 1];
    for (uint32_t e0 = off[r]; e0 < end; e0 += ILP)
      {
        uint32_t ps[ILP];
        double   t[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          ps[u] = pos[min(e0 + u, end - 1)];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          t[u] = buf[(size_t)ps[u] * B + v];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
          if (e0 + u < end)
            s += t[u];
      }
    *d = s;
  }

  int
  launch_pack(hx_plan *p, const double *x, uint32_t B, const uint32_t *ids, uint32_t n, double *buf)
  {
    if (n == 0)
      return HX_OK;
    pack_rows_kernel<<<nblk((size_t)n * B), 256, 0, p->stream>>>(x, B, ids, n, buf);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }
  int
  launch_unpack(hx_plan *p, const double *buf, uint32_t B, const uint32_t *ids, uint32_t n, double *x)
  {
    if (n == 0)
      return HX_OK;
    unpack_rows_kernel<<<nblk((size_t)n * B), 256, 0, p->stream>>>(buf, B, ids, n, x);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }
  int
  launch_add_rows(hx_plan *p, const double *buf, uint32_t B, const uint32_t *rows, const uint32_t *off,
                  const uint32_t *pos, uint32_t nrows, double *x)
  {
    if (nrows == 0)
      return HX_OK;
    add_rows_kernel<<<nblk((size_t)nrows * B), 256, 0, p->stream>>>(buf, B, rows, off, pos, nrows, x);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // ---- elementwise ---------------------------------------------------------------------------------
  // khatriRaoProduct(ColMajor,1,B,N) (src/linearAlgebra/BlasLapackKernels.cpp:356-372): y[i,:] = d[i] x[i,:]
  __global__ void
  row_scale_kernel(const double *d, const double *x, double *y, uint32_t B, size_t total)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total)
      y[i] = d[i / B] * x[i];
  }
  int
  launch_row_scale(hx_plan *p, const double *d, const double *x, double *y, uint32_t B, size_t nrows)
  {
    const size_t tot = nrows * B;
    if (tot == 0)
      return HX_OK;
    row_scale_kernel<<<nblk(tot), 256, 0, p->stream>>>(d, x, y, B, tot);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  __global__ void
  axpby_kernel(size_t n, double a, const double *x, double b, const double *y, double *z)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      z[i] = a * x[i] + b * y[i];
  }
  int
  launch_axpby(hx_plan *p, size_t n, double a, const double *x, double b, const double *y, double *z)
  {
    if (n == 0)
      return HX_OK;
    axpby_kernel<<<nblk(n), 256, 0, p->stream>>>(n, a, x, b, y, z);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  __global__ void
  axpby_blocked_kernel(size_t total, uint32_t B, double a1, const double *a, const double *x, double b1,
                       const double *b, const double *y, double *z)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total)
      {
        const uint32_t j = (uint32_t)(i % B);
        z[i]             = a1 * a[j] * x[i] + b1 * b[j] * y[i];
      }
  }
  int
  launch_axpby_blocked(hx_plan *p, size_t nrows, uint32_t B, double a1, const double *a, const double *x, double b1,
                       const double *b, const double *y, double *z)
  {
    const size_t tot = nrows * B;
    if (tot == 0)
      return HX_OK;
    axpby_blocked_kernel<<<nblk(tot), 256, 0, p->stream>>>(tot, B, a1, a, x, b1, b, y, z);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // ---- column sums of squares (MultiVector::l2Norms, src/linearAlgebra/MultiVector.t.cpp:553-578) ----
  // deterministic two-stage reduction: each block reduces a contiguous slab of rows per column in a fixed
  // order, a second kernel adds the block partials in block order.
  constexpr int CS_BLOCKS = 592; // 4 x 148 SMs
  __global__ void
  colsumsq_partial_kernel(const double *x, const double *y, uint32_t B, size_t nrows, double *partial)
  {
    extern __shared__ double sh[]; // [rowsPerIter][B]
    const uint32_t           rpi   = blockDim.x / B; // rows handled per iteration (>=1 because B <= blockDim)
    const uint32_t           rr    = threadIdx.x / B, c = threadIdx.x % B;
    const size_t             per   = (nrows + gridDim.x - 1) / gridDim.x;
    const size_t             begin = (size_t)blockIdx.x * per;
    const size_t             end   = begin + per < nrows ? begin + per : nrows;
    double                   s     = 0.0;
    if (rr < rpi)
      for (size_t r = begin + rr; r < end; r += rpi)
        {
          s += x[r * B + c] * y[r * B + c];
        }
    if (rr < rpi)
      sh[rr * B + c] = s;
    __syncthreads();
    if (threadIdx.x < B)
      {
        double t = 0.0;
        for (uint32_t q = 0; q < rpi; ++q)
          t += sh[q * B + threadIdx.x];
        partial[(size_t)blockIdx.x * B + threadIdx.x] = t;
      }
  }
  __global__ void
  colsumsq_final_kernel(const double *partial, uint32_t B, uint32_t nb, double *out)
  {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B)
      return;
    double s = 0.0;
    for (uint32_t b = 0; b < nb; ++b)
      s += partial[(size_t)b * B + c];
    out[c] = s;
  }
  int
  launch_colsumsq(hx_plan *p, const double *x, uint32_t B, size_t nrows, double *out_dev)
  {
    return launch_coldot(p, x, x, B, nrows, out_dev);
  }

  // column dot products over the first nrows rows (MultiVector dot / l2Norms, src/linearAlgebra/MultiVector.t.cpp:
  // 553-578, 805-870): same deterministic two-stage reduction
  int
  launch_coldot(hx_plan *p, const double *x, const double *y, uint32_t B, size_t nrows, double *out_dev)
  {
    HX_CHECK(B <= 256, HX_ERR_UNSUPPORTED, "column reductions: B > 256 must be called per column batch");
    HX_TRY(p->ensure_small((size_t)CS_BLOCKS * B + B));
    double *partial = p->d_small.p;
    const unsigned threads = 256;
    colsumsq_partial_kernel<<<CS_BLOCKS, threads, (threads / B) * B * sizeof(double), p->stream>>>(x, y, B, nrows, partial);
    colsumsq_final_kernel<<<nblk(B), 256, 0, p->stream>>>(partial, B, CS_BLOCKS, out_dev);
    p->launches += 2;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // ---- fused BLAS-1 of one CG iteration (CGLinearSolver.t.cpp:170-260) for a diagonal preconditioner ----
  // cg_dots2: partial column sums of z.r and p.w over the owned rows (one pass instead of two reductions)
  __global__ void
  cg_dots2_partial_kernel(const double *z, const double *r, const double *pd, const double *w, uint32_t B, size_t nrows,
                          double *partial)
  {
    extern __shared__ double sh[]; // [2][rowsPerIter][B]
    const uint32_t           rpi = blockDim.x / B, rr = threadIdx.x / B, c = threadIdx.x % B;
    const size_t             per = (nrows + gridDim.x - 1) / gridDim.x;
    const size_t             begin = (size_t)blockIdx.x * per, end = begin + per < nrows ? begin + per : nrows;
    double                   s0 = 0.0, s1 = 0.0;
    if (rr < rpi)
      for (size_t q = begin + rr; q < end; q += rpi)
        {
          s0 += z[q * B + c] * r[q * B + c];
          s1 += pd[q * B + c] * w[q * B + c];
        }
    if (rr < rpi)
      {
        sh[rr * B + c]             = s0;
        sh[(rpi + rr) * B + c]     = s1;
      }
    __syncthreads();
    if (threadIdx.x < 2 * B)
      {
        const uint32_t which = threadIdx.x / B, col = threadIdx.x % B;
        double         t = 0.0;
        for (uint32_t q = 0; q < rpi; ++q)
          t += sh[(which * rpi + q) * B + col];
        partial[((size_t)blockIdx.x * 2 + which) * B + col] = t;
      }
  }
  // sums the block partials in block order; out0 = first dot, out1 = second dot, quot = out0 / den (den = out1 when
  // den_in is null), nquot = -quot
  __global__ void
  cg_finalize_kernel(const double *partial, uint32_t B, uint32_t nb, double *out0, double *out1, const double *den_in,
                     int quot_of_first_over_den, double *quot, double *nquot)
  {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B)
      return;
    double s0 = 0.0, s1 = 0.0;
    for (uint32_t b = 0; b < nb; ++b)
      {
        s0 += partial[((size_t)b * 2 + 0) * B + c];
        s1 += partial[((size_t)b * 2 + 1) * B + c];
      }
    out0[c] = s0;
    out1[c] = s1;
    if (quot_of_first_over_den)
      {
        const double q = s0 / (den_in ? den_in[c] : s1);
        quot[c]        = q;
        if (nquot)
          nquot[c] = -q;
      }
  }
  // x += alpha p, r -= alpha w over the owned rows; z = dinv .* r over the local rows (PreconditionerJacobi::apply
  // with both ghost flags false); partial column sums of z.r and r.r over the owned rows
  __global__ void
  cg_update_partial_kernel(double *x, const double *pd, double *r, const double *w, double *z, const double *dinv,
                           const double *alpha, uint32_t B, size_t nowned, size_t nlocal, double *partial)
  {
    extern __shared__ double sh[];
    const uint32_t           rpi = blockDim.x / B, rr = threadIdx.x / B, c = threadIdx.x % B;
    const size_t             per = (nlocal + gridDim.x - 1) / gridDim.x;
    const size_t             begin = (size_t)blockIdx.x * per, end = begin + per < nlocal ? begin + per : nlocal;
    double                   s0 = 0.0, s1 = 0.0;
    if (rr < rpi)
      {
        const double al = alpha[c];
        for (size_t q = begin + rr; q < end; q += rpi)
          {
            const size_t i = q * B + c;
            double       rv = r[i];
            if (q < nowned)
              {
                x[i] = 1.0 * x[i] + al * pd[i];
                rv   = 1.0 * rv + (-al) * w[i];
                r[i] = rv;
              }
            const double zv = dinv[q] * rv;
            z[i]            = zv;
            if (q < nowned)
              {
                s0 += zv * rv;
                s1 += rv * rv;
              }
          }
      }
    if (rr < rpi)
      {
        sh[rr * B + c]         = s0;
        sh[(rpi + rr) * B + c] = s1;
      }
    __syncthreads();
    if (threadIdx.x < 2 * B)
      {
        const uint32_t which = threadIdx.x / B, col = threadIdx.x % B;
        double         t = 0.0;
        for (uint32_t q = 0; q < rpi; ++q)
          t += sh[(which * rpi + q) * B + col];
        partial[((size_t)blockIdx.x * 2 + which) * B + col] = t;
      }
  }
  int
  launch_cg_dots2(hx_plan *p, const double *z, const double *r, const double *pd, const double *w, uint32_t B, double *zdotr,
                  double *pdotw, double *alpha, double *nalpha)
  {
    HX_CHECK(B <= 128, HX_ERR_UNSUPPORTED, "fused CG reductions: B <= 128");
    HX_TRY(p->ensure_small((size_t)2 * CS_BLOCKS * B + 16 * (size_t)B));
    const unsigned threads = 256;
    cg_dots2_partial_kernel<<<CS_BLOCKS, threads, 2 * (threads / B) * B * sizeof(double), p->stream>>>(z, r, pd, w, B, p->n_owned,
                                                                                                      p->d_small.p);
    cg_finalize_kernel<<<nblk(B), 256, 0, p->stream>>>(p->d_small.p, B, CS_BLOCKS, zdotr, pdotw, nullptr, 1, alpha, nalpha);
    p->launches += 2;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }
  int
  launch_cg_update(hx_plan *p, double *x, const double *pd, double *r, const double *w, double *z, const double *dinv,
                   const double *alpha, uint32_t B, double *zdotr_new, double *rr, const double *zdotr_old, double *beta)
  {
    HX_CHECK(B <= 128, HX_ERR_UNSUPPORTED, "fused CG reductions: B <= 128");
    HX_TRY(p->ensure_small((size_t)2 * CS_BLOCKS * B + 16 * (size_t)B));
    const unsigned threads = 256;
    cg_update_partial_kernel<<<CS_BLOCKS, threads, 2 * (threads / B) * B * sizeof(double), p->stream>>>(
      x, pd, r, w, z, dinv, alpha, B, p->n_owned, p->n_local, p->d_small.p);
    cg_finalize_kernel<<<nblk(B), 256, 0, p->stream>>>(p->d_small.p, B, CS_BLOCKS, zdotr_new, rr, zdotr_old, 1, beta, nullptr);
    p->launches += 2;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // per-column quotient of two reduction results, kept on the device (step lengths of the CG solver)
  __global__ void
  col_divide_kernel(const double *num, const double *den, double *out, double *out_neg, uint32_t B)
  {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B)
      return;
    const double q = num[j] / den[j];
    out[j]         = q;
    if (out_neg)
      out_neg[j] = -q;
  }
  int
  launch_col_divide(hx_plan *p, const double *num, const double *den, double *out, double *out_neg, uint32_t B)
  {
    col_divide_kernel<<<nblk(B), 256, 0, p->stream>>>(num, den, out, out_neg, B);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // ---- shared-row (enrichment) reduction after the coloured scatter --------------------------------
  // s = sum over e in [begin, end), ascending, of src[slot(e) * B + v]
  template <bool INDIRECT>
  __device__ __forceinline__ double
  ordered_slot_sum(const double *src, const uint32_t *slots, uint32_t begin, uint32_t end, uint32_t B, uint32_t v)
  {
    double s = 0.0;
    for (uint32_t e0 = begin; e0 < end; e0 += ILP_SH)
      {
        uint32_t sl[ILP_SH];
        double   t[ILP_SH];
#pragma unroll
        for (int u = 0; u < ILP_SH; ++u)
          {
            const uint32_t e = min(e0 + u, end - 1); // clamped: every load is unconditional and in range
            sl[u]            = INDIRECT ? slots[e] : e;
          }
#pragma unroll
        for (int u = 0; u < ILP_SH; ++u)
          t[u] = src[(size_t)sl[u] * B + v];
#pragma unroll
        for (int u = 0; u < ILP_SH; ++u)
          if (e0 + u < end)
            s += t[u];
      }
    return s;
  }
  __global__ void
  shared_reduce_kernel(double *Y, const double *stage, const uint32_t *rows, const uint32_t *off,
                       const uint32_t *slots, uint32_t nrows, uint32_t B)
  {
    pdl_wait();
    pdl_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * B)
      return;
    const uint32_t r = (uint32_t)(i / B), v = (uint32_t)(i % B);
    // shared rows receive every contribution through a staging slot: Y is written, not accumulated
    Y[(size_t)rows[r] * B + v] = ordered_slot_sum<true>(stage, slots, off[r], off[r + 1], B, v);
  }
  // rows shared by very many cells (an enrichment function spans every cell inside its cutoff): two-stage
  // fixed-order reduction - chunks of SH_CHUNK consecutive slots are summed in parallel, then the chunk partials
  // of a row in chunk order.  Deterministic; the grouping differs from the plain ascending sum only in rounding.
  __global__ void
  shared_reduce_chunks_kernel(const double *stage, const uint32_t *chBegin, const uint32_t *chEnd, const uint32_t *slots,
                              double *partial, uint32_t nChunks, uint32_t B)
  {
    pdl_wait();
    pdl_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nChunks * B)
      return;
    const uint32_t k = (uint32_t)(i / B), v = (uint32_t)(i % B);
    partial[i] = ordered_slot_sum<true>(stage, slots, chBegin[k], chEnd[k], B, v);
  }
  __global__ void
  shared_reduce_final_kernel(double *Y, const double *partial, const uint32_t *rows, const uint32_t *chOff, uint32_t nrows,
                             uint32_t B)
  {
    pdl_wait();
    pdl_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * B)
      return;
    const uint32_t r = (uint32_t)(i / B), v = (uint32_t)(i % B);
    Y[(size_t)rows[r] * B + v] = ordered_slot_sum<false>(partial, nullptr, chOff[r], chOff[r + 1], B, v);
  }

  int
  launch_shared_reduce(hx_plan *p, double *Y, uint32_t B)
  {
    if (p->n_shared == 0)
      return HX_OK;
    if (p->n_sh_chunks)
      {
        HX_CUDA(launch_pdl(shared_reduce_chunks_kernel, nblk((size_t)p->n_sh_chunks * B), 256, 0, p->stream, p->d_stage.p,
                           p->d_sh_ch_begin.p, p->d_sh_ch_end.p, p->d_sh_slots.p, p->d_sh_partial.p, p->n_sh_chunks, B));
        HX_CUDA(launch_pdl(shared_reduce_final_kernel, nblk((size_t)p->n_shared * B), 256, 0, p->stream, Y, p->d_sh_partial.p,
                           p->d_sh_rows.p, p->d_sh_ch_off.p, p->n_shared, B));
        p->launches += 2;
        return HX_OK;
      }
    HX_CUDA(launch_pdl(shared_reduce_kernel, nblk((size_t)p->n_shared * B), 256, 0, p->stream, Y, p->d_stage.p, p->d_sh_rows.p,
                       p->d_sh_off.p, p->d_sh_slots.p, p->n_shared, B));
    p->launches++;
    return HX_OK;
  }

  // ---- atom-block enrichment matrix: Yenr (B x nE) = Xenr (B x nE) * blk (nE x nE), col-major ----------
  // (src/basis/OEFEAtomBlockOverlapInvOpContextGLL.t.cpp:1005-1021)
  __global__ void
  enr_block_kernel(const double *blk, uint32_t nE, const double *Xenr, double *Yenr, uint32_t B)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nE * B)
      return;
    const uint32_t j = (uint32_t)(i / B), v = (uint32_t)(i % B);
    double         s = 0.0;
    for (uint32_t k = 0; k < nE; ++k)
      s += Xenr[(size_t)k * B + v] * blk[(size_t)k + (size_t)j * nE];
    Yenr[i] = s;
  }
  int
  launch_enr_block(hx_plan *p, const double *blk, uint32_t nE, const double *Xenr, double *Yenr, uint32_t B)
  {
    if (nE == 0)
      return HX_OK;
    enr_block_kernel<<<nblk((size_t)nE * B), 256, 0, p->stream>>>(blk, nE, Xenr, Yenr, B);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // ---- fused Chebyshev recurrence step ---------------------------------------------------------------
  // One pass over the owned rows does, for the diagonal (mass-lumped) M^-1 of the reference,
  //   t      = C2P( dinv .* P2C(s1) ) [+ atom-block rows]      (M^-1 apply, a11)
  //   out    = a*t + b*xcur + c*xprev                           (the two axpby of ChebyshevFilter.t.cpp:105-124)
  // s1 must already have its constrained rows filled (p2c launched before).  rowinfo[i]: 0xFFFFFFFF = free row
  // without children, 0xFFFFFFFE = constrained row (-> t = 0), else index into the parent-side CSR.
  __global__ void
  cheb_fused_kernel(const double *s1, const double *xcur, const double *xprev, double *out, const double *dinv,
                    const uint32_t *rowinfo, const uint32_t *parOff, const uint32_t *parChild, const double *parW,
                    const double *blk, uint32_t ncl, uint32_t nE, uint32_t nRows, uint32_t B, double a, double b,
                    double c, const uint32_t *rows)
  {
    pdl_wait();
    pdl_launch();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nRows * B)
      return;
    uint32_t r = (uint32_t)(i / B);
    const uint32_t v = (uint32_t)(i % B);
    if (rows)
      {
        r = rows[r];
        i = (size_t)r * B + v;
      }
    double         t;
    const uint32_t info = rowinfo[r];
    if (info == 0xFFFFFFFEu)
      t = 0.0;
    else
      {
        if (r >= ncl && nE > 0)
          {
            t = 0.0;
            const uint32_t j = r - ncl;
            for (uint32_t k = 0; k < nE; ++k)
              t += s1[(size_t)(ncl + k) * B + v] * blk[(size_t)k + (size_t)j * nE];
          }
        else
          t = dinv[r] * s1[i];
        if (info != 0xFFFFFFFFu)
          {
            const uint32_t end = parOff[info + 1];
            for (uint32_t e0 = parOff[info]; e0 < end; e0 += ILP)
              {
                uint32_t ch[ILP];
                double   w[ILP], dv[ILP], sv[ILP];
#pragma unroll
                for (int u = 0; u < ILP; ++u)
                  {
                    const uint32_t e = min(e0 + u, end - 1);
                    ch[u]            = parChild[e];
                    w[u]             = parW[e];
                  }
#pragma unroll
                for (int u = 0; u < ILP; ++u)
                  {
                    dv[u] = dinv[ch[u]];
                    sv[u] = s1[(size_t)ch[u] * B + v];
                  }
#pragma unroll
                for (int u = 0; u < ILP; ++u)
                  if (e0 + u < end)
                    t += w[u] * (dv[u] * sv[u]);
              }
          }
      }
    out[i] = cheb_combine(a, t, b, xcur[i], c, c != 0.0 ? xprev[i] : 0.0);
  }
} // namespace hx

namespace hx
{
  int
  launch_cheb_fused(hx_plan *p, hx_op *binv, const double *s1, const double *xcur, const double *xprev, double *out,
                    uint32_t B, double a, double b, double c, bool use_row_list, const uint32_t *rows, uint32_t n_rows)
  {
    if (!use_row_list)
      rows = nullptr;
    const uint32_t nr  = use_row_list ? n_rows : p->n_owned;
    const size_t   tot = (size_t)nr * B;
    if (tot == 0)
      return HX_OK;
    HX_CUDA(launch_pdl(cheb_fused_kernel, nblk(tot), 256, 0, p->stream, s1, xcur, xprev ? xprev : xcur, out, binv->d_diag.p,
                       p->d_rowinfo.p, p->d_par_off.p, p->d_par_child.p, p->d_par_w.p, binv->d_enr_block.p,
                       p->n_owned_classical, binv->variant == HX_DIAG_CFE ? 0u : binv->nE, nr, B, a, b, xprev ? c : 0.0, rows));
    p->launches++;
    return HX_OK;
  }
} // namespace hx
