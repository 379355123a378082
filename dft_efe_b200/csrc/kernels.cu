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

  // ---- shared machinery of the row kernels -----------------------------------------------------------
  // A thread owns VEC consecutive columns of one row (VEC = 2: 16-byte accesses, B even and 16-B aligned blocks), so a
  // row of B doubles is moved by B / VEC lanes with whole 32-byte sectors.  Each column is computed exactly as in the
  // scalar form, so VEC does not change a single bit of the result.
  template <int VEC>
  struct VD
  {
    double v[VEC];
  };
  template <int VEC>
  __device__ __forceinline__ VD<VEC>
  ldv(const double *p)
  {
    VD<VEC> r;
    if constexpr (VEC == 2)
      {
        const double2 t = *reinterpret_cast<const double2 *>(p);
        r.v[0] = t.x, r.v[1] = t.y;
      }
    else
      r.v[0] = *p;
    return r;
  }
  template <int VEC>
  __device__ __forceinline__ void
  stv(double *p, const VD<VEC> &x)
  {
    if constexpr (VEC == 2)
      *reinterpret_cast<double2 *>(p) = make_double2(x.v[0], x.v[1]);
    else
      *p = x.v[0];
  }
  static inline bool
  aligned16(const void *a, const void *b = nullptr, const void *c = nullptr, const void *d = nullptr)
  {
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d)) & 15) == 0;
  }

  // The fixed-order sums below (constraint rows, parent-side transposes, shared-row slots) are chains of
  // index -> value loads.  Only the additions have to be serial: the index loads of U consecutive terms are issued
  // together, then their value loads, so a chain of m terms costs about 2 m / U memory latencies instead of 2 m; the
  // summation order is the sequential one.  Indices beyond the end
  // are clamped to the last entry (every load unconditional and in range), their terms are not accumulated.
  template <int U, typename I, typename V, typename FI, typename FV, typename FA>
  __device__ __forceinline__ void
  ordered_chain(uint32_t begin, uint32_t end, FI load_idx, FV load_val, FA accumulate)
  {
    if (begin >= end)
      return;
    I idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      idx[u] = load_idx(min(begin + (uint32_t)u, end - 1));
    for (uint32_t e0 = begin; e0 < end; e0 += U)
      {
        V val[U];
        I cur[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          {
            cur[u] = idx[u];
            val[u] = load_val(cur[u]);
          }
        // (kept conditional on purpose: written unconditionally ptxas interleaves the value loads with the additions
        // that wait for them; in this form all U value loads are issued back to back)
        if (e0 + U < end)
          {
#pragma unroll
            for (int u = 0; u < U; ++u)
              idx[u] = load_idx(min(e0 + U + (uint32_t)u, end - 1));
          }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e0 + u < end)
            accumulate(cur[u], val[u]);
      }
  }
  // Chain depth by the longest chain of the data (known at plan creation) and by the size of the launch.  Depth 16 costs
  // 130-190 registers, i.e. ONE 256-thread block per SM: worth it only while the whole launch is a wave or two (the C1
  // mesh: 1352 parents, 18.7 k row-list rows at 8 columns); a large launch (an adaptive mesh of C2's size: 1.8 M
  // row-list threads, 46 waves at one block per SM) is better off with depth 8 and two to three blocks per SM.
  static inline int
  chain_depth(uint32_t longest, size_t threads, int sm_count)
  {
    if (longest <= 1)
      return 1;
    const size_t two_waves = (size_t)2 * 256 * (size_t)(sm_count > 0 ? sm_count : 148);
    return (longest > 32 && threads <= two_waves) ? 16 : 8;
  }
  struct WeightedRow
  {
    uint32_t row;
    double   w;
  };

  // ---- constraints -------------------------------------------------------------------------------
  // distributeParentToChild (src/basis/ConstraintsInternal.cpp:35-108): X[r,:] = inh_r + sum_j w_rj X[col_rj,:].
  // Rows are independent once the constraints are closed (checked at plan creation), so one thread per
  // (row, VEC columns); the j-order of the reference is kept.
  template <int VEC, int U>
  __global__ void
  p2c_kernel(double *X, uint32_t B, uint32_t nR, const uint32_t *rowIds, const uint32_t *rowSizes,
             const uint32_t *rowOffsets, const uint32_t *colIds, const double *colVals, const double *inhom)
  {
    pdl_wait();
    pdl_launch();
    const uint32_t bv = B / VEC;
    const size_t   i  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nR * bv)
      return;
    const uint32_t r = (uint32_t)(i / bv), v = (uint32_t)(i % bv) * VEC;
    VD<VEC>        s;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      s.v[k] = inhom[r];
    const uint32_t o = rowOffsets[r], m = rowSizes[r];
    ordered_chain<U, WeightedRow, VD<VEC>>(
      o, o + m, [&](uint32_t e) { return WeightedRow{colIds[e], colVals[e]}; },
      [&](const WeightedRow &q) { return ldv<VEC>(X + (size_t)q.row * B + v); },
      [&](const WeightedRow &q, const VD<VEC> &x) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          s.v[k] += q.w * x.v[k];
      });
    stv<VEC>(X + (size_t)rowIds[r] * B + v, s);
  }

  // distributeChildToParent (src/basis/ConstraintsInternal.cpp:110-170) without atomics: one thread per
  // (parent, VEC columns) walks the parent-side transpose in the reference's (row, entry) order.
  template <int VEC, int U>
  __global__ void
  c2p_kernel(double *Y, uint32_t B, uint32_t nPar, const uint32_t *parIds, const uint32_t *parOff,
             const uint32_t *parChild, const double *parW)
  {
    pdl_wait();
    pdl_launch();
    const uint32_t bv = B / VEC;
    const size_t   i  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nPar * bv)
      return;
    const uint32_t q = (uint32_t)(i / bv), v = (uint32_t)(i % bv) * VEC;
    double *       y = Y + (size_t)parIds[q] * B + v;
    VD<VEC>        s = ldv<VEC>(y);
    ordered_chain<U, WeightedRow, VD<VEC>>(
      parOff[q], parOff[q + 1], [&](uint32_t e) { return WeightedRow{parChild[e], parW[e]}; },
      [&](const WeightedRow &c) { return ldv<VEC>(Y + (size_t)c.row * B + v); },
      [&](const WeightedRow &c, const VD<VEC> &x) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          s.v[k] += c.w * x.v[k];
      });
    stv<VEC>(y, s);
  }

  template <int VEC>
  __global__ void
  zero_rows_kernel(double *Y, uint32_t B, uint32_t nR, const uint32_t *rowIds)
  {
    pdl_wait();
    pdl_launch();
    const uint32_t bv = B / VEC;
    const size_t   i  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nR * bv)
      return;
    VD<VEC> z;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      z.v[k] = 0.0;
    stv<VEC>(Y + (size_t)rowIds[i / bv] * B + (i % bv) * VEC, z);
  }

// picks the <VEC, U> instantiation: vec_ (bool) and depth_ (1 / 8 / 16) are run-time values
#define HX_VEC_DEPTH_DISPATCH(vec_, depth_, CALL) \
  do                                              \
    {                                             \
      if (vec_)                                   \
        {                                         \
          if ((depth_) == 1)                      \
            CALL(2, 1);                           \
          else if ((depth_) == 8)                 \
            CALL(2, 8);                           \
          else                                    \
            CALL(2, 16);                          \
        }                                         \
      else                                        \
        {                                         \
          if ((depth_) == 1)                      \
            CALL(1, 1);                           \
          else if ((depth_) == 8)                 \
            CALL(1, 8);                           \
          else                                    \
            CALL(1, 16);                          \
        }                                         \
    }                                             \
  while (0)

  int
  launch_p2c(hx_plan *p, double *X, uint32_t B, uint32_t set)
  {
    const ConstraintView c = p->constraint_view(set);
    if (c.nR == 0)
      return HX_OK;
    const bool vec = (B % 2 == 0) && aligned16(X);
    const size_t nthr = (size_t)c.nR * (B / (vec ? 2 : 1));
    const int  dep = chain_depth(c.max_row, nthr, p->sm_count);
    const unsigned nb = nblk(nthr);
#define HX_CALL(V_, U_)                                                                                               \
  HX_CUDA(launch_pdl(p2c_kernel<V_, U_>, nb, 256, 0, p->stream, X, B, c.nR, c.row_ids, c.row_sizes, c.row_offsets, \
                     c.col_ids, c.col_vals, c.inhom))
    HX_VEC_DEPTH_DISPATCH(vec, dep, HX_CALL);
#undef HX_CALL
    p->launches++;
    return HX_OK;
  }

  int
  launch_zero_rows(hx_plan *p, double *Y, uint32_t B, const uint32_t *rows, uint32_t n)
  {
    if (n == 0)
      return HX_OK;
    if ((B % 2 == 0) && aligned16(Y))
      HX_CUDA(launch_pdl(zero_rows_kernel<2>, nblk((size_t)n * (B / 2)), 256, 0, p->stream, Y, B, n, rows));
    else
      HX_CUDA(launch_pdl(zero_rows_kernel<1>, nblk((size_t)n * B), 256, 0, p->stream, Y, B, n, rows));
    p->launches++;
    return HX_OK;
  }

  int
  launch_zero_constrained(hx_plan *p, double *Y, uint32_t B, uint32_t set)
  {
    const ConstraintView c = p->constraint_view(set);
    return launch_zero_rows(p, Y, B, c.row_ids, c.nR);
  }

  int
  launch_c2p(hx_plan *p, double *Y, uint32_t B, uint32_t set, bool zero_rows)
  {
    const ConstraintView c = p->constraint_view(set);
    if (c.nR == 0)
      return HX_OK;
    if (c.nPar)
      {
        const bool     vec = (B % 2 == 0) && aligned16(Y);
        const size_t   nthr = (size_t)c.nPar * (B / (vec ? 2 : 1));
        const int      dep  = chain_depth(c.max_child, nthr, p->sm_count);
        const unsigned nb   = nblk(nthr);
#define HX_CALL(V_, U_) \
  HX_CUDA(launch_pdl(c2p_kernel<V_, U_>, nb, 256, 0, p->stream, Y, B, c.nPar, c.par_ids, c.par_off, c.par_child, c.par_w))
        HX_VEC_DEPTH_DISPATCH(vec, dep, HX_CALL);
#undef HX_CALL
        p->launches++;
      }
    return zero_rows ? launch_zero_constrained(p, Y, B, set) : HX_OK;
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
    ordered_chain<8, uint32_t, double>(
      off[r], off[r + 1], [&](uint32_t e) { return pos[e]; }, [&](uint32_t ps) { return buf[(size_t)ps * B + v]; },
      [&](uint32_t, double t) { s += t; });
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

  // ---- shared-row (enrichment) reduction after the cell kernel's scatter -----------------------------
  // s = sum over e in [begin, end), ascending, of src[slot(e) * B + v .. v + VEC)
  template <int VEC, bool INDIRECT>
  __device__ __forceinline__ VD<VEC>
  ordered_slot_sum(const double *src, const uint32_t *slots, uint32_t begin, uint32_t end, uint32_t B, uint32_t v)
  {
    VD<VEC> s;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      s.v[k] = 0.0;
    ordered_chain<16, uint32_t, VD<VEC>>(
      begin, end, [&](uint32_t e) { return INDIRECT ? slots[e] : e; },
      [&](uint32_t sl) { return ldv<VEC>(src + (size_t)sl * B + v); },
      [&](uint32_t, const VD<VEC> &t) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          s.v[k] += t.v[k];
      });
    return s;
  }
  template <int VEC>
  __global__ void
  shared_reduce_kernel(double *Y, const double *stage, const uint32_t *rows, const uint32_t *off,
                       const uint32_t *slots, uint32_t nrows, uint32_t B)
  {
    pdl_wait();
    pdl_launch();
    const uint32_t bv = B / VEC;
    const size_t   i  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * bv)
      return;
    const uint32_t r = (uint32_t)(i / bv), v = (uint32_t)(i % bv) * VEC;
    // shared rows receive every contribution through a staging slot: Y is written, not accumulated
    stv<VEC>(Y + (size_t)rows[r] * B + v, ordered_slot_sum<VEC, true>(stage, slots, off[r], off[r + 1], B, v));
  }
  // rows shared by very many cells (an enrichment function spans every cell inside its cutoff): two-stage
  // fixed-order reduction - chunks of SH_CHUNK consecutive slots are summed in parallel, then the chunk partials
  // of a row in chunk order.  Deterministic; the grouping differs from the plain ascending sum only in rounding.
  template <int VEC>
  __global__ void
  shared_reduce_chunks_kernel(const double *stage, const uint32_t *chBegin, const uint32_t *chEnd, const uint32_t *slots,
                              double *partial, uint32_t nChunks, uint32_t B)
  {
    pdl_wait();
    pdl_launch();
    const uint32_t bv = B / VEC;
    const size_t   i  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nChunks * bv)
      return;
    const uint32_t k = (uint32_t)(i / bv), v = (uint32_t)(i % bv) * VEC;
    stv<VEC>(partial + (size_t)k * B + v, ordered_slot_sum<VEC, true>(stage, slots, chBegin[k], chEnd[k], B, v));
  }
  template <int VEC>
  __global__ void
  shared_reduce_final_kernel(double *Y, const double *partial, const uint32_t *rows, const uint32_t *chOff, uint32_t nrows,
                             uint32_t B)
  {
    pdl_wait();
    pdl_launch();
    const uint32_t bv = B / VEC;
    const size_t   i  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * bv)
      return;
    const uint32_t r = (uint32_t)(i / bv), v = (uint32_t)(i % bv) * VEC;
    stv<VEC>(Y + (size_t)rows[r] * B + v, ordered_slot_sum<VEC, false>(partial, nullptr, chOff[r], chOff[r + 1], B, v));
  }

  template <int VEC>
  static int
  launch_shared_reduce_v(hx_plan *p, double *Y, uint32_t B)
  {
    const uint32_t bv = B / VEC;
    if (p->n_sh_chunks)
      {
        HX_CUDA(launch_pdl(shared_reduce_chunks_kernel<VEC>, nblk((size_t)p->n_sh_chunks * bv), 256, 0, p->stream, p->d_stage.p,
                           p->d_sh_ch_begin.p, p->d_sh_ch_end.p, p->d_sh_slots.p, p->d_sh_partial.p, p->n_sh_chunks, B));
        HX_CUDA(launch_pdl(shared_reduce_final_kernel<VEC>, nblk((size_t)p->n_shared * bv), 256, 0, p->stream, Y,
                           p->d_sh_partial.p, p->d_sh_rows.p, p->d_sh_ch_off.p, p->n_shared, B));
        p->launches += 2;
        return HX_OK;
      }
    HX_CUDA(launch_pdl(shared_reduce_kernel<VEC>, nblk((size_t)p->n_shared * bv), 256, 0, p->stream, Y, p->d_stage.p,
                       p->d_sh_rows.p, p->d_sh_off.p, p->d_sh_slots.p, p->n_shared, B));
    p->launches++;
    return HX_OK;
  }
  int
  launch_shared_reduce(hx_plan *p, double *Y, uint32_t B)
  {
    if (p->n_shared == 0)
      return HX_OK;
    if ((B % 2 == 0) && aligned16(Y, p->d_stage.p, p->d_sh_partial.p))
      return launch_shared_reduce_v<2>(p, Y, B);
    return launch_shared_reduce_v<1>(p, Y, B);
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

  // ---- global enrichment block: Yenr_owned (nE x B) = block[off .. off+nE, :] (nEg x nEg, column-major) . Xg (nEg x B) ----
  // (gemm('N','T', B, nEg, nEg, Xg, block) of src/basis/OrthoEFEOverlapInverseOpContextGLL.t.cpp:1246-1262, owned rows only)
  __global__ void
  enr_block_global_kernel(const double *blk, uint32_t nEg, uint32_t off, uint32_t nE, const double *Xg, double *Yenr, uint32_t B)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nE * B)
      return;
    const uint32_t j = (uint32_t)(i / B), v = (uint32_t)(i % B);
    double         s = 0.0;
    for (uint32_t k = 0; k < nEg; ++k)
      s += Xg[(size_t)k * B + v] * blk[(size_t)(off + j) + (size_t)k * nEg];
    Yenr[i] = s;
  }
  int
  launch_enr_block_global(hx_plan *p, const double *blk, uint32_t nEg, uint32_t off, uint32_t nE, const double *Xg, double *Yenr,
                          uint32_t B)
  {
    if (nE == 0)
      return HX_OK;
    enr_block_global_kernel<<<nblk((size_t)nE * B), 256, 0, p->stream>>>(blk, nEg, off, nE, Xg, Yenr, B);
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
  template <int VEC>
  struct ChildVal
  {
    double  dinv;
    VD<VEC> s;
  };
  template <int VEC, int U>
  __global__ void
  cheb_fused_kernel(const double *s1, const double *xcur, const double *xprev, double *out, const double *dinv,
                    const uint32_t *rowinfo, const uint32_t *parOff, const uint32_t *parChild, const double *parW,
                    const double *blk, uint32_t ncl, uint32_t nE, uint32_t nRows, uint32_t B, double a, double b,
                    double c, const uint32_t *rows)
  {
    pdl_wait();
    pdl_launch();
    const uint32_t bv = B / VEC;
    const size_t   g  = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)nRows * bv)
      return;
    uint32_t       r = (uint32_t)(g / bv);
    const uint32_t v = (uint32_t)(g % bv) * VEC;
    if (rows)
      r = rows[r];
    const size_t   i = (size_t)r * B + v;
    // the recurrence operands do not depend on the M^-1 part: their loads go first (out may alias xprev - the same
    // thread reads before it writes)
    const VD<VEC> xc = ldv<VEC>(xcur + i);
    VD<VEC>       xp;
    if (c != 0.0)
      xp = ldv<VEC>(xprev + i);
    else
      {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          xp.v[k] = 0.0;
      }
    VD<VEC>        t;
    const uint32_t info = rowinfo[r];
    if (info == 0xFFFFFFFFu && !(r >= ncl && nE > 0))
      {
        // free row of the diagonal M^-1 without hanging-node children: the form the cell kernel's epilogue uses
        const double  d  = dinv[r];
        const VD<VEC> sv = ldv<VEC>(s1 + i);
        VD<VEC>       o;
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          o.v[k] = cheb_combine_diag(a, d, sv.v[k], b, xc.v[k], c, xp.v[k]);
        stv<VEC>(out + i, o);
        return;
      }
    if (info == 0xFFFFFFFEu)
      {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          t.v[k] = 0.0;
      }
    else
      {
        if (r >= ncl && nE > 0)
          {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
              t.v[k] = 0.0;
            const uint32_t j = r - ncl;
            for (uint32_t e = 0; e < nE; ++e)
              {
                const VD<VEC> sv = ldv<VEC>(s1 + (size_t)(ncl + e) * B + v);
                const double  m  = blk[(size_t)e + (size_t)j * nE];
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                  t.v[k] += sv.v[k] * m;
              }
          }
        else
          {
            const double  d  = dinv[r];
            const VD<VEC> sv = ldv<VEC>(s1 + i);
#pragma unroll
            for (int k = 0; k < VEC; ++k)
              t.v[k] = d * sv.v[k];
          }
        if (info != 0xFFFFFFFFu)
          ordered_chain<U, WeightedRow, ChildVal<VEC>>(
            parOff[info], parOff[info + 1], [&](uint32_t e) { return WeightedRow{parChild[e], parW[e]}; },
            [&](const WeightedRow &q) { return ChildVal<VEC>{dinv[q.row], ldv<VEC>(s1 + (size_t)q.row * B + v)}; },
            [&](const WeightedRow &q, const ChildVal<VEC> &ch) {
#pragma unroll
              for (int k = 0; k < VEC; ++k)
                t.v[k] += q.w * (ch.dinv * ch.s.v[k]);
            });
      }
    VD<VEC> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      o.v[k] = cheb_combine(a, t.v[k], b, xc.v[k], c, xp.v[k]);
    stv<VEC>(out + i, o);
  }
} // namespace hx

namespace hx
{
  int
  launch_cheb_fused(hx_plan *p, hx_op *binv, const double *s1, const double *xcur, const double *xprev, double *out,
                    uint32_t B, double a, double b, double c, bool use_row_list, const uint32_t *rows, uint32_t n_rows,
                    int force_depth)
  {
    if (!use_row_list)
      rows = nullptr;
    const uint32_t nr  = use_row_list ? n_rows : p->n_owned;
    const size_t   tot = (size_t)nr * B;
    if (tot == 0)
      return HX_OK;
    const double * xp  = xprev ? xprev : xcur;
    const bool     vec = (B % 2 == 0) && aligned16(s1, xcur, xp, out);
    // this kernel is also the full-vector pass of the unfused filter (a bandwidth kernel over all owned rows): the chain
    // depth stops at 8 so that two 256-thread blocks stay resident per SM
    const size_t   nthr = (size_t)nr * (B / (vec ? 2 : 1));
    const bool     deep = p->max_child > 1 && force_depth != 1;
    // a short row list with long child lists is latency-bound: depth 16; anything longer than two waves keeps depth 8
    const bool     deep16 = deep && use_row_list && chain_depth(p->max_child, nthr, p->sm_count) == 16;
    const unsigned nb   = nblk(nthr);
#define HX_CALL(V_, U_)                                                                                                  \
  HX_CUDA(launch_pdl(cheb_fused_kernel<V_, U_>, nb, 256, 0, p->stream, s1, xcur, xp, out, binv->d_diag.p, p->d_rowinfo.p, \
                     p->d_par_off.p, p->d_par_child.p, p->d_par_w.p, binv->d_enr_block.p, p->n_owned_classical,          \
                     binv->variant == HX_DIAG_CFE ? 0u : binv->nE, nr, B, a, b, xprev ? c : 0.0, rows))
    if (vec)
      {
        if (deep16)
          HX_CALL(2, 16);
        else if (deep)
          HX_CALL(2, 8);
        else
          HX_CALL(2, 1);
      }
    else
      {
        if (deep16)
          HX_CALL(1, 16);
        else if (deep)
          HX_CALL(1, 8);
        else
          HX_CALL(1, 1);
      }
#undef HX_CALL
    p->launches++;
    return HX_OK;
  }
} // namespace hx
