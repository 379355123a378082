// gram.cu — the subspace projections as FP64 tensor-core (DMMA) GEMMs.
//
//  gram_block : S[(B-j0) x b] = X[:, j0:]^T . OpX_batch   over the owned rows
//               (blasLapack::gemm('N','C', B-j0, b, nOwned, X+j0, B, OpXb, b) of
//                src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:782-796), split-K over DoF slabs with a
//               fixed-order second-stage reduction (deterministic), tiles strictly above the diagonal skipped
//               (the reference keeps the lower trapezoid only, :819-836).
//  rotate     : X[dof,:] <- X[dof,:] . Qeff   (subspaceRotation, src/linearAlgebra/ElpaScalapackOperations.t.cpp:303-330)
//
// Both stage 64-wide operand tiles in shared memory with cp.async (16-B, double-buffered) and feed
// mma.sync.m8n8k4.f64 from conflict-free padded rows.
#include <stdlib.h>

#include <algorithm>

#include "hx_internal.h"

namespace hx
{
  __device__ __forceinline__ void
  dmma884g(double &d0, double &d1, const double a, const double b)
  {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
  }
  __device__ __forceinline__ void
  cp_async16(void *smem, const void *gmem)
  {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
  }
  __device__ __forceinline__ void
  cp_async_commit()
  {
    asm volatile("cp.async.commit_group;");
  }
  template <int N>
  __device__ __forceinline__ void
  cp_async_wait()
  {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
  }

  constexpr int GT  = 64; // tile edge (outputs)
  constexpr int GKC = 16; // k rows per stage
  constexpr int GLD = GT + 4;

  // load a [GKC x 64] tile of a row-major matrix (row stride ld, columns c0.., rows r0..) into smem[GKC][GLD];
  // out-of-range rows/columns are zero-filled.
  __device__ __forceinline__ void
  load_tile_rows(double *sm, const double *g, size_t ld, size_t r0, size_t rend, uint32_t c0, uint32_t cend,
                 bool aligned, int tid)
  {
    // 16 rows x 32 double2 chunks = 512 chunks, 256 threads -> 2 each
#pragma unroll
    for (int it = 0; it < 2; ++it)
      {
        const int      ch = tid + it * 256;
        const int      r  = ch >> 5;
        const int      cc = (ch & 31) * 2;
        double *       d  = sm + r * GLD + cc;
        const size_t   gr = r0 + r;
        const uint32_t gc = c0 + cc;
        if (gr < rend && gc + 1 < cend && aligned)
          cp_async16(d, g + gr * ld + gc);
        else
          {
            d[0] = (gr < rend && gc < cend) ? g[gr * ld + gc] : 0.0;
            d[1] = (gr < rend && gc + 1 < cend) ? g[gr * ld + gc + 1] : 0.0;
          }
      }
  }

  // grid: (tiles, nSplit).  W: [nSplit][M*N] col-major partials.
  __global__ void __launch_bounds__(256)
  gram_kernel(const double *X, uint32_t B, uint32_t j0, const double *O, uint32_t b, size_t nOwned, uint32_t M,
              uint32_t N, uint32_t tilesM, size_t slab, double *W, int alignedX, int alignedO)
  {
    __shared__ __align__(16) double As[2][GKC * GLD];
    __shared__ __align__(16) double Bs[2][GKC * GLD];
    const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tm = blockIdx.x % tilesM, tn = blockIdx.x / tilesM;
    const uint32_t m0 = tm * GT, n0 = tn * GT;
    if (m0 + GT <= n0)
      return; // tile strictly above the diagonal (rows j < cols i): not part of the lower trapezoid
    const size_t kb = (size_t)blockIdx.y * slab;
    const size_t ke = kb + slab < nOwned ? kb + slab : nOwned;
    double       acc[2][4][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        acc[j][t][0] = acc[j][t][1] = 0.0;
    const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
    if (kb < ke)
      {
        const int nchunks = (int)((ke - kb + GKC - 1) / GKC);
        load_tile_rows(As[0], X, B, kb, ke, j0 + m0, j0 + M, alignedX, tid);
        load_tile_rows(Bs[0], O, b, kb, ke, n0, N, alignedO, tid);
        cp_async_commit();
        for (int c = 0; c < nchunks; ++c)
          {
            const int cur = c & 1;
            if (c + 1 < nchunks)
              {
                load_tile_rows(As[cur ^ 1], X, B, kb + (size_t)(c + 1) * GKC, ke, j0 + m0, j0 + M, alignedX, tid);
                load_tile_rows(Bs[cur ^ 1], O, b, kb + (size_t)(c + 1) * GKC, ke, n0, N, alignedO, tid);
                cp_async_commit();
                cp_async_wait<1>();
              }
            else
              cp_async_wait<0>();
            __syncthreads();
            const double *as = As[cur] + (lane & 3) * GLD + wm + (lane >> 2);
            const double *bs = Bs[cur] + (lane & 3) * GLD + wn + (lane >> 2);
#pragma unroll
            for (int k4 = 0; k4 < GKC / 4; ++k4)
              {
                double a[2], bb[4];
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  a[j] = as[k4 * 4 * GLD + j * 8];
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  bb[t] = bs[k4 * 4 * GLD + t * 8];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                  for (int t = 0; t < 4; ++t)
                    dmma884g(acc[j][t][0], acc[j][t][1], a[j], bb[t]);
              }
            __syncthreads();
          }
      }
    double *w = W + (size_t)blockIdx.y * M * N;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        {
          const uint32_t r = m0 + wm + j * 8 + (lane >> 2);
          const uint32_t c = n0 + wn + t * 8 + (lane & 3) * 2;
          if (r < M)
            {
              if (c < N)
                w[(size_t)r + (size_t)c * M] = acc[j][t][0];
              if (c + 1 < N)
                w[(size_t)r + (size_t)(c + 1) * M] = acc[j][t][1];
            }
        }
  }

  __global__ void
  gram_reduce_kernel(const double *W, uint32_t M, uint32_t N, uint32_t nSplit, double *S)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)M * N)
      return;
    const uint32_t r = (uint32_t)(i % M), c = (uint32_t)(i / M);
    // tiles strictly above the diagonal were skipped
    if ((r / GT) * GT + GT <= (c / GT) * GT)
      {
        S[i] = 0.0;
        return;
      }
    double s = 0.0;
    for (uint32_t k = 0; k < nSplit; ++k)
      s += W[(size_t)k * M * N + i];
    S[i] = s;
  }

  // ---------------------------------------------------------------------------------------------------
  // Narrow blocks (B - j0 <= 32 and b <= 32: the C1 / C2 shapes): the Gram block is one 32 x 32 tile and the GEMM is
  // HBM-bound (16 N B bytes), so there is nothing to stage: every warp streams its own k-steps (4 rows of X and of
  // Op X) straight from global memory into DMMA fragments - each 256-B row segment is consumed whole by the warp -
  // and the 8 warps of a CTA are summed in warp order through one shared tile (deterministic).  grid = row slabs.
  __global__ void __launch_bounds__(256)
  gram_small_kernel(const double *X, uint32_t B, uint32_t j0, const double *O, uint32_t b, size_t nOwned, uint32_t M,
                    uint32_t N, size_t slab, double *W)
  {
    __shared__ double tile[32 * 33];
    const int         tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t      kb = (size_t)blockIdx.x * slab;
    const size_t      ke = kb + slab < nOwned ? kb + slab : nOwned;
    double            acc[4][4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        acc[j][t][0] = acc[j][t][1] = 0.0;
    const uint32_t mi = lane >> 2; // row of the fragment inside an 8-wide tile
    auto           load = [&](size_t r0, double(&a)[4], double(&q)[4]) {
      const size_t row = r0 + (lane & 3);
      const bool   ok  = row < ke;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        {
          const uint32_t m = j * 8 + mi;
          a[j]             = (ok && m < M) ? __ldg(X + row * B + j0 + m) : 0.0;
          q[j]             = (ok && m < N) ? __ldg(O + row * b + m) : 0.0;
        }
    };
    double a0[4], q0[4], a1[4], q1[4];
    size_t r = kb + (size_t)warp * 4;
    if (r < ke)
      load(r, a0, q0);
    for (; r < ke; r += 64)
      {
        const bool more = r + 32 < ke;
        if (more)
          load(r + 32, a1, q1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int t = 0; t < 4; ++t)
            dmma884g(acc[j][t][0], acc[j][t][1], a0[j], q0[t]);
        if (!more)
          break;
        if (r + 64 < ke)
          load(r + 64, a0, q0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int t = 0; t < 4; ++t)
            dmma884g(acc[j][t][0], acc[j][t][1], a1[j], q1[t]);
      }
    // sum the warps in warp order
    for (int w = 0; w < 8; ++w)
      {
        if (warp == w)
          {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                  {
                    double *d = tile + (j * 8 + mi) * 33 + t * 8 + (lane & 3) * 2 + e;
                    *d        = (w == 0) ? acc[j][t][e] : *d + acc[j][t][e];
                  }
          }
        __syncthreads();
      }
    double *w_ = W + (size_t)blockIdx.x * M * N;
    for (uint32_t i = tid; i < M * N; i += 256)
      {
        const uint32_t rr = i % M, cc = i / M;
        w_[i]             = tile[rr * 33 + cc];
      }
  }

  int
  gram_block(hx_plan *p, const double *X, uint32_t B, uint32_t j0, const double *OpXb, uint32_t b, size_t nOwned,
             double *S_dev)
  {
    const uint32_t M = B - j0, N = b;
    if (M <= 32 && N <= 32 && !getenv("HXB200_GRAM_TILED"))
      {
        // narrow block: one 32 x 32 tile, split over row slabs (2 resident CTAs per SM, >= 256 rows each)
        uint32_t nSplit = 2u * (uint32_t)std::max(p->sm_count, 1);
        nSplit          = (uint32_t)std::max<size_t>(1, std::min<size_t>(nSplit, (nOwned + 255) / 256));
        size_t slab     = (nOwned + nSplit - 1) / nSplit;
        slab            = std::max<size_t>(32, (slab + 31) / 32 * 32);
        nSplit          = (uint32_t)std::max<size_t>(1, (nOwned + slab - 1) / slab);
        double *W = S_dev + (size_t)M * N;
        HX_CHECK(p->d_small.p && S_dev >= p->d_small.p &&
                   (size_t)(S_dev - p->d_small.p) + (size_t)M * N * (1 + (size_t)nSplit) <= p->d_small.n,
                 HX_ERR_INVALID, "gram workspace too small");
        gram_small_kernel<<<nSplit, 256, 0, p->stream>>>(X, B, j0, OpXb, b, nOwned, M, N, slab, W);
        gram_reduce_kernel<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, p->stream>>>(W, M, N, nSplit, S_dev);
        p->launches += 2;
        HX_CUDA(cudaGetLastError());
        return HX_OK;
      }
    const uint32_t tilesM = (M + GT - 1) / GT, tilesN = (N + GT - 1) / GT;
    const uint32_t tiles  = tilesM * tilesN;
    // tiles on and below the diagonal do the work (the others exit at once): split K so that about 4 CTAs per SM
    // (64 registers x 256 threads) are resident and all of them run in one wave
    uint32_t active = 0;
    for (uint32_t tn = 0; tn < tilesN; ++tn)
      active += tilesM > tn ? tilesM - tn : 0;
    active = std::max(active, 1u);
    const uint32_t slots    = 4u * (uint32_t)std::max(p->sm_count, 1);
    uint32_t       nSplit   = std::max(1u, slots / active);
    const uint32_t maxSplit = (uint32_t)std::max<size_t>(1, (nOwned + 255) / 256);
    nSplit                  = std::max(1u, std::min(nSplit, maxSplit));
    size_t slab             = (nOwned + nSplit - 1) / nSplit;
    slab                    = (slab + GKC - 1) / GKC * GKC;
    if (slab == 0)
      slab = GKC;
    nSplit = (uint32_t)std::max<size_t>(1, (nOwned + slab - 1) / slab);
    // workspace lives behind S in the same small buffer (caller sized it: see hx_xtopx)
    double *W = S_dev + (size_t)M * N;
    HX_CHECK(p->d_small.p && S_dev >= p->d_small.p &&
               (size_t)(S_dev - p->d_small.p) + (size_t)M * N * (1 + (size_t)nSplit) <= p->d_small.n,
             HX_ERR_INVALID, "gram workspace too small");
    const int alignedX = (B % 2 == 0) && (j0 % 2 == 0) && (((uintptr_t)X & 15) == 0);
    const int alignedO = (b % 2 == 0) && (((uintptr_t)OpXb & 15) == 0);
    dim3      grid(tiles, nSplit);
    gram_kernel<<<grid, 256, 0, p->stream>>>(X, B, j0, OpXb, b, nOwned, M, N, tilesM, slab, W, alignedX, alignedO);
    gram_reduce_kernel<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, p->stream>>>(W, M, N, nSplit, S_dev);
    p->launches += 2;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  // doubles gram_block needs behind S_dev for an M x N block over nOwned rows: the block itself + one partial per K split
  // (the same split counts gram_block chooses)
  size_t
  gram_workspace_doubles(const hx_plan *p, uint32_t B, uint32_t batch, size_t nOwned)
  {
    // the column batches of computeXTransOpX: block j0 is (B - j0) x b; the later (smaller) blocks have fewer active tiles
    // and are split further, so the maximum is taken over all of them
    batch        = std::max(1u, std::min(batch, B));
    size_t worst = 0;
    for (uint32_t j0 = 0; j0 < B; j0 += batch)
      {
        const uint32_t M = B - j0, N = std::min(batch, B - j0);
        uint32_t       nSplit;
        if (M <= 32 && N <= 32)
          nSplit = 2u * (uint32_t)std::max(p->sm_count, 1);
        else
          {
            const uint32_t tilesM = (M + GT - 1) / GT, tilesN = (N + GT - 1) / GT;
            uint32_t       active = 0;
            for (uint32_t tn = 0; tn < tilesN; ++tn)
              active += tilesM > tn ? tilesM - tn : 0;
            nSplit = std::max(1u, 4u * (uint32_t)std::max(p->sm_count, 1) / std::max(active, 1u));
          }
        nSplit = (uint32_t)std::max<size_t>(1, std::min<size_t>(nSplit, (nOwned + 255) / 256)) + 1; // + 1: slab rounding
        worst  = std::max(worst, (size_t)M * N * (1 + (size_t)nSplit));
      }
    return worst;
  }

  // ---------------------------------------------------------------------------------------------------
  constexpr int RLDA = GKC + 4;
  // Out[r, n0..n0+63] = sum_k X[r, k] Q[k, n]   (Q row-major K x N = B x B)
  __global__ void __launch_bounds__(256)
  rotate_kernel(const double *X, uint32_t B, size_t nRows, const double *Q, double *Out, int lowerTri, int transpose,
                int aligned)
  {
    __shared__ __align__(16) double As[2][GT * RLDA];
    __shared__ __align__(16) double Bs[2][GKC * GLD];
    const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tilesN = (B + GT - 1) / GT;
    const uint32_t tn     = blockIdx.x % tilesN;
    const size_t   tmi    = blockIdx.x / tilesN;
    const size_t   r0     = tmi * GT;
    const uint32_t n0     = tn * GT;
    uint32_t       kbeg = 0, kend = B;
    if (lowerTri)
      {
        if (transpose)
          kbeg = (n0 / GKC) * GKC; // Qeff[i][j] = Q(i,j) != 0 only for i >= j
        else
          kend = (n0 + GT < B) ? n0 + GT : B; // Qeff[i][j] = Q(j,i) != 0 only for i <= j
      }
    double acc[2][4][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        acc[j][t][0] = acc[j][t][1] = 0.0;
    const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;

    auto load = [&](int buf, uint32_t k0) {
      // A: 64 rows x 16 k  -> 64*8 double2 chunks = 512 -> 2 per thread
#pragma unroll
      for (int it = 0; it < 2; ++it)
        {
          const int      ch = tid + it * 256;
          const int      r  = ch >> 3;
          const int      kk = (ch & 7) * 2;
          double *       d  = As[buf] + r * RLDA + kk;
          const size_t   gr = r0 + r;
          const uint32_t gk = k0 + kk;
          if (gr < nRows && gk + 1 < kend && aligned)
            cp_async16(d, X + gr * B + gk);
          else
            {
              d[0] = (gr < nRows && gk < kend) ? X[gr * B + gk] : 0.0;
              d[1] = (gr < nRows && gk + 1 < kend) ? X[gr * B + gk + 1] : 0.0;
            }
        }
      load_tile_rows(Bs[buf], Q, B, k0, kend, n0, B, aligned, tid);
    };

    const int nchunks = (kend > kbeg) ? (int)((kend - kbeg + GKC - 1) / GKC) : 0;
    if (nchunks)
      {
        load(0, kbeg);
        cp_async_commit();
      }
    for (int c = 0; c < nchunks; ++c)
      {
        const int cur = c & 1;
        if (c + 1 < nchunks)
          {
            load(cur ^ 1, kbeg + (uint32_t)(c + 1) * GKC);
            cp_async_commit();
            cp_async_wait<1>();
          }
        else
          cp_async_wait<0>();
        __syncthreads();
        const double *as = As[cur] + (wm + (lane >> 2)) * RLDA + (lane & 3);
        const double *bs = Bs[cur] + (lane & 3) * GLD + wn + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < GKC / 4; ++k4)
          {
            double a[2], bb[4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
              a[j] = as[j * 8 * RLDA + k4 * 4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              bb[t] = bs[k4 * 4 * GLD + t * 8];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int t = 0; t < 4; ++t)
                dmma884g(acc[j][t][0], acc[j][t][1], a[j], bb[t]);
          }
        __syncthreads();
      }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        {
          const size_t   r = r0 + wm + j * 8 + (lane >> 2);
          const uint32_t c = n0 + wn + t * 8 + (lane & 3) * 2;
          if (r < nRows)
            {
              if (c < B)
                Out[r * B + c] = acc[j][t][0];
              if (c + 1 < B)
                Out[r * B + c + 1] = acc[j][t][1];
            }
        }
  }

  // Narrow blocks (B <= 32): the whole rotation matrix lives in registers as DMMA B fragments, every warp streams
  // tiles of 16 rows of X straight from global memory, and writes them back IN PLACE (a row of the result depends on
  // that row of X only, and the warp holds all of it before it stores) - no staging buffer, no copy back: 16 N B bytes.
  __global__ void __launch_bounds__(128, 3)
  rotate_small_kernel(double *X, uint32_t B, size_t nRows, const double *Q)
  {
    const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ki = lane & 3, mi = lane >> 2;
    double         q[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int t = 0; t < 4; ++t)
        {
          const uint32_t k = ks * 4 + ki, n = t * 8 + mi;
          q[ks][t]         = (k < B && n < B) ? __ldg(Q + (size_t)k * B + n) : 0.0;
        }
    const size_t nTiles = (nRows + 15) / 16;
    for (size_t rt = (size_t)blockIdx.x * 4 + warp; rt < nTiles; rt += (size_t)gridDim.x * 4)
      {
        double a[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
          {
            const size_t row = rt * 16 + j * 8 + mi;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              {
                const uint32_t k = ks * 4 + ki;
                a[j][ks]         = (row < nRows && k < B) ? X[row * B + k] : 0.0;
              }
          }
        double acc[2][4][2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int t = 0; t < 4; ++t)
            acc[j][t][0] = acc[j][t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int t = 0; t < 4; ++t)
              dmma884g(acc[j][t][0], acc[j][t][1], a[j][ks], q[ks][t]);
        __syncwarp(); // every lane's loads of these 16 rows have been consumed by the (warp-synchronous) mma
#pragma unroll
        for (int j = 0; j < 2; ++j)
          {
            const size_t row = rt * 16 + j * 8 + mi;
            if (row < nRows)
#pragma unroll
              for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                  {
                    const uint32_t c = t * 8 + ki * 2 + e;
                    if (c < B)
                      X[row * B + c] = acc[j][t][e];
                  }
          }
      }
  }

  int
  rotate(hx_plan *p, double *X, uint32_t B, size_t nOwned, const double *Q_dev, int transpose, int lowerTri,
         double *tmp, size_t tmp_rows)
  {
    if (nOwned == 0)
      return HX_OK;
    if (B <= 32 && !getenv("HXB200_GRAM_TILED"))
      {
        const size_t nTiles = (nOwned + 15) / 16;
        unsigned     grid   = (unsigned)std::min<size_t>((nTiles + 3) / 4, (size_t)3 * std::max(p->sm_count, 1));
        rotate_small_kernel<<<grid, 128, 0, p->stream>>>(X, B, nOwned, Q_dev);
        p->launches++;
        HX_CUDA(cudaGetLastError());
        return HX_OK;
      }
    // a row's image depends on that row only: slabs of tmp_rows rows (a multiple of the tile height) are rotated into the
    // scratch block and copied back in place
    const uint32_t tilesN  = (B + GT - 1) / GT;
    const int      aligned = (B % 2 == 0) && (((uintptr_t)X & 15) == 0) && (((uintptr_t)Q_dev & 15) == 0);
    const size_t   slab    = tmp_rows ? std::max<size_t>(GT, tmp_rows / GT * GT) : nOwned;
    for (size_t r0 = 0; r0 < nOwned; r0 += slab)
      {
        const size_t rows   = std::min(slab, nOwned - r0);
        const size_t tilesM = (rows + GT - 1) / GT;
        double *     Xs     = X + r0 * B;
        rotate_kernel<<<(unsigned)(tilesM * tilesN), 256, 0, p->stream>>>(Xs, B, rows, Q_dev, tmp, lowerTri, transpose, aligned);
        p->launches++;
        HX_CUDA(cudaGetLastError());
        HX_CUDA(cudaMemcpyAsync(Xs, tmp, rows * B * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
      }
    return HX_OK;
  }
} // namespace hx
