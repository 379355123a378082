// dense.cu — the dense B x B problems of the subspace iteration, on the device.
//
// The reference hands these to ELPA / ScaLAPACK on the host: elpa_cholesky + ScaLAPACKMatrix::invert
// (src/linearAlgebra/OrthonormalizationFunctions.t.cpp:204-309; potrf + trtri in the serial branch :361-374) and
// elpa_eigenvectors (src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:154-164; heevd in the serial branch :245-254).
// They are library calls there and library calls here: cuSOLVER (potrf, trtri, syevd), resolved with dlopen like NCCL
// so that the hot path carries no link-time dependency, run on the plan's stream so the projected matrix never leaves
// the device between the Gram GEMM and the subspace rotation.  O(B^3), not a roofline target.
#include <dlfcn.h>

#include "hx_internal.h"

namespace hx
{
  typedef struct cusolverDnContext *cusolverDnHandle_t;
  enum
  {
    CUSOLVER_OK = 0
  };

  struct Dense
  {
    void *             lib    = nullptr;
    cusolverDnHandle_t handle = nullptr;
    int (*Create)(cusolverDnHandle_t *)                                                                   = nullptr;
    int (*Destroy)(cusolverDnHandle_t)                                                                    = nullptr;
    int (*SetStream)(cusolverDnHandle_t, cudaStream_t)                                                    = nullptr;
    int (*PotrfBufferSize)(cusolverDnHandle_t, int, int, double *, int, int *)                            = nullptr;
    int (*Potrf)(cusolverDnHandle_t, int, int, double *, int, double *, int, int *)                       = nullptr;
    int (*TrtriBufferSize)(cusolverDnHandle_t, int, int, int64_t, int, void *, int64_t, size_t *, size_t *) = nullptr;
    int (*Trtri)(cusolverDnHandle_t, int, int, int64_t, int, void *, int64_t, void *, size_t, void *, size_t,
                 int *)                                                                                   = nullptr;
    int (*SyevdBufferSize)(cusolverDnHandle_t, int, int, int, const double *, int, const double *, int *) = nullptr;
    int (*Syevd)(cusolverDnHandle_t, int, int, int, double *, int, double *, double *, int, int *)        = nullptr;
    DevBuf<double>    work;
    DevBuf<int>       info;
    std::vector<char> host_work;
  };

  void
  dense_destroy(Dense *d)
  {
    if (!d)
      return;
    if (d->handle && d->Destroy)
      d->Destroy(d->handle);
    // the library stays loaded (other plans may share it)
    delete d;
  }

  static int
  dense_get(hx_plan *p, Dense **out)
  {
    if (p->dense)
      {
        *out = p->dense;
        return HX_OK;
      }
    const char *names[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so",
                           "/usr/local/cuda/lib64/libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so"};
    void *      h       = nullptr;
    for (const char *n : names)
      {
        h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (h)
          break;
      }
    HX_CHECK(h, HX_ERR_UNSUPPORTED, "cannot dlopen libcusolver.so.11 (dense B x B solves): %s", dlerror());
    Dense *d = new Dense();
    d->lib   = h;
#define HX_SYM(field, name)                       \
  *(void **)(&d->field) = dlsym(h, name);         \
  if (!d->field)                                  \
    {                                             \
      set_error("cuSOLVER symbol %s missing", name); \
      delete d;                                   \
      return HX_ERR_UNSUPPORTED;                  \
    }
    HX_SYM(Create, "cusolverDnCreate");
    HX_SYM(Destroy, "cusolverDnDestroy");
    HX_SYM(SetStream, "cusolverDnSetStream");
    HX_SYM(PotrfBufferSize, "cusolverDnDpotrf_bufferSize");
    HX_SYM(Potrf, "cusolverDnDpotrf");
    HX_SYM(TrtriBufferSize, "cusolverDnXtrtri_bufferSize");
    HX_SYM(Trtri, "cusolverDnXtrtri");
    HX_SYM(SyevdBufferSize, "cusolverDnDsyevd_bufferSize");
    HX_SYM(Syevd, "cusolverDnDsyevd");
#undef HX_SYM
    int rc = d->Create(&d->handle);
    if (rc != CUSOLVER_OK)
      {
        set_error("cusolverDnCreate failed (%d)", rc);
        d->handle = nullptr;
        delete d;
        return HX_ERR_CUDA;
      }
    rc = d->SetStream(d->handle, p->stream);
    if (rc != CUSOLVER_OK || d->info.alloc(4) != HX_OK)
      {
        set_error("cusolverDnSetStream failed (%d)", rc);
        dense_destroy(d);
        return HX_ERR_CUDA;
      }
    p->dense = d;
    *out     = d;
    return HX_OK;
  }

#define HX_SOLVER(call)                                               \
  do                                                                  \
    {                                                                 \
      int s_ = (call);                                                \
      if (s_ != CUSOLVER_OK)                                          \
        {                                                             \
          set_error("%s:%d: %s -> cusolver status %d", __FILE__, __LINE__, #call, s_); \
          return HX_ERR_CUDA;                                         \
        }                                                             \
    }                                                                 \
  while (0)

  // zero the strict upper triangle of a column-major B x B matrix (what the reference's "extract LConj" loop does,
  // OrthonormalizationFunctions.t.cpp:263-276)
  __global__ void
  zero_strict_upper_kernel(double *A, uint32_t B)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * B)
      return;
    const uint32_t r = (uint32_t)(i % B), c = (uint32_t)(i / B);
    if (r < c)
      A[i] = 0.0;
  }

  // S (lower triangle valid) -> full symmetric: the reference's projHam + projHam^T with the diagonal halved
  // (RayleighRitzEigenSolver.t.cpp:137-152)
  __global__ void
  symmetrize_from_lower_kernel(double *A, uint32_t B)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * B)
      return;
    const uint32_t r = (uint32_t)(i % B), c = (uint32_t)(i / B);
    if (r < c)
      A[i] = A[(size_t)c + (size_t)r * B];
  }

  __global__ void
  transpose_kernel(const double *A, double *At, uint32_t B)
  {
    __shared__ double t[32][33];
    const uint32_t    c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y)
      {
        const uint32_t r = r0 + threadIdx.x, c = c0 + j;
        t[j][threadIdx.x] = (r < B && c < B) ? A[(size_t)r + (size_t)c * B] : 0.0;
      }
    __syncthreads();
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y)
      {
        const uint32_t r = c0 + threadIdx.x, c = r0 + j; // At(r, c) = A(c, r)
        if (r < B && c < B)
          At[(size_t)r + (size_t)c * B] = t[threadIdx.x][j];
      }
  }

  // Gram block of one column batch ((B-j0) x b, column-major, leading dimension M) into the B x B matrix: only the
  // entries the reference writes into the ScaLAPACK matrix, i.e. rows j >= column (RayleighRitzEigenSolver.t.cpp:819-836)
  __global__ void
  place_gram_block_kernel(const double *Sd, uint32_t M, uint32_t b, uint32_t j0, double *S, uint32_t B)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * b)
      return;
    const uint32_t r = (uint32_t)(i % B), c = (uint32_t)(i / B);
    S[(size_t)r + (size_t)(c + j0) * B] = (r >= c + j0) ? Sd[(size_t)c * M + (r - j0)] : 0.0;
  }

  int
  dense_place_gram_block(hx_plan *p, const double *Sd, uint32_t M, uint32_t b, uint32_t j0, double *S, uint32_t B)
  {
    const size_t tot = (size_t)B * b;
    place_gram_block_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, p->stream>>>(Sd, M, b, j0, S, B);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  int
  dense_transpose(hx_plan *p, const double *A, double *At, uint32_t B)
  {
    dim3 grid((B + 31) / 32, (B + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, p->stream>>>(A, At, B);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }

  static int
  fetch_info(hx_plan *p, Dense *d, int *info_host)
  {
    int h = 0;
    HX_CUDA(cudaMemcpyAsync(&h, d->info.p, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    HX_TRY(plan_sync(p));
    *info_host = h;
    return HX_OK;
  }

  // S (lower triangle) = L L^T; on return S holds L^-1 (lower triangular, strict upper triangle zero).
  // info: 0 ok; k > 0 the leading minor of order k is not positive definite (potrf) or L(k,k) == 0 (trtri).
  int
  dense_cholesky_inverse(hx_plan *p, double *S, uint32_t B, int *info_host)
  {
    Dense *d;
    HX_TRY(dense_get(p, &d));
    const int CUBLAS_FILL_MODE_LOWER = 0, CUBLAS_DIAG_NON_UNIT = 0, CUDA_R_64F = 1;
    int       lwork = 0;
    HX_SOLVER(d->PotrfBufferSize(d->handle, CUBLAS_FILL_MODE_LOWER, (int)B, S, (int)B, &lwork));
    size_t wdev = 0, whost = 0;
    HX_SOLVER(d->TrtriBufferSize(d->handle, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, (int64_t)B, CUDA_R_64F, S,
                                 (int64_t)B, &wdev, &whost));
    const size_t need = std::max((size_t)lwork, (wdev + 7) / 8) + 8;
    if (d->work.n < need)
      HX_TRY(d->work.alloc(need));
    if (d->host_work.size() < whost + 8)
      d->host_work.resize(whost + 8);
    HX_SOLVER(d->Potrf(d->handle, CUBLAS_FILL_MODE_LOWER, (int)B, S, (int)B, d->work.p, lwork, d->info.p));
    p->launches++;
    HX_TRY(fetch_info(p, d, info_host));
    if (*info_host != 0)
      return HX_OK;
    zero_strict_upper_kernel<<<(unsigned)(((size_t)B * B + 255) / 256), 256, 0, p->stream>>>(S, B);
    p->launches++;
    HX_SOLVER(d->Trtri(d->handle, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, (int64_t)B, CUDA_R_64F, S, (int64_t)B,
                       d->work.p, wdev, d->host_work.data(), whost, d->info.p));
    p->launches++;
    HX_TRY(fetch_info(p, d, info_host));
    return HX_OK;
  }

  // S: lower triangle of a symmetric B x B matrix (column-major).  On return S holds the orthonormal eigenvectors
  // (column j <-> evals[j], ascending) and evals_dev the eigenvalues.
  int
  dense_sym_eig(hx_plan *p, double *S, uint32_t B, double *evals_dev, int *info_host)
  {
    Dense *d;
    HX_TRY(dense_get(p, &d));
    const int CUBLAS_FILL_MODE_LOWER = 0, CUSOLVER_EIG_MODE_VECTOR = 1;
    symmetrize_from_lower_kernel<<<(unsigned)(((size_t)B * B + 255) / 256), 256, 0, p->stream>>>(S, B);
    p->launches++;
    int lwork = 0;
    HX_SOLVER(d->SyevdBufferSize(d->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)B, S, (int)B,
                                 evals_dev, &lwork));
    if (d->work.n < (size_t)lwork + 8)
      HX_TRY(d->work.alloc((size_t)lwork + 8));
    HX_SOLVER(d->Syevd(d->handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)B, S, (int)B, evals_dev,
                       d->work.p, lwork, d->info.p));
    p->launches++;
    HX_TRY(fetch_info(p, d, info_host));
    return HX_OK;
  }
} // namespace hx
