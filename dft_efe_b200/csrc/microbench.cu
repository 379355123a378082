// microbench.cu — roofline denominators measured on the device the plan runs on:
//   * FP64 tensor pipe: register-resident mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) chains,
//   * FP64 FMA pipe   : register-resident DFMA chains,
//   * HBM             : 16-B vectorised device copy of 2 x 1 GiB.
// tcgen05.mma has no FP64 kind, so the first number is the "FP64 tensor peak" of SURVEY.md 8(d).
#include "hx_internal.h"

namespace hx
{
  __global__ void __launch_bounds__(256)
  dmma_peak_kernel(double *out, int iters, double seed)
  {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      acc[i][0] = acc[i][1] = 0.0;
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it)
      {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][0]), "+d"(acc[i][1])
                       : "d"(a), "d"(b));
      }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      s += acc[i][0] + acc[i][1];
    if (s == 123.456)
      out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }

  __global__ void __launch_bounds__(256)
  dfma_peak_kernel(double *out, int iters, double seed)
  {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      acc[i] = i * 1e-3;
    const double a = seed + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it)
      {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          acc[i] = fma(acc[i], a, b);
      }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      s += acc[i];
    if (s == 123.456)
      out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }

  __global__ void
  copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n)
  {
    size_t       i      = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
      dst[i] = src[i];
  }
} // namespace hx

using namespace hx;

extern "C" int
hx_microbench(double *dmma_tflops, double *dfma_tflops, double *copy_gbs)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
      set_error("no CUDA device available");
      return HX_ERR_CUDA;
    }
  cudaDeviceProp prop;
  int            dev = 0;
  HX_CUDA(cudaGetDevice(&dev));
  HX_CUDA(cudaGetDeviceProperties(&prop, dev));
  const int   sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  HX_CUDA(cudaEventCreate(&e0));
  HX_CUDA(cudaEventCreate(&e1));
  DevBuf<double> out;
  HX_TRY(out.alloc((size_t)sms * 8 * 256));
  float      ms     = 0.f;
  const int  blocks = sms * 4; // 4 x 8 warps per SM = 8 warps per SMSP
  const int  iters  = 20000;
  double     best;
  // DMMA
  best = 0.0;
  for (int rep = 0; rep < 4; ++rep)
    {
      HX_CUDA(cudaEventRecord(e0));
      dmma_peak_kernel<<<blocks, 256>>>(out.p, iters, 1.0);
      HX_CUDA(cudaEventRecord(e1));
      HX_CUDA(cudaEventSynchronize(e1));
      HX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double fl = (double)blocks * 8 /*warps*/ * iters * 8.0 * 512.0;
      best            = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
  if (dmma_tflops)
    *dmma_tflops = best;
  best = 0.0;
  for (int rep = 0; rep < 4; ++rep)
    {
      HX_CUDA(cudaEventRecord(e0));
      dfma_peak_kernel<<<blocks, 256>>>(out.p, iters, 1.0);
      HX_CUDA(cudaEventRecord(e1));
      HX_CUDA(cudaEventSynchronize(e1));
      HX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double fl = (double)blocks * 256 * iters * 16.0 * 2.0;
      best            = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
  if (dfma_tflops)
    *dfma_tflops = best;
  // copy
  {
    const size_t    bytes = 1ull << 30;
    DevBuf<double2> a, b;
    HX_TRY(a.alloc(bytes / 16));
    HX_TRY(b.alloc(bytes / 16));
    HX_CUDA(cudaMemset(a.p, 1, bytes));
    best = 0.0;
    for (int rep = 0; rep < 6; ++rep)
      {
        HX_CUDA(cudaEventRecord(e0));
        copy_kernel<<<sms * 16, 512>>>(a.p, b.p, bytes / 16);
        HX_CUDA(cudaEventRecord(e1));
        HX_CUDA(cudaEventSynchronize(e1));
        HX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
      }
    if (copy_gbs)
      *copy_gbs = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  HX_CUDA(cudaGetLastError());
  return HX_OK;
}
