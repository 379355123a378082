// How many warps per SM sub-partition does mma.sync.m8n8k4.f64 (DMMA.8x8x4) need to reach its peak issue rate, with 8 / 16
// independent accumulator pairs per warp, with and without the shared-memory fragment loads of the cell kernel's k loop?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_warps_probe dmma_warps_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC, bool LDS>
__global__ void k(double *out, int iters, double seed)
{
  __shared__ double sm[4 * 1024];
  for (int i = threadIdx.x; i < 4 * 1024; i += blockDim.x)
    sm[i] = seed + i * 1e-9;
  __syncthreads();
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    acc[i][0] = acc[i][1] = 0.0;
  double a = seed + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double *base = sm + (warp % 4) * 1024 + lane;
  for (int it = 0; it < iters; ++it)
    {
      if (LDS)
        {
          // 4 A fragments + 4 B fragments per 16 DMMA (NACC = 16), as in the cell kernel
          double af[4], bf[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            af[j] = base[((it & 7) * 4 + j) * 32];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            bf[j] = base[512 + (((it & 7) * 4 + j) * 9) % 480];
#pragma unroll
          for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(af[i & 3]), "d"(bf[(i >> 2) & 3]));
        }
      else
        {
#pragma unroll
          for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(a), "d"(b));
        }
    }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    s += acc[i][0] + acc[i][1];
  if (s == 123.456)
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, bool LDS>
double run(int sms, int warps_per_sm, double *out)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000 * 8 / NACC;
  // one block per SM with warps_per_sm warps (warp w runs on sub-partition w % 4)
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep)
    {
      cudaEventRecord(e0);
      k<NACC, LDS><<<sms, warps_per_sm * 32>>>(out, iters, 1.0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
  return (double)sms * warps_per_sm * iters * NACC * 512.0 / (best * 1e-3) / 1e12;
}
int main()
{
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  printf("warps/SM (per SMSP) | 8 acc | 16 acc | 16 acc + LDS fragments   [TFLOP/s]\n");
  for (int w : {4, 8, 12, 16, 32})
    printf("%2d (%d) | %6.2f | %6.2f | %6.2f\n", w, w / 4, run<8, false>(sms, w, out), run<16, false>(sms, w, out), run<16, true>(sms, w, out));
  return cudaGetLastError() != cudaSuccess;
}
