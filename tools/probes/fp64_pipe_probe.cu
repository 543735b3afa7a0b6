// Micro-probe: is the fp64 tensor path (DMMA) on sm_100a a separate pipe from the DFMA pipe?
// Measures DFMA-only, DMMA-only (m8n8k4 and, if available, m16n8k8) and a 1:1 interleave.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters, double a, double b)
{
   double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
   for (int i = 0; i < iters; i++)
   {
#pragma unroll
      for (int u = 0; u < 8; u++)
      {
         x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
         x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
      }
   }
   out[blockIdx.x*blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double *out, int iters, double a, double b)
{
   double c[8][2];
#pragma unroll
   for (int k = 0; k < 8; k++) { c[k][0] = threadIdx.x + k; c[k][1] = k; }
   for (int i = 0; i < iters; i++)
   {
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
         for (int k = 0; k < 8; k++) { dmma884(c[k][0], c[k][1], a, b); }
   }
   double s = 0; for (int k = 0; k < 8; k++) { s += c[k][0] + c[k][1]; }
   out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
__global__ void k_both(double *out, int iters, double a, double b)
{
   double c[4][2];
   double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll
   for (int k = 0; k < 4; k++) { c[k][0] = threadIdx.x + k; c[k][1] = k; }
   for (int i = 0; i < iters; i++)
   {
#pragma unroll
      for (int u = 0; u < 4; u++)
      {
         // 4 DMMA (4*256 = 1024 FMA-equivalents per warp) : 32 DFMA warp instructions (1024 FMA per warp)
#pragma unroll
         for (int k = 0; k < 4; k++)
         {
            dmma884(c[k][0], c[k][1], a, b);
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
         }
      }
   }
   double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7; for (int k = 0; k < 4; k++) { s += c[k][0] + c[k][1]; }
   out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
// shared-memory bandwidth probe: LDS.64 conflict-free
__global__ void k_lds(double *out, int iters)
{
   __shared__ double s[4096];
   for (int i = threadIdx.x; i < 4096; i += blockDim.x) { s[i] = i; }
   __syncthreads();
   double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
   int p = threadIdx.x;
   for (int i = 0; i < iters; i++)
   {
#pragma unroll
      for (int u = 0; u < 8; u++)
      {
         a0 += s[(p + 0) & 4095]; a1 += s[(p + 256) & 4095]; a2 += s[(p + 512) & 4095]; a3 += s[(p + 768) & 4095];
         p += 1024;
      }
   }
   out[blockIdx.x*blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

template<typename F> static float time_ms(F f)
{
   cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
   f(); cudaDeviceSynchronize();
   cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
   float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main()
{
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   const int sms = p.multiProcessorCount, T = 256, BPS = 4, G = sms*BPS, iters = 2000;
   double *out; cudaMalloc(&out, sizeof(double)*G*T);
   printf("device %s, %d SMs\n", p.name, sms);
   float ms = time_ms([&] { k_dfma<<<G, T>>>(out, iters, 1.0000001, 1e-9); });
   double fma_n = (double)G*T*iters*64;
   printf("DFMA only : %.3f ms  %.2f TFLOP/s\n", ms, 2*fma_n/ms*1e-9);
   ms = time_ms([&] { k_dmma<<<G, T>>>(out, iters, 1.0000001, 1e-9); });
   double mma_n = (double)G*(T/32)*iters*32*256;
   printf("DMMA only : %.3f ms  %.2f TFLOP/s\n", ms, 2*mma_n/ms*1e-9);
   ms = time_ms([&] { k_both<<<G, T>>>(out, iters, 1.0000001, 1e-9); });
   double both_n = (double)G*(T/32)*iters*16*(256 + 8*32);
   printf("DMMA+DFMA : %.3f ms  %.2f TFLOP/s (1:1 flop mix)\n", ms, 2*both_n/ms*1e-9);
   ms = time_ms([&] { k_lds<<<G, T>>>(out, iters); });
   double bytes = (double)G*T*iters*32*8;
   printf("LDS.64    : %.3f ms  %.1f B/clk/SM at %.0f MHz\n", ms, bytes/(ms*1e-3)/sms/(p.clockRate*1e3), p.clockRate*1e-3);
   cudaFree(out);
   return 0;
}
