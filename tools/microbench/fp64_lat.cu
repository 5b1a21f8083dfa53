// Microbenchmark: dependent-issue latency and per-SM throughput of fp64 ops on sm_100a (development aid).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_kernel(double* out, long long* cyc, double a, double b) {
  double x = a, y = b;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) x = fma(x, y, a);
  }
  long long t1 = clock64();
  double z = a;
#pragma unroll 1
  for (int i = 0; i < 256; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) z = z + y;
  }
  long long t2 = clock64();
  double w = a;
#pragma unroll 1
  for (int i = 0; i < 256; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) w = w * y;
  }
  long long t3 = clock64();
  double s = a;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) s = __shfl_xor_sync(0xffffffffu, s, 1) + y;
  }
  long long t4 = clock64();
  double r = a;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) r = rsqrt(r + 1.5);
  }
  long long t5 = clock64();
  double q = a;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) q = 1.0 / (q + 1.5);
  }
  long long t6 = clock64();
  double u = a;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) u = sqrt(u + 1.5);
  }
  long long t7 = clock64();
  float ff = (float)a;
#pragma unroll 1
  for (int i = 0; i < 256; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) ff = fmaf(ff, (float)b, (float)a);
  }
  long long t8 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7;
  }
  out[threadIdx.x + blockIdx.x * blockDim.x] = x + z + w + s + r + q + u + ff;
}
// throughput: ILP=8 independent chains per thread
__global__ void thr_kernel(double* out, long long* cyc, double a, double b) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = a + k;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 512; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = fma(x[k], b, a);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 << 20); cudaMalloc(&cyc, 64);
  long long h[8];
  lat_kernel<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
  cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles/op): DFMA %.1f  DADD %.1f  DMUL %.1f  SHFL64+DADD %.1f  rsqrt(+add) %.1f  div(+add) %.1f  sqrt(+add) %.1f  FFMA %.1f\n",
         h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 1024.0, h[4] / 256.0, h[5] / 256.0, h[6] / 256.0, h[7] / 4096.0);
  for (int warps = 1; warps <= 32; warps *= 2) {
    thr_kernel<<<148, warps * 32>>>(out, cyc, 1.0000001, 0.9999999);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    double fma_per_clk_sm = (double)warps * 32 * 8 * 512 / h[0];
    printf("throughput: %2d warps/SM x ILP8: %.1f DFMA/clk/SM (%lld cycles)\n", warps, fma_per_clk_sm, h[0]);
  }
  return 0;
}
