// Microbenchmark of the L-BFGS two-loop recursion variants (development aid, not product code).
// Includes the product translation unit to time the very functions the solver calls.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out_in/two_loop_bench tools/microbench/two_loop_bench.cu dftpav_b200/csrc/dftpav_api.cu
#include "../../dftpav_b200/csrc/dftpav_kernels.cu"

#include <cstdio>
#include <vector>

using namespace dftpav;

// V2: prefetch distance 2 with explicit register double-buffers (NPL = 2 only)
__device__ __noinline__ void two_loop_pd2(double* d_sm, double* alpha_sm, const double* __restrict__ lm_s,
                                          const double* __restrict__ lm_y, const double* __restrict__ lm_ys, int ld, int m,
                                          int end, int bound_in, int n, int lane, double ys, double yy) {
  const int bound = __shfl_sync(FULL, bound_in, 0);
  const bool ok0 = lane < n, ok1 = lane + 32 < n;
  double d0 = ok0 ? d_sm[lane] : 0.0, d1 = ok1 ? d_sm[lane + 32] : 0.0;
  double s0[3], s1[3], y0[3], y1[3], r[3];
  auto load = [&](int q, int j) {
    const double* sj = lm_s + (size_t)j * ld + lane;
    const double* yj = lm_y + (size_t)j * ld + lane;
    s0[q] = ok0 ? sj[0] : 0.0, s1[q] = ok1 ? sj[32] : 0.0;
    y0[q] = ok0 ? yj[0] : 0.0, y1[q] = ok1 ? yj[32] : 0.0;
    r[q] = lm_ys[j];
  };
  int j = end - 1;
  if (j < 0) j += m;
  int jf = j;
  load(0, jf);
  jf = jf == 0 ? m - 1 : jf - 1;
  if (bound > 1) load(1, jf);
  jf = jf == 0 ? m - 1 : jf - 1;
  for (int i = 0; i < bound; i += 3) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
      if (i + q < bound) {
        if (i + q + 2 < bound) {
          load((q + 2) % 3, jf);
          jf = jf == 0 ? m - 1 : jf - 1;
        }
        const double a = warp_sum(s0[q] * d0 + s1[q] * d1) * r[q];
        if (lane == 0) alpha_sm[j] = a;
        d0 -= a * y0[q], d1 -= a * y1[q];
        if (i + q + 1 < bound) j = j == 0 ? m - 1 : j - 1;
      }
    }
  }
  const double scl = ys / yy;
  d0 *= scl, d1 *= scl;
  __syncwarp();
  jf = j;
  load(0, jf);
  jf = jf + 1 == m ? 0 : jf + 1;
  if (bound > 1) load(1, jf);
  jf = jf + 1 == m ? 0 : jf + 1;
  for (int i = 0; i < bound; i += 3) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
      if (i + q < bound) {
        if (i + q + 2 < bound) {
          load((q + 2) % 3, jf);
          jf = jf + 1 == m ? 0 : jf + 1;
        }
        const double beta = warp_sum(y0[q] * d0 + y1[q] * d1) * r[q];
        const double a = alpha_sm[j] - beta;
        d0 += a * s0[q], d1 += a * s1[q];
        j = j + 1 == m ? 0 : j + 1;
      }
    }
  }
  if (ok0) d_sm[lane] = d0;
  if (ok1) d_sm[lane + 32] = d1;
  __syncwarp();
}

// V4: 16 lanes x 3 elements (n <= 48): 4-level butterfly inside the half warp; both halves compute the same
__device__ __noinline__ void two_loop_h16(double* d_sm, double* alpha_sm, const double* __restrict__ lm_s,
                                          const double* __restrict__ lm_y, const double* __restrict__ lm_ys, int ld, int m,
                                          int end, int bound_in, int n, int lane, double ys, double yy) {
  const int bound = __shfl_sync(FULL, bound_in, 0);
  const int l = lane & 15;
  bool ok[3];
  double d[3], sc[3], yc[3], sn[3], yn[3], ysc, ysn = 0.0;
#pragma unroll
  for (int e = 0; e < 3; e++) {
    ok[e] = l + 16 * e < n;
    d[e] = ok[e] ? d_sm[l + 16 * e] : 0.0;
    sn[e] = yn[e] = 0.0;
  }
  auto hsum = [&](double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  };
  int j = end - 1;
  if (j < 0) j += m;
  auto ld3 = [&](double* s_, double* y_, double& r_, int jj) {
    const double* sj = lm_s + (size_t)jj * ld + l;
    const double* yj = lm_y + (size_t)jj * ld + l;
#pragma unroll
    for (int e = 0; e < 3; e++) s_[e] = ok[e] ? sj[16 * e] : 0.0, y_[e] = ok[e] ? yj[16 * e] : 0.0;
    r_ = lm_ys[jj];
  };
  ld3(sc, yc, ysc, j);
#pragma unroll 1
  for (int i = 0; i < bound; ++i) {
    int jn = j - 1;
    if (jn < 0) jn += m;
    if (i + 1 < bound) ld3(sn, yn, ysn, jn);
    const double a = hsum(sc[0] * d[0] + sc[1] * d[1] + sc[2] * d[2]) * ysc;
    if (lane == 0) alpha_sm[j] = a;
#pragma unroll
    for (int e = 0; e < 3; e++) d[e] -= a * yc[e], sc[e] = sn[e], yc[e] = yn[e];
    ysc = ysn;
    if (i + 1 < bound) j = jn;
  }
  const double scl = ys / yy;
#pragma unroll
  for (int e = 0; e < 3; e++) d[e] *= scl;
  __syncwarp();
  ld3(sc, yc, ysc, j);
#pragma unroll 1
  for (int i = 0; i < bound; ++i) {
    int jn = j + 1;
    if (jn >= m) jn -= m;
    if (i + 1 < bound) ld3(sn, yn, ysn, jn);
    const double beta = hsum(yc[0] * d[0] + yc[1] * d[1] + yc[2] * d[2]) * ysc;
    const double a = alpha_sm[j] - beta;
#pragma unroll
    for (int e = 0; e < 3; e++) d[e] += a * sc[e], sc[e] = sn[e], yc[e] = yn[e];
    ysc = ysn;
    j = jn;
  }
  if (lane < 16) {
#pragma unroll
    for (int e = 0; e < 3; e++)
      if (ok[e]) d_sm[l + 16 * e] = d[e];
  }
  __syncwarp();
}

template <int V>
__global__ void __launch_bounds__(384, 1) bench_kernel(double* hist_s, double* hist_y, double* hist_ys, double* gram, int ld,
                                                       int m, int bound, int n, int reps, int smem_hist,
                                                       unsigned long long* cycles, double* sink) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int slot = blockIdx.x * nw + warp;
  double* d_sm = sm + warp * 600;
  double* alpha = d_sm + 64;
  double* lm_s = hist_s + (size_t)slot * m * ld;
  double* lm_y = hist_y + (size_t)slot * m * ld;
  double* lm_ys = hist_ys + (size_t)slot * m;
  (void)gram;
  if (smem_hist) {  // chain-only variant: the (small) history lives in shared memory
    double* hs = sm + nw * 600 + warp * (2 * smem_hist * ld + smem_hist);
    for (int i = lane; i < smem_hist * ld; i += 32) hs[i] = lm_s[i], hs[smem_hist * ld + i] = lm_y[i];
    for (int i = lane; i < smem_hist; i += 32) hs[2 * smem_hist * ld + i] = lm_ys[i];
    lm_s = hs, lm_y = hs + smem_hist * ld, lm_ys = hs + 2 * smem_hist * ld;
    m = smem_hist;
  }
  for (int i = lane; i < 64; i += 32) d_sm[i] = 1.0 / (1 + i);
  __syncwarp();
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    const int end = bound % m;
    if (V == 0) two_loop<2>(d_sm, alpha, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, 1.0, 1.0);
    if (V == 2) two_loop_pd2(d_sm, alpha, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, 1.0, 1.0);
    if (V == 4) two_loop_h16(d_sm, alpha, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, 1.0, 1.0);
    if (V == 5) two_loop_generic(d_sm, alpha, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, 1.0, 1.0);
  }
  long long t1 = clock64();
  if (lane == 0) {
    atomicAdd(cycles, (unsigned long long)(t1 - t0));
    sink[slot] = d_sm[3];
  }
}

int main() {
  const int n = 47, ld = 48, m = 256, maxslots = 148 * 12;
  std::vector<double> hs((size_t)maxslots * m * ld), hy(hs.size()), hys((size_t)maxslots * m);
  for (size_t i = 0; i < hs.size(); i++) hs[i] = 1e-3 * ((i * 2654435761u) % 1000) / 1000.0, hy[i] = 1e-3 * ((i * 40503u) % 1000) / 1000.0;
  for (size_t i = 0; i < hys.size(); i++) hys[i] = 1.0 + (i % 7);
  double *ds, *dy, *dys, *dg, *sink;
  unsigned long long* cyc;
  cudaMalloc(&ds, hs.size() * 8), cudaMalloc(&dy, hs.size() * 8), cudaMalloc(&dys, hys.size() * 8);
  cudaMalloc(&dg, hys.size() * 32), cudaMalloc(&sink, maxslots * 8), cudaMalloc(&cyc, 8);
  cudaMemcpy(ds, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dy, hy.data(), hs.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dys, hys.data(), hys.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(dg, 0, hys.size() * 32);
  const char* names[] = {"two_loop<2> (product)", "", "prefetch distance 2", "", "16 lanes x 3", "generic (smem d)"};
  auto run = [&](int V, int grid, int warps, int bound, int smem_hist, bool small_l1 = false) {
    const int reps = 20;
    size_t smem = (size_t)warps * 600 * 8 + (smem_hist ? (size_t)warps * (2 * smem_hist * ld + smem_hist) * 8 : 0);
    if (small_l1) smem = 185 * 1024;  // the solver's carve-out: ~40 KB of L1 left
    cudaMemset(cyc, 0, 8);
    auto launch = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, warps * 32, smem>>>(ds, dy, dys, dg, ld, m, bound, n, 2, smem_hist, cyc, sink);  // warm-up
      cudaMemset(cyc, 0, 8);
      kern<<<grid, warps * 32, smem>>>(ds, dy, dys, dg, ld, m, bound, n, reps, smem_hist, cyc, sink);
    };
    if (V == 0) launch(bench_kernel<0>);
    if (V == 2) launch(bench_kernel<2>);
    if (V == 4) launch(bench_kernel<4>);
    if (V == 5) launch(bench_kernel<5>);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_step = (double)c / ((double)grid * warps * reps * 2.0 * bound);
    printf("%-24s grid %3d warps %2d bound %3d %s: %7.1f cycles per history step%s\n", names[V], grid, warps, bound,
           smem_hist ? "smem-history" : (small_l1 ? "global,L1~40K" : "global       "), per_step, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  for (int V : {0, 2}) {
    for (int bound : {16, 64, 128}) run(V, 1, 1, bound, 0);       // a lone warp on the GPU
    run(V, 1, 1, 16, 16);                                         // chain only: history in shared memory
    run(V, 1, 1, 128, 0, true);                                   // a lone warp, L1 as small as in the solver
    run(V, 148, 12, 64, 0, true);
    run(V, 148, 12, 64, 0);                                       // every SM full of two-loops
    run(V, 148, 4, 64, 0);
  }
  return 0;
}
