// Development aid: times the GBM Euler-Maruyama kernel (Philox noise, endpoint) directly, for occupancy experiments.
// nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a [-DMINB=n] [-DTYPE=float]
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../simplediffeq.jl_b200/csrc/device/sde_em.cuh"
#ifndef MINB
#define MINB 1
#endif
#ifndef TYPE
#define TYPE double
#endif
typedef TYPE real;
__global__ void __launch_bounds__(128, MINB) kern(const __grid_constant__ sde::EMArgs<real> a) {
  sde::em_body<sde::EmGBM, real, sde::kSaveEndpoint, sde::kNoisePhilox>(a);
}
int main() {
  const long long n = 1 << 22, steps = 1000;
  std::vector<real> u0(n, 1), p(2 * n);
  for (long long i = 0; i < n; ++i) { p[i] = (real)0.1; p[n + i] = (real)0.2; }
  real *du0, *dp, *dout;
  cudaMalloc(&du0, n * sizeof(real)); cudaMalloc(&dp, 2 * n * sizeof(real)); cudaMalloc(&dout, n * sizeof(real));
  cudaMemcpy(du0, u0.data(), n * sizeof(real), cudaMemcpyHostToDevice); cudaMemcpy(dp, p.data(), 2 * n * sizeof(real), cudaMemcpyHostToDevice);
  sde::EMArgs<real> a; memset(&a, 0, sizeof a);
  a.u0 = du0; a.p = dp; a.n_traj = n; a.ld_in = n; a.t0 = 0; a.dt = (real)1e-3; a.n_steps = steps; a.out_u = dout; a.ld_out = n; a.seed = 1;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); kern<<<(unsigned)(n / 128), 128>>>(a); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  printf("MINB=%d regs=%d blocks/SM=%d: %.3f ms  steps/s=%.4g  %s\n", MINB, fa.numRegs, per_sm, best, (double)n * steps / best * 1e3,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
