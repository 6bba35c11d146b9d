// Development aid: times one adaptive kernel instantiation directly (no library rebuild needed), for
// tuning experiments.  nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a
//   [-DMINB=n] [-DSYS=VanDerPol -DNSTATE=2 -DNPAR=1] [-DMETHOD=Vern9Method -DV9=true]
// run: ./a.out [tol] [sorted: 0|1 (Van der Pol)] [log2 n] [max persistent CTAs per SM, 0 = occupancy]
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../simplediffeq.jl_b200/csrc/device/sde_kernels.cuh"
#include "../../simplediffeq.jl_b200/csrc/device/sde_systems.cuh"
#ifndef SYS
#define SYS Lorenz
#endif
#ifndef METHOD
#define METHOD Tsit5Method
#endif
#ifndef V9
#define V9 false
#endif
#ifndef MINB
#define MINB 1
#endif
#ifndef BLOCK
#define BLOCK 128
#endif
using Sys = sde::SYS;
__global__ void __launch_bounds__(BLOCK, MINB) kern(const __grid_constant__ sde::KArgs<double> a) {
  sde::adaptive_body<Sys, double, sde::METHOD<Sys, double>, sde::kSaveEndpoint, V9, false>(a);
}
int main(int argc, char** argv) {
  const long long n = 1LL << (argc > 3 ? atoi(argv[3]) : 20);
  const int N = Sys::N, NP = Sys::NP;
  std::vector<double> u0((size_t)N * n, 0.0), p((size_t)NP * n);
  double tf = 10.0, tol = argc > 1 ? atof(argv[1]) : 1e-8;
  if (N == 3) { for (long long i = 0; i < n; ++i) { u0[i] = 1; p[i] = 10; p[n + i] = 21.0 * i / (n - 1); p[2 * n + i] = 8.0 / 3.0; } }
  else { tf = 20.0; const bool sorted = argc > 2 && atoi(argv[2]) == 1; for (long long i = 0; i < n; ++i) { u0[i] = 2; long long j = sorted ? i : (i * 2654435761LL) % n; p[i] = 0.1 + 49.9 * j / (n - 1); } }
  double *du0, *dp, *dout; int *na, *nr; unsigned long long* q;
  cudaMalloc(&du0, u0.size() * 8); cudaMalloc(&dp, p.size() * 8); cudaMalloc(&dout, u0.size() * 8);
  cudaMalloc(&na, n * 4); cudaMalloc(&nr, n * 4); cudaMalloc(&q, 8);
  cudaMemcpy(du0, u0.data(), u0.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dp, p.data(), p.size() * 8, cudaMemcpyHostToDevice);
  sde::KArgs<double> a; memset(&a, 0, sizeof a);
  a.u0 = du0; a.p = dp; a.n_traj = n; a.ld_in = n; a.t0 = 0; a.tf = tf; a.dt = (double)0.1f; a.abstol = tol; a.reltol = tol;
  a.out_u = dout; a.ld_out = n; a.n_out = 1; a.naccept = na; a.nreject = nr; a.queue = q;
  int per_sm = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, 0);
  const int occ = per_sm;
  if (argc > 4 && atoi(argv[4]) > 0 && atoi(argv[4]) < per_sm) per_sm = atoi(argv[4]);   // cap persistent CTAs per SM
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaMemset(q, 0, 8);
    cudaEventRecord(e0); kern<<<sms * per_sm, BLOCK>>>(a); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  std::vector<int> hna(n), hnr(n); cudaMemcpy(hna.data(), na, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hnr.data(), nr, n * 4, cudaMemcpyDeviceToHost);
  double acc = 0, rej = 0; for (long long i = 0; i < n; ++i) { acc += hna[i]; rej += hnr[i]; }
  printf("%s tol=%g n=%lld regs=%d occupancy=%d CTAs/SM=%d block=%d: %.3f ms  attempts/s=%.4g (acc %.0f rej %.0f) %s\n", argv[0], tol, n, fa.numRegs, occ, per_sm, BLOCK, best,
         (acc + rej) / best * 1e3, acc, rej, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
