// Development aid / evidence: accuracy of the controller's fast FP64 helpers (sde_common.cuh)
// against the CUDA math library, over the ranges the controller uses.
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../simplediffeq.jl_b200/csrc/device/sde_common.cuh"
__global__ void k(const double* x, double* out, int n) {
  __shared__ double tab[sde::kC_count];
  for (int i = threadIdx.x; i < sde::kC_count; i += blockDim.x) tab[i] = sde::k_ctrl[i];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i];
  out[3 * i + 0] = sde::sde_rcp_fast(v, 1.0) * v - 1.0;                       // relative error of 1/v
  out[3 * i + 1] = sde::sde_log2_fast(v, tab) - log2(v);                      // absolute error of log2
  double e = -3.4 + 6.8 * (i / (double)n);
  out[3 * i + 2] = sde::sde_exp2_fast(e, tab) / exp2(e) - 1.0;                // relative error of exp2
}
int main() {
  const int n = 1 << 22;
  std::vector<double> h(n);
  for (int i = 0; i < n; ++i) h[i] = pow(10.0, -30.0 + 40.0 * ((i * 2654435761u) % 1000003) / 1000003.0);  // 1e-30 .. 1e10
  double *dx, *dout; cudaMalloc(&dx, n * 8); cudaMalloc(&dout, 3 * n * 8);
  cudaMemcpy(dx, h.data(), n * 8, cudaMemcpyHostToDevice);
  k<<<(n + 255) / 256, 256>>>(dx, dout, n);
  std::vector<double> o(3 * n); cudaMemcpy(o.data(), dout, 3 * n * 8, cudaMemcpyDeviceToHost);
  double m[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) m[j] = fmax(m[j], fabs(o[3 * i + j]));
  printf("max |rel err| rcp_fast   : %.3g\nmax |abs err| log2_fast  : %.3g (|log2| up to 100)\nmax |rel err| exp2_fast  : %.3g\n%s\n", m[0], m[1], m[2], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
