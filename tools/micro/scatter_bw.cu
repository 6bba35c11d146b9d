// Microbenchmark (development aid): DRAM write efficiency of the trajectory-major series layout.
// Every warp owns 32 rows of ROW bytes (24 024 B = 1001 x 3 doubles); per "flush" it writes a contiguous
// CHUNK of each of its 32 rows (consecutive lanes -> consecutive 8-byte words), then advances.
// Compares chunk sizes and a fully coalesced SoA-style stream.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void scatter(double* out, long long n_rows, int row_elems, int chunk_elems) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long row0 = warp * 32;
  if (row0 >= n_rows) return;
  for (int s0 = 0; s0 < row_elems; s0 += chunk_elems) {
    const int cnt = min(chunk_elems, row_elems - s0);
    for (int tr = 0; tr < 32; ++tr) {
      double* p = out + (row0 + tr) * row_elems + s0;
      for (int off = lane; off < cnt; off += 32) p[off] = (double)(off + tr);
    }
  }
}
__global__ void stream(double* out, long long n_rows, int row_elems) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  for (int s = 0; s < row_elems; ++s) out[(long long)s * n_rows + i] = (double)s;
}
int main() {
  const long long n_rows = 1 << 20; const int row = 3003;
  double* d; cudaMalloc(&d, n_rows * row * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int chunks[] = {24, 45, 48, 96, 192, 384, 768, 3003};
  for (int bs : {128, 256}) for (int c : chunks) {
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(a); scatter<<<(n_rows + bs - 1) / bs, bs>>>(d, n_rows, row, c); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("scatter block=%d chunk=%4d doubles (%5d B): %.3f ms  %.0f GB/s\n", bs, c, c * 8, best, n_rows * row * 8.0 / best / 1e6);
  }
  float best = 1e9;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(a); stream<<<(n_rows + 127) / 128, 128>>>(d, n_rows, row); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  printf("stream (SoA): %.3f ms  %.0f GB/s\n", best, n_rows * row * 8.0 / best / 1e6);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
