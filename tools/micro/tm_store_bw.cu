// Microbenchmark (development aid): what bounds the trajectory-major series layout -- the run length, or the
// number of rows being written at the same time?  Every LANE owns one row of ROW doubles at a time (like the
// staged writer of sde_kernels.cuh) and writes it front to back in bulk copies shared -> global of `sub` doubles
// (cp.async.bulk); the grid is persistent (148 * cps CTAs), so the number of rows in flight is cps * threads * 148.
// Optional L2 eviction hints: evict_last on the copies and `applypriority evict_normal` once a run of
// `subs_per_run` copies is complete (does L2 then write the run back as one piece?).
// A dependent DFMA chain of `pace` iterations per copied slot mimics the integration between copies.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tm_store_bw tm_store_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
extern __shared__ __align__(128) unsigned char smem[];

__device__ __forceinline__ u64 policy(int kind) {
  u64 p = 0;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(dst), "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_hint(void* dst, const void* src, unsigned bytes, u64 pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(dst), "r"((unsigned)__cvta_generic_to_shared(src)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void apply_normal(const void* p) {
  asm volatile("applypriority.global.L2::evict_normal [%0], 128;" :: "l"(p) : "memory");
}

// mode 0 plain | 1 evict_last + applypriority per finished run | 2 evict_last only | 3 evict_first
template <int NBUF>
__global__ void tm_store(double* out, long long n_rows, int row_elems, int sub, int subs_per_run, int pace, int mode,
                         double* sink) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  double* my = reinterpret_cast<double*>(smem) + (long long)threadIdx.x * sub * NBUF;
  const u64 pol = policy(mode == 3 ? 2 : (mode == 0 ? 0 : 1));
  double acc = (double)threadIdx.x;
  int buf = 0;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
    double* g = out + row * row_elems;
    int k = 0;
    for (int s0 = 0; s0 < row_elems; s0 += sub, ++k) {
      const int cnt = min(sub, row_elems - s0);
      if (NBUF == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      if (NBUF == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      const int it = pace * (cnt / 3);
      for (int i = 0; i < it; ++i) acc = fma(acc, 1.0000001, 0.5);
      double* s = my + buf * sub;
      s[0] = acc;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (mode == 0) bulk(g + s0, s, (unsigned)cnt * 8u);
      else bulk_hint(g + s0, s, (unsigned)cnt * 8u, pol);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (mode == 1 && ((k + 1) % subs_per_run == 0 || s0 + sub >= row_elems)) {
        // the run that just became complete: [run0, s0 + cnt)
        const int run0 = (k / subs_per_run) * subs_per_run * sub;
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // the lines must have reached L2
        for (int e = run0; e < s0 + cnt; e += 16) apply_normal(g + e);
      }
      buf = (buf + 1) % NBUF;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 12345.678) *sink = acc;
}


// The real row: ROW = 3003 doubles = 24 024 B, so rows start at every multiple of 8 bytes mod 128.
//  pattern 0 "slots": what the staged writer does -- runs of `sub` doubles from the row start; the aligned 16-byte
//             middle as one bulk copy, the odd element in front / behind by a scalar store of the lane.
//  pattern 1 "lines": the same bytes, but a lane only ever writes WHOLE 128-byte lines of global memory (the
//             partial line at the end of a run waits for the next run); row ends by scalar stores.
__global__ void tm_real(double* out, long long n_rows, int row_elems, int sub, int pace, int pattern, double* sink) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int cap = sub + 18;                                      // doubles of staging per lane
  double* my = reinterpret_cast<double*>(smem) + (long long)threadIdx.x * cap;
  double acc = (double)threadIdx.x;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
    const long long e_row = row * row_elems;                     // element offset of the row
    double* g = out + e_row;
    long long done = 0;                                          // elements of the row already written
    for (int s0 = 0; s0 < row_elems; s0 += sub) {
      const int cnt = min(sub, row_elems - s0);
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      const int it = pace * (cnt / 3);
      for (int i = 0; i < it; ++i) acc = fma(acc, 1.0000001, 0.5);
      my[0] = acc;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const long long avail = s0 + cnt;                          // elements produced so far
      long long lo = done, hi = avail;                           // candidate range [lo, hi)
      if (pattern == 1 && avail < row_elems) {
        // keep the tail that does not fill a 128-byte line (16 doubles) of global memory
        const long long abs_hi = e_row + hi;
        hi -= (abs_hi & 15);
      }
      if (hi > lo) {
        // scalar stores up to the first 16-byte boundary (pattern 0) / 128-byte boundary is implied by `done` (pattern 1)
        long long a0 = lo;
        if ((e_row + a0) & 1) { g[a0] = acc; ++a0; }
        long long a1 = hi;
        if ((e_row + a1) & 1) { --a1; g[a1] = acc; }
        if (a1 > a0) bulk(g + a0, my + ((e_row + a0) & 1), (unsigned)(a1 - a0) * 8u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        done = hi;
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 12345.678) *sink = acc;
}

// the round-1 pattern for comparison: a warp writes one chunk of each of its 32 rows with plain coalesced stores
__global__ void coop(double* out, long long n_rows, int row_elems, int chunk) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp * 32 < n_rows; warp += nwarps) {
    const long long row0 = warp * 32;
    for (int s0 = 0; s0 < row_elems; s0 += chunk) {
      const int cnt = min(chunk, row_elems - s0);
      for (int tr = 0; tr < 32; ++tr) {
        double* p = out + (row0 + tr) * row_elems + s0;
        for (int off = lane; off < cnt; off += 32) p[off] = (double)(off + tr);
      }
    }
  }
}

int main(int argc, char** argv) {
  const long long n_rows = 1 << 20;
  const int row = 3008;                       // 24 064 B = 188 lines of 128 B (the real row is 24 024 B)
  double *d, *sink;
  cudaMalloc(&d, n_rows * row * 8);
  cudaMalloc(&sink, 8);
  cudaMemset(d, 0, n_rows * row * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int dev = 0, sms = 0, l2 = 0, maxp = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
  cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
  printf("SMs %d, L2 %d MB, max persisting L2 %d MB\n", sms, l2 >> 20, maxp >> 20);
  const double gb = n_rows * row * 8.0 / 1e9;

  auto run = [&](int threads, int cps, int nbuf, int sub, int spr, int pace, int mode) {
    const size_t sh = (size_t)threads * sub * 8 * nbuf;
    if (sh > 227 * 1024 || sh * cps > 227 * 1024) return;
    void (*kern)(double*, long long, int, int, int, int, int, double*) =
        nbuf == 1 ? tm_store<1> : nbuf == 2 ? tm_store<2> : tm_store<4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, sh);
    if (occ < cps) { printf("skip threads=%d cps=%d nbuf=%d sub=%d: occupancy %d\n", threads, cps, nbuf, sub, occ); return; }
    float best = 1e9;
    for (int r = 0; r < 2; ++r) {
      cudaEventRecord(a);
      kern<<<sms * cps, threads, sh>>>(d, n_rows, row, sub, spr, pace, mode, sink);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("tm mode=%d threads=%3d cps=%d rows_in_flight=%6d nbuf=%d sub=%4d B run=%5d B pace=%2d : %8.3f ms %6.0f GB/s %s\n",
           mode, threads, cps, sms * cps * threads, nbuf, sub * 8, sub * 8 * (mode == 1 ? spr : 1), pace, best,
           gb / best * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
  };


  // (0) the real row length (3003 doubles): slot-aligned runs (what the kernel does) against whole-line flushes
  if (argc > 1) {
    const int row3 = 3003;
    const double gb3 = n_rows * row3 * 8.0 / 1e9;
    for (int pace : {0, 20, 37})
      for (int cps : {2, 3, 4})
        for (int sub : {36, 48, 60, 96})
          for (int pattern : {0, 1}) {
            const size_t sh = (size_t)128 * (sub + 18) * 8;
            if ((sh + 1024) * cps > 227 * 1024) continue;
            cudaFuncSetAttribute(tm_real, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
            float best = 1e9;
            for (int r = 0; r < 2; ++r) {
              cudaEventRecord(a);
              tm_real<<<sms * cps, 128, sh>>>(d, n_rows, row3, sub, pace, pattern, sink);
              cudaEventRecord(b);
              cudaEventSynchronize(b);
              float ms; cudaEventElapsedTime(&ms, a, b);
              if (ms < best) best = ms;
            }
            printf("real-row %s cps=%d sub=%4d B pace=%2d : %8.3f ms %6.0f GB/s %s\n", pattern ? "lines" : "slots", cps, sub * 8,
                   pace, best, gb3 / best * 1e3, cudaGetErrorString(cudaGetLastError()));
            fflush(stdout);
          }
    return 0;
  }

  // (1) plain bulk copies: rows in flight x copy size, unpaced and paced
  for (int pace : {0, 9, 37})
    for (int threads : {32, 64, 128})
      for (int cps : {1, 2, 4})
        for (int nbuf : {1, 2})
          for (int sub : {48, 96, 192, 384, 768}) run(threads, cps, nbuf, sub, 1, pace, 0);
  // (2) L2 hints, without and with a persisting set-aside
  for (int setaside : {0, 1}) {
    if (setaside) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp);
      printf("persisting L2 set-aside %d MB: %s\n", maxp >> 20, cudaGetErrorString(e));
    }
    for (int pace : {0, 9, 37})
      for (int mode : {2, 3, 1})
        for (int cps : {1, 2, 4})
          for (int sub : {48, 96})
            for (int spr : {4, 8, 16}) {
              if (mode != 1 && spr != 4) continue;
              run(128, cps, 2, sub, spr, pace, mode);
            }
  }
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  // (3) the round-1 cooperative pattern with a persistent grid: rows in flight x chunk
  for (int cps : {1, 2, 4, 16})
    for (int chunk : {48, 192, 384, 768, 3008}) {
      float best = 1e9;
      for (int r = 0; r < 2; ++r) {
        cudaEventRecord(a);
        coop<<<sms * cps, 128>>>(d, n_rows, row, chunk);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      printf("coop cps=%2d rows_in_flight=%6d chunk=%5d B : %8.3f ms %6.0f GB/s\n", cps, sms * cps * 128, chunk * 8, best,
             gb / best * 1e3);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
