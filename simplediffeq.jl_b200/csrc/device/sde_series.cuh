// sde_series.cuh -- series-output writer of the fixed-step kernels (sde_kernels.cuh: fixed_body) and of the
// SimpleEM kernels (sde_em.cuh: em_body): direct stores in either layout, or trajectory-major rows staged in
// shared memory and written in whole 128-byte lines.  Self-contained for NVRTC.
#pragma once
#include "sde_common.cuh"

namespace sde {

// ------------------------------------------------------------------------------------------
// series output writer.
//   STAGED = false: direct stores in the layout the launch asked for.  Coalesced for kLayoutSoA
//            (consecutive lanes -> consecutive addresses).
//   STAGED = true : kLayoutTrajMajor, fixed step.  Each trajectory owns a contiguous row
//            out_u[traj][slot][c]; a thread storing its own N values per save point would write 32
//            different rows per instruction, 8 bytes each.  Instead every LANE stages its row's byte stream in
//            shared memory, and the warp writes it out in WHOLE 128-BYTE LINES of global memory:
//              * a lane's staging region is congruent with global memory mod 128 bytes (region byte 0 = the line
//                that holds the first unwritten byte of the row), so a line of the row is a line of the region;
//              * as soon as K0 lines are complete in every lane (a warp-uniform count: fixed-step lanes write in
//                lockstep) the warp copies them row after row -- lane q moves the q-th 16-byte piece (LDS.128 ->
//                STG.128; with K0 = 2 a pass covers two rows) -- and every lane slides what is left of its region
//                (less than 128 bytes + one slot) to the front;
//              * only the first line of a row (shared with the previous row) and the bytes left at the end are
//                written by their owner lane with scalar stores.
//            Why lines: rows start at every multiple of 8 bytes mod 128 (24 024-byte rows), and runs cut at slot
//            boundaries leave a partially written sector / line at both ends of every run, which L2 hands to DRAM
//            separately.  Pure-store microbenchmark with the real row length (tools/micro/tm_store_bw.cu,
//            profiles/r2_tm_store_microbench.txt): slot-aligned 384-byte runs 3.1-3.4 TB/s, whole lines 4.5-4.7 TB/s.
//            Shared memory per warp: 32 lanes x LS elements of stage + kRingBytes for the dense-output
//            weights of the step (fixed_body).  Every lane of the warp must stay alive (lanes without a
//            trajectory compute a copy of the last one and never write).
// ------------------------------------------------------------------------------------------
template <class T, int N>
struct StageCfg {
#ifndef SDE_STAGE_ELEMS_F64
#define SDE_STAGE_ELEMS_F64 48    // staging capacity per lane in elements (+ 16 bytes): 400 B, 4 CTAs of 4 warps per SM
#define SDE_STAGE_ELEMS_F32 96
#endif
  static constexpr int kSz = (int)sizeof(T);
  static constexpr int kLineE = 128 / kSz;                      // elements per 128-byte line
  static constexpr int kElems = (kSz == 8 ? SDE_STAGE_ELEMS_F64 : SDE_STAGE_ELEMS_F32);
  // worst case before a flush: (128 - sz) bytes of line offset + K0 lines + one slot that just crossed the boundary
  static constexpr int kNeed = (128 - kSz) + 128 + (N * kSz - kSz);
  static constexpr int kWant = kElems * kSz + 16;
  static constexpr int kRawB = ((kWant > kNeed ? kWant : kNeed) + 15) / 16 * 16;
  static constexpr int kCapB = ((kRawB / 16) % 2 == 0) ? kRawB + 16 : kRawB;   // lane stride: an odd number of 16-byte units
  static constexpr int LS = kCapB / kSz;
  static constexpr int K0 = (kCapB - (128 - kSz) - (N * kSz - kSz)) / 128;     // lines per flush (>= 1)
  static constexpr int kRingBytes = 1024;                               // dense-output weights of a step, per warp
  static constexpr int kRingElems = kRingBytes / kSz;
  static constexpr int kBytesPerWarp = 32 * kCapB + kRingBytes;
};

extern __shared__ __align__(16) unsigned char sde_dyn_smem[];

template <class T, int N, bool STAGED>
struct SeriesWriter {
  using Cfg = StageCfg<T, N>;
  const KArgs<T>& a;
  i64 traj;
  bool valid;
  i64 slot;      // next slot to be written by put()
  // staged writer
  unsigned char* wstage;   // the warp's staging region: lane l owns bytes [l * kCapB, (l + 1) * kCapB)
  T* buf;        // this lane's region
  int wpos;      // element index in buf of the next value            (lane specific: includes the row's line offset)
  int wb;        // elements staged beyond the flushed lines by a row that starts on a line boundary (warp uniform)
  i64 lines;     // lines of the rows written so far                  (warp uniform)
  char* gline;   // global address of the line that holds the start of this lane's row
  unsigned lane;
  u64 gbase0;    // byte address of the warp's first row               (warp uniform)
  int nrows;     // trajectories of the warp that exist                (warp uniform)
  // direct stores
  T* q0;         // component 0 of the next slot
  i64 cs, ss;    // element strides between components / between slots (warp uniform)

  __device__ __forceinline__ SeriesWriter(const KArgs<T>& a_, i64 traj_, bool valid_)
      : a(a_), traj(traj_), valid(valid_), slot(0), wstage(nullptr), buf(nullptr), wpos(0), wb(0), lines(0),
        gline(nullptr), lane(0), gbase0(0), nrows(0), q0(nullptr), cs(0), ss(0) {
    if (!STAGED) {
      if (a.layout == kLayoutTrajMajor) {      // out_u[(traj * n_out + slot) * N + c]
        q0 = a.out_u + traj * a.n_out * N; cs = 1; ss = N;
      } else {                                 // out_u[(slot * N + c) * ld_out + traj]
        q0 = a.out_u + traj; cs = a.ld_out; ss = N * a.ld_out;
      }
    }
    if (STAGED) {
      lane = threadIdx.x & 31u;
      // the warp's index as a value ptxas knows to be warp-uniform
      const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
      const i64 traj0 = (i64)blockIdx.x * blockDim.x + (i64)warp * 32;
      const i64 left = a.n_traj - traj0;
      const int here = (int)blockDim.x - warp * 32;            // (host emulation: one-lane blocks)
      nrows = (int)(left < 32 ? (left < 0 ? 0 : left) : 32);
      if (nrows > here) nrows = here;
      wstage = sde_dyn_smem + (size_t)warp * Cfg::kBytesPerWarp;
      buf = reinterpret_cast<T*>(wstage + (size_t)lane * Cfg::kCapB);
      const i64 rowbytes = a.n_out * N * (i64)sizeof(T);
      gbase0 = (u64)a.out_u + (u64)(traj0 * rowbytes);
      const u64 mine = (u64)a.out_u + (u64)(traj * rowbytes);
      gline = reinterpret_cast<char*>(mine & ~(u64)127);
      wpos = (int)(mine & (u64)127) / Cfg::kSz;
    }
  }
  // the warp's ring for the dense-output weights of a step (behind the stage, 16-byte aligned)
  __device__ __forceinline__ T* ring() const { return reinterpret_cast<T*>(wstage + 32 * Cfg::kCapB); }

  // the warp writes lines [first, K) of every lane's region to the rows (all arguments warp uniform)
  __device__ __forceinline__ void copy_lines(int first, int K) {
    constexpr int P = Cfg::K0 * 8;                       // 16-byte pieces per row in the common case
    constexpr int RP = 32 / P > 0 ? 32 / P : 1;          // rows per pass of the warp
    const i64 rowbytes = a.n_out * N * (i64)sizeof(T);
    const u64 foff = (u64)lines * 128u;
    if (first == 0 && K == Cfg::K0 && P <= 32 && nrows == 32) {
      // common case: K0 whole lines per row, RP rows per pass, four passes in flight
      typedef typename Vec16<T>::type V16;
      const unsigned rr = lane / P, q = lane % P;        // (P is a power of two or 24: folded at compile time)
      const bool act = rr < (unsigned)RP;
      const unsigned char* sp = wstage + rr * Cfg::kCapB + q * 16u;
      u64 brow = gbase0 + (u64)rr * (u64)rowbytes;
      const u64 bstep = (u64)RP * (u64)rowbytes;
      constexpr int kBatch = 4;
      static_assert((32 / RP) % kBatch == 0, "passes per flush must be a multiple of the batch");
      for (int r0 = 0; r0 < 32; r0 += RP * kBatch) {
        V16 v[kBatch];
#pragma unroll
        for (int i = 0; i < kBatch; ++i)
          if (act) v[i] = load16<T>(sp + (size_t)(r0 + i * RP) * Cfg::kCapB);
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
          if (act) store16<T>(reinterpret_cast<char*>((brow & ~(u64)127) + foff + q * 16u), v[i]);
          brow += bstep;
        }
      }
    } else {
      u64 brow = gbase0;
      const unsigned char* sp = wstage;
      for (int r = 0; r < nrows; ++r) {
        char* g = reinterpret_cast<char*>((brow & ~(u64)127) + foff);
        for (unsigned q = (unsigned)first * 128u + lane * 16u; q < (unsigned)K * 128u; q += 512u) copy16<T>(g + q, sp + q);
        brow += (u64)rowbytes;
        sp += Cfg::kCapB;
      }
    }
  }

  // all lanes hold at least K = wb / kLineE complete lines: write them, keep the rest
  __device__ __forceinline__ void flush_lines() {
    const int K = wb / Cfg::kLineE;
    __syncwarp();
    if (lines == 0) {
      // line 0 is shared with the previous row: the owner writes its part
      if (valid) {
        const int e0 = wpos - wb;                      // the row's offset in its first line (wpos = offset + wb throughout)
        T* g = reinterpret_cast<T*>(gline);
        for (int e = e0; e < Cfg::kLineE; ++e) g[e] = buf[e];
      }
      copy_lines(1, K);
    } else {
      copy_lines(0, K);
    }
    __syncwarp();
    // slide the unwritten rest of the region to the front (16-byte pieces; at most one line + one slot)
    const int keep = wpos - K * Cfg::kLineE;                        // elements (lane specific)
    for (int e = 0; e < keep; e += 16 / Cfg::kSz) copy16<T>(buf + e, buf + K * Cfg::kLineE + e);
    wpos = keep;
    wb -= K * Cfg::kLineE;
    lines += K;
  }

  // fixed-step kernels call put() with the same slot in every lane
  __device__ __forceinline__ void put(const T* v) {
    if (STAGED) {
      T* my = buf + wpos;
#pragma unroll
      for (int c = 0; c < N; ++c) my[c] = v[c];
      wpos += N;
      wb += N;
      ++slot;
      if (wb >= Cfg::K0 * Cfg::kLineE) flush_lines();
    } else {
      // one running pointer; the layout only enters through the two strides (no branch, no re-derived
      // (slot * N + c) * ld_out + traj per save point: 9 instead of 19 instructions for the three Lorenz stores)
      if (valid) {
#pragma unroll
        for (int c = 0; c < N; ++c) q0[c * cs] = v[c];
      }
      q0 += ss;
      ++slot;
    }
  }

  // the end of the rows: what is still staged goes out by scalar stores of the owner lanes
  __device__ __forceinline__ void finish() {
    if (STAGED) {
      if (valid) {
        T* g = reinterpret_cast<T*>(gline + lines * 128);
        const int e0 = (lines == 0) ? (wpos - wb) : 0;
        for (int e = e0; e < wpos; ++e) g[e] = buf[e];
      }
    }
  }
};

}  // namespace sde
