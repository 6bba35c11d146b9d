#!/bin/bash
# Persistent-CTA count sweep for the adaptive kernels (development aid, run on a B200):
# fewer CTAs per SM = fewer trajectories in flight = each warp advances faster, so the latency-bound
# drain at the end of the queue is shorter -- as long as the FP64 pipe stays saturated.
# Binaries: built beforehand into tools/micro/bin (nvcc ... adaptive_bench.cu, see its header;
# ab_vdp: -DSYS=VanDerPol, ab_avern9: -DMETHOD=Vern9Method -DV9=true, ab_avern7: -DMETHOD=Vern7Method, *_b64: -DBLOCK=64).
set -u
B="$(dirname "$0")/bin"
for lg in 18 20 22; do
  for k in 0 1 2 3; do $B/ab_lorenz 1e-8 0 $lg $k; done
  for k in 0 2 3 4 5 6; do $B/ab_vdp 1e-6 0 $lg $k; done
  for k in 0 2 3 4 5 6; do $B/ab_vdp 1e-6 1 $lg $k; done
done
for k in 0 1 2; do $B/ab_avern9 1e-12 0 20 $k; done
for k in 0 1 2; do $B/ab_avern7 1e-10 0 20 $k; done
for k in 0 2 4 6 8; do $B/ab_lorenz_b64 1e-8 0 20 $k; done
for k in 0 4 6 8 12 16; do $B/ab_vdp_b64 1e-6 0 20 $k; done
