#!/bin/bash
# __launch_bounds__(128, MINB) sweep for the adaptive kernels (development aid, run on a B200; binaries prebuilt
# into tools/micro/bin: adaptive_bench.cu with -DMINB=n).  More resident warps vs fewer registers.
set -u
B="$(dirname "$0")/bin"
for v in ab_lorenz ab_lorenz_m5 ab_lorenz_m6; do $B/$v 1e-8 0 20 0; $B/$v 1e-8 0 22 0; done
for v in ab_vdp ab_vdp_m6 ab_vdp_m7; do $B/$v 1e-6 0 20 0; $B/$v 1e-6 1 20 0; $B/$v 1e-6 0 22 0; done
for v in ab_avern9 ab_avern9_m4; do $B/$v 1e-12 0 20 0; done
for v in ab_avern7 ab_avern7_m4; do $B/$v 1e-10 0 20 0; done
