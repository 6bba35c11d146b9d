#!/bin/bash
# production register allocation (__launch_bounds__(128), = MINB 0) vs explicit minimum CTAs per SM, production header
set -u
B="$(dirname "$0")/bin"
for rep in 1 2; do
for v in ab_lorenz_p0 ab_lorenz_p5; do $B/$v 1e-8 0 20 0; done
for v in ab_vdp_p0 ab_vdp_p4 ab_vdp_p5; do $B/$v 1e-6 0 20 0; $B/$v 1e-6 1 20 0; done
for v in ab_avern9_p0 ab_avern9_new; do $B/$v 1e-12 0 20 0; done
for v in ab_avern7_p0 ab_avern7_new; do $B/$v 1e-10 0 20 0; done
done
