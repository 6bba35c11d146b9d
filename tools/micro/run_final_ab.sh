#!/bin/bash
# round-1 kernels (bin/ab_X, built before the change) vs the production header (bin/ab_X_new)
set -u
B="$(dirname "$0")/bin"
for rep in 1 2; do
for v in ab_lorenz ab_lorenz_new; do $B/$v 1e-8 0 20 0; done
for v in ab_vdp ab_vdp_new; do $B/$v 1e-6 0 20 0; $B/$v 1e-6 1 20 0; done
for v in ab_avern9 ab_avern9_new; do $B/$v 1e-12 0 20 0; done
for v in ab_avern7 ab_avern7_new; do $B/$v 1e-10 0 20 0; done
done
