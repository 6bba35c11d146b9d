#!/bin/bash
# A/B of the attempt tail: _bf = accept / reject with selects (no branch), _op = branch on accept taken late
# (after the controller code both outcomes share), _fin = finish predicate instead of ret >= 0, plain = round-1 code.
set -u
B="$(dirname "$0")/bin"
for rep in 1 2; do
for v in ab_lorenz ab_lorenz_fin ab_lorenz_finop; do $B/$v 1e-8 0 20 0; done
for v in ab_vdp ab_vdp_fin ab_vdp_finop ab_vdp_finbf; do $B/$v 1e-6 0 20 0; $B/$v 1e-6 1 20 0; done
for v in ab_avern9 ab_avern9_fin ab_avern9_finop ab_avern9_finbf; do $B/$v 1e-12 0 20 0; done
for v in ab_avern7 ab_avern7_fin ab_avern7_finop; do $B/$v 1e-10 0 20 0; done
done
for v in ab_lorenz ab_lorenz_fin ab_lorenz_finop; do $B/$v 1e-8 0 22 0; done
for v in ab_vdp ab_vdp_fin ab_vdp_finop; do $B/$v 1e-6 0 22 0; done
