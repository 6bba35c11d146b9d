#!/bin/bash
# GPU call for the SDE_COMPAT_FAST_STAGES kernels (gpurun --timeout 1200 -- 'bash tools/r2c_fast_stages.sh'):
# whole GPU test suite, the config-2 bench line with FAST_RHS and with FAST_RHS | FAST_STAGES, the measured deviation
# from the oracle, compute-sanitizer over the workload of tools/sanitize.sh (now with SimpleEM's staged rows)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c_gpu_tests.txt
cat gpurun_out/r2c_gpu_tests.txt
for c in 2faster 2fast; do
  python bench.py --config $c --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2c_bench_$c.json 2> gpurun_out/r2c_bench_$c.err
  head -c 400 gpurun_out/r2c_bench_$c.json; echo
done
for f in 24 16; do python tools/fast_rhs_diff.py $f; done > gpurun_out/r2c_fast_diff.txt 2>&1
cat gpurun_out/r2c_fast_diff.txt
timeout 700 bash tools/sanitize.sh > gpurun_out/r2c_sanitize.txt 2>&1
cat gpurun_out/r2c_sanitize.txt
