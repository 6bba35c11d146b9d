#!/bin/bash
# Last evidence call of round 2 (one B200): GPU test suite, smoke(), the CTA-cap probe with the launcher's rule in place,
# the driver-style bench line + reference arm, compute-sanitizer over tools/sanitize.sh's workload (small-solve context included)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_gpu_tests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.txt 2>&1
python tools/cta_cap_probe.py > gpurun_out/r2c_cta_cap_probe_after.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_n1_final.err
timeout 400 bash tools/sanitize.sh > gpurun_out/r2c_sanitize.txt 2>&1
cat gpurun_out/r2_gpu_tests_final.txt gpurun_out/r2_smoke_final.txt gpurun_out/r2c_cta_cap_probe_after.txt; head -c 400 gpurun_out/r2_bench_n1_final.json; echo; cat gpurun_out/r2c_sanitize.txt
