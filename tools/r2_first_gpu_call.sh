#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU minutes were spent.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_call.sh'
# 1. the new device-side assertion alone (so that its result is on record even if something else stops -x)
# 2. the whole GPU suite   3. literal-vs-log2 controller A/B   4. the contract bench line
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_strict_bitexact.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_strict_bitexact.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_gpu_tests.txt
python tools/strict_ab.py > gpurun_out/r2_strict_ab.txt 2>&1
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -5 gpurun_out/r2_strict_bitexact.txt gpurun_out/r2_gpu_tests.txt; cat gpurun_out/r2_strict_ab.txt; cat gpurun_out/r2_bench.json
