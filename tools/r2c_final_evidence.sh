#!/bin/bash
# Final evidence of round 2 on ONE B200 (gpurun --timeout 1200 -- 'bash tools/r2c_final_evidence.sh'): the GPU test suite,
# smoke(), the contract bench line the way the driver runs it (N = 1, --steps 20 --warmup 5) with the reference arm beside
# it, the ncu launch list of the same bench command, and one ncu --set full capture of the SDE_COMPAT_FAST_RHS |
# SDE_COMPAT_FAST_STAGES kernel (config 2faster).  The other ncu captures of profiles/r2_ncu_* are of kernels whose SASS
# did not change (tools/r2_evidence.sh made them).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_gpu_tests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.json 2> /dev/null
ncu --set full --clock-control none --import-source on -k regex:fixed_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_config2faster \
    python bench.py --config 2faster --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
for r in gpurun_out/r2_ncu_config2faster.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r > $b.summary.txt 2>&1
  python tools/ncu_source_regions.py $r 0.5 > $b.regions.txt 2>&1
  python tools/ncu_opcode_mix.py $r > $b.opcodes.txt 2>&1
done
cat gpurun_out/r2_gpu_tests_final.txt gpurun_out/r2_smoke_final.txt; head -c 500 gpurun_out/r2_bench_n1_final.json; echo; head -c 300 gpurun_out/r2_bench_reference_arm.json; echo
head -20 gpurun_out/r2_ncu_config2faster.summary.txt
