#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun --timeout 1500 -- 'bash tools/r2_evidence.sh'):
#  1. the whole GPU test suite
#  2. the contract bench line the way the driver runs it (N = 1, --steps 20 --warmup 5), reference arm beside it
#  3. ncu launch list of the same bench command (every launch with its device time; shares, not absolutes)
#  4. ncu --set full of the headline kernel (config 2) and of the HBM-bound saveat kernels (config 5 SoA / trajectory-major)
#  5. ncu source-level captures of the adaptive kernels (instruction mix per attempt)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_gpu_tests_final.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.json 2> /dev/null
ncu --set full --clock-control none --import-source on -k regex:fixed_kernel -s 3 -c 1 -o gpurun_out/r2_ncu_bench_tsit5_10m \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fixed_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_config5_soa \
    python tools/prof_saveat.py 1 4000000 0.1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fixed_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_config5_tm \
    python tools/prof_saveat.py 0 4000000 0.1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fixed_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_config5f_soa \
    python tools/prof_saveat.py 1 1048576 0.01 > /dev/null 2>&1
for w in atsit5 vdp; do
  ncu --set full --clock-control none --import-source on -k regex:adaptive_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_${w}_log2 \
      python tools/prof_adaptive.py $w 0 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:adaptive_kernel -s 1 -c 1 -o gpurun_out/r2_ncu_avern9_literal_final \
    python tools/prof_adaptive.py avern9 0 > /dev/null 2>&1
# summaries are made here (ncu reads its own reports without a GPU too); only two reports travel back (64 MiB cap)
for r in gpurun_out/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r > $b.summary.txt 2>&1
  python tools/ncu_source_regions.py $r 0.5 > $b.regions.txt 2>&1
  python tools/ncu_opcode_mix.py $r > $b.opcodes.txt 2>&1
done
mkdir -p gpurun_out/keep && mv gpurun_out/r2_ncu_bench_tsit5_10m.ncu-rep gpurun_out/r2_ncu_config5_tm.ncu-rep gpurun_out/keep/
rm -f gpurun_out/*.ncu-rep && mv gpurun_out/keep/* gpurun_out/ && rmdir gpurun_out/keep
cat gpurun_out/r2_gpu_tests_final.txt; head -c 600 gpurun_out/r2_bench_n1_final.json; echo; ls -la gpurun_out | tail -40
