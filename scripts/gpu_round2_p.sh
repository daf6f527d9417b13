#!/bin/bash
# round-2 GPU call P (1 GPU): ncu of the compact-direction kernels (launch list + --set full), m = 6 at 1e8 and m = 20 at 2^28
mkdir -p gpurun_out
python scripts/prof_compact.py > gpurun_out/p_plain.log 2>&1 || { cat gpurun_out/p_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/p_launches_compact.csv python scripts/prof_compact.py > gpurun_out/p_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gram|k_direction|k_compact_solve" -s 15 -c 6 -o gpurun_out/p_prof_compact_m6 python scripts/prof_compact.py > gpurun_out/p_ncu2.log 2>&1
PROF_N=268435456 PROF_M=20 python scripts/prof_compact.py > gpurun_out/p_plain20.log 2>&1 &&
PROF_N=268435456 PROF_M=20 ncu --set full --clock-control none --import-source on -k regex:"k_gram|k_direction_gen" -s 80 -c 5 -o gpurun_out/p_prof_compact_m20 python scripts/prof_compact.py > gpurun_out/p_ncu3.log 2>&1
ls -la gpurun_out/p_*; tail -n 3 gpurun_out/p_ncu2.log gpurun_out/p_ncu3.log; cat gpurun_out/p_plain.log gpurun_out/p_plain20.log
