#!/bin/bash
# round-2 GPU call AA (1 GPU): ncu --set full of the multi-step probe and of the commit fused with pass A (compact mode, n = 1e8, m = 6)
mkdir -p gpurun_out
python scripts/prof_compact.py > gpurun_out/aa_plain.log 2>&1 || { cat gpurun_out/aa_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"probe_multi|commit_gram" -s 8 -c 4 -o gpurun_out/aa_prof_probe_multi_commit_gram python scripts/prof_compact.py > gpurun_out/aa_ncu.log 2>&1
tail -n 3 gpurun_out/aa_ncu.log; cat gpurun_out/aa_plain.log; ls -la gpurun_out/aa_*
