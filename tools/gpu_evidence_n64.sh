#!/bin/bash
# ncu --set full of the first BLOCK_N=64 conv launches of one training step (stem convs), with source-level stalls.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_n64 --profile-from-start off \
  --kernel-name-base demangled -k regex:"conv_igemm_kernel<\(int\)64" -c 3 python tools/profile_step.py --ncu > gpurun_out/ncu_n64.log 2>&1
echo "n64 capture rc=$?"
ncu -i gpurun_out/prof_n64.ncu-rep --page raw --csv > gpurun_out/prof_n64_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_n64.ncu-rep --page source --csv > gpurun_out/prof_n64_source.csv 2>/dev/null
rm -f gpurun_out/prof_kernels.ncu-rep gpurun_out/prof_n64.ncu-rep   # only the CSV exports travel back (64 MiB cap)
ls -la gpurun_out | grep n64
