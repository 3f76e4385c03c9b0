#!/bin/bash
# One-GPU evidence run (through gpurun): ncu launch list of one training step + ncu --set full of each hot kernel.
# Numbers printed by anything under ncu are never bench values.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_step.csv python tools/profile_step.py --ncu > gpurun_out/launches_step.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_step.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_kernels \
  -k regex:"conv_igemm|conv_wgrad|bn_|upsample_ce|psamask|ppm_|wgrad_reduce" python tools/ncu_kernels.py > gpurun_out/ncu_kernels.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
rm -f gpurun_out/prof_kernels.ncu-rep gpurun_out/prof_n64.ncu-rep   # only the CSV exports travel back (64 MiB cap)
ls -la gpurun_out/ | head -20
