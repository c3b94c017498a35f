#!/bin/bash
# ncu --set full captures of the dominant kernels inside one training step (bench.py --ncu-step); raw CSV pages are
# produced on the box (the reports themselves are large: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; cnt=$4; src=$5
  timeout -s KILL 600 ncu --profile-from-start off --set full --clock-control none $src \
    -k regex:$regex -s $skip -c $cnt -o gpurun_out/$name -f python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ls -la gpurun_out/$name.ncu-rep; }
cap tc3_c64f pair_tc3_kernel 0 2 "--import-source on"
cap tc3_c128f pair_tc3_kernel 10 2 "--import-source on"
cap tc3_c128b pair_tc3_kernel 34 2 ""
cap tc3_c64b pair_tc3_kernel 58 2 ""
cap wgrad_tc wgrad_tc_kernel 0 4 "--import-source on"
cap s1 "pair_kernel|wgrad_small|conv_taps_kernel|wgrad_taps" 0 30 ""
rm -f gpurun_out/s1.ncu-rep gpurun_out/tc3_c128b.ncu-rep gpurun_out/tc3_c64b.ncu-rep
du -sh gpurun_out
