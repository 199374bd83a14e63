#!/bin/bash
# ncu --set full captures of the two kernels outside scan_kernel (run under gpurun from the repo root): the time-bucketed
# reduce (hourly windows case of the sweep) and the string match pre-pass (fixed 20-byte addresses).  The plain command
# runs first (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
cap() {   # name, sweep tag, kernel regex, launches to skip
  python profiles/sweep_configs.py --only $2 --out gpurun_out/plain_$1.json > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/prof_$1 python profiles/sweep_configs.py --only $2 --out gpurun_out/ncu_$1.json > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/prof_$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$1.ncu-rep
}
cap r1_bucket c3b bucket_kernel 2
cap r1_strmatch c4s strmatch_kernel 2
