#!/bin/bash
# One ncu --set full capture of the scan kernel for a sweep case (run under gpurun from the repo root):
#   profiles/run_ncu_case.sh <name> <only> <packs> "<case substring>"
# The same command runs first WITHOUT ncu (B200_PROFILING.md); text exports are made on the box.
set -u
mkdir -p gpurun_out
name=$1; only=$2; packs=$3; match=$4
C="python profiles/ncu_cases.py --only $only --packs $packs --match"
$C "$match" > gpurun_out/plain_$name.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_ -s 2 -c 1 -f -o gpurun_out/prof_$name $C "$match" > gpurun_out/ncu_$name.log 2>&1
echo "$name rc=$? $(tail -1 gpurun_out/plain_$name.log)"
ncu -i gpurun_out/prof_$name.ncu-rep --page details > gpurun_out/prof_${name}_details.txt 2>/dev/null
ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/prof_${name}_source.csv.gz
case " ${KEEP_REP:-} " in *" $name "*) ;; *) rm -f gpurun_out/prof_$name.ncu-rep ;; esac
