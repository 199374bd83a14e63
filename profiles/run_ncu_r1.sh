#!/bin/bash
# Round-1 ncu evidence (run under gpurun from the repo root).  Every ncu command is preceded by the same command
# without ncu (B200_PROFILING.md).  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --packs 256 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"

cap() {   # name, only, packs, match
  local C="python profiles/ncu_cases.py --only $2 --packs $3 --match"
  $C "$4" > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1_$1 $C "$4" > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$? $(tail -1 gpurun_out/plain_$1.log)"
  # text exports are made on the box (gpurun_out/ may carry at most 64 MiB back); only KEEP_REP reports travel
  ncu -i gpurun_out/prof_r1_$1.ncu-rep --page details > gpurun_out/prof_r1_$1_details.txt 2>/dev/null
  ncu -i gpurun_out/prof_r1_$1.ncu-rep --page raw --csv > gpurun_out/prof_r1_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_r1_$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/prof_r1_$1_source.csv.gz
  case " $KEEP_REP " in *" $1 "*) ;; *) rm -f gpurun_out/prof_r1_$1.ncu-rep ;; esac
}
KEEP_REP="${KEEP_REP:-w20}"
cap w20   c2  256 "w=20 lt(median) count"
cap w8    c2  256 "w=8 lt(median) count"
cap raw64 c1  128 "raw u64 1Mi bw(2^58,3*2^58) count"
cap dictin c2s 64 "dict in{64} count"
cap agg90 c3  128 "ts range(90%) sum/min/max i64"
cap hash4096 c3 128 "acct(bitpack40) in{4096} sum/min/max i64"
cap and2  c3  128 "AND acct(bitpack40) lt(median) bitset+count"
# stay below the 64 MiB pull limit: drop the binary reports first if the directory grew too large
if [ "$(du -sm gpurun_out | cut -f1)" -gt 50 ]; then rm -f gpurun_out/*.ncu-rep; fi
ls -la gpurun_out/ | head -60; du -sh gpurun_out
