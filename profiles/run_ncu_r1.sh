#!/bin/bash
# Round-1 ncu evidence (run under gpurun from the repo root).  Every ncu command is preceded by the same command
# without ncu (B200_PROFILING.md).  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --packs 256 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# name, only, packs, case substring (profiles/run_ncu_case.sh: plain run, then ncu --set full on the scan kernel)
profiles/run_ncu_case.sh r1_w20    c2  256 "w=20 lt(median) count"
profiles/run_ncu_case.sh r1_w8     c2  256 "w=8 lt(median) count"
profiles/run_ncu_case.sh r1_raw64  c1  128 "raw u64 1Mi bw(2^58,3*2^58) count"
profiles/run_ncu_case.sh r1_dictin c2s 64  "dict in{64} count"
profiles/run_ncu_case.sh r1_c3dict c3  128 "acct(dict15) in{64} sum/min/max i64"
profiles/run_ncu_case.sh r1_hash64 c3  128 "acct(bitpack40) in{64} sum/min/max i64"
profiles/run_ncu_case.sh r1_agg90  c3  128 "ts range(90%) sum/min/max i64"
profiles/run_ncu_case.sh r1_and2   c3  128 "AND acct(bitpack40) lt(median) bitset+count"
ls -la gpurun_out/ | head -80; du -sh gpurun_out
