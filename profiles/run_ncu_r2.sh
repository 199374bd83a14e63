#!/bin/bash
# Round-2 ncu evidence (run under gpurun from the repo root).  Every ncu command is preceded by the same command
# without ncu (B200_PROFILING.md).  Outputs land in gpurun_out/; profiles/collect_r2.py copies the summaries to profiles/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --packs 256 --no-cpu-baseline --no-configs --no-strong"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# name, only, packs, case substring (profiles/run_ncu_case.sh: plain run, then ncu --set full on the scan kernel)
profiles/run_ncu_case.sh r2_w20    c2  256 "w=20 lt(median) count"
profiles/run_ncu_case.sh r2_c3dict c3  128 "acct(dict15) in{64} sum/min/max i64"
profiles/run_ncu_case.sh r2_hash64 c3  128 "acct(bitpack40) in{64} sum/min/max i64"
profiles/run_ncu_case.sh r2_ts01   c3  128 "ts range(0.1%) sum/min/max i64"
profiles/run_ncu_case.sh r2_agg90  c3  128 "ts range(90%) sum/min/max i64"
ls -la gpurun_out/ | tail -40; du -sh gpurun_out
