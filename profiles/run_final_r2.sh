#!/bin/bash
# Final round-2 evidence in one gpurun call (run from the repo root): the full -m gpu suite, both bench arms, the sweep over
# the BASELINE configs, the ncu launch list of the bench and the ncu --set full captures (each after the same command ran
# without ncu, B200_PROFILING.md).  Outputs land in gpurun_out/; profiles/collect_r2.py copies the summaries to profiles/.
set -u
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/tests_final.log 2>&1; tail -4 gpurun_out/tests_final.log
python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python profiles/sweep_configs.py --out gpurun_out/sweep_final.json > gpurun_out/sweep_final.log 2>&1; echo "sweep rc=$?"
B="python bench.py --steps 5 --warmup 3 --packs 256 --no-cpu-baseline --no-configs --no-strong"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
profiles/run_ncu_case.sh r2_w20     c2  256 "w=20 lt(median) count"
profiles/run_ncu_case.sh r2_w8      c2  256 "w=8 lt(median) count"
profiles/run_ncu_case.sh r2b_c3dict c3  128 "acct(dict15) in{64} sum/min/max i64"
profiles/run_ncu_case.sh r2b_hash64 c3  128 "acct(bitpack40) in{64} sum/min/max i64"
profiles/run_ncu_case.sh r2b_ts01   c3  128 "ts range(0.1%) sum/min/max i64"
profiles/run_ncu_case.sh r2b_agg90  c3  128 "ts range(90%) sum/min/max i64"
du -sh gpurun_out
