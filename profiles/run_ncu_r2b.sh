#!/bin/bash
# Round-2 ncu captures of the warp-autonomous kernel (kx_warp.cu); run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
for c in "$@"; do
case $c in
c3dict) profiles/run_ncu_case.sh r2b_c3dict c3  128 "acct(dict15) in{64} sum/min/max i64" ;;
hash64) profiles/run_ncu_case.sh r2b_hash64 c3  128 "acct(bitpack40) in{64} sum/min/max i64" ;;
ts01)   profiles/run_ncu_case.sh r2b_ts01   c3  128 "ts range(0.1%) sum/min/max i64" ;;
agg90)  profiles/run_ncu_case.sh r2b_agg90  c3  128 "ts range(90%) sum/min/max i64" ;;
esac
done
