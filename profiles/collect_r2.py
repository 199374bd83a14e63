#!/usr/bin/env python3
"""collect_r2.py — copies the round-2 ncu summaries from gpurun_out/ (scratch) into profiles/ (tracked) and writes the two
traffic files bench.py reads (`roofline.traffic`, `roofline_c3.traffic`).

usage: python profiles/collect_r2.py            (after profiles/run_ncu_r2.sh ran under gpurun)"""
import csv
import gzip
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
PACK_ROWS = 1 << 20   # rows per pack of the config-3 cases (profiles/sweep_configs.py)

CASES = {
    # name: (rows per launch, algorithmic note)
    "w20": ("256 packs x 4 Mi rows, w = 20, Less(median) -> count", 256 * (4 << 20)),
    "w8": ("256 packs x 4 Mi rows, w = 8, Less(median) -> count (issue-bound narrow width)", 256 * (4 << 20)),
    "c3dict": ("128 packs x 1 Mi rows: ts(w=40) range 0.1 % AND acct(dict, 15-bit codes) IN{64} -> count + sum/min/max(int64 raw)", 128 * PACK_ROWS),
    "hash64": ("128 packs x 1 Mi rows: ts(w=40) range AND acct(bitpack40) IN{64} -> sum/min/max(int64 raw)", 128 * PACK_ROWS),
    "ts01": ("128 packs x 1 Mi rows: ts(w=40) range 0.1 % -> sum/min/max(int64 raw)", 128 * PACK_ROWS),
    "agg90": ("128 packs x 1 Mi rows: ts(w=40) range 90 % -> sum/min/max(int64 raw)", 128 * PACK_ROWS),
}


def to_bytes(txt):
    v, unit = txt.split()
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    head, units, last = rows[0], rows[1], rows[-1]
    return {k: (v + " " + u).strip() for k, u, v in zip(head, units, last)}


def main():
    for name, (launch, rows) in CASES.items():
        tag = "r2" if name in ("w20", "w8") else "r2b"   # r2b: captures of the warp-autonomous kernel (profiles/run_ncu_r2b.sh)
        det = os.path.join(SRC, f"prof_{tag}_{name}_details.txt")
        raw = os.path.join(SRC, f"prof_{tag}_{name}_raw.csv")
        if not os.path.exists(det):
            print("missing", det)
            continue
        shutil.copy(det, os.path.join(DST, f"r2_ncu_{name}_details.txt"))
        m = raw_metrics(raw)
        rec = {
            "kernel": "kx::" + m["Kernel Name"].replace("void ", ""),
            "launch": launch + f" (profiles/run_ncu_{tag}.sh, case {tag}_{name})",
            "rows": rows,
            "dram_bytes_read": to_bytes(m["dram__bytes_read.sum"]),
            "dram_bytes_write": to_bytes(m["dram__bytes_write.sum"]),
            "duration_us_under_ncu": float(m["gpu__time_duration.sum"].split()[0]),
            "warp_instructions": float(m["smsp__inst_executed.sum"].split()[0]),
            "source": f"profiles/r2_ncu_{name}_details.txt (ncu --set full --clock-control none)",
        }
        if name == "w20":
            rec["algorithmic_bytes"] = rows * 20 // 8
            shutil.copy(raw, os.path.join(DST, "r2_ncu_w20_raw.csv"))
        json.dump(rec, open(os.path.join(DST, f"r2_ncu_{name}_traffic.json"), "w"), indent=1)
        print(name, rec["duration_us_under_ncu"], "us", rec["dram_bytes_read"] / 1e6, "MB read")
    for name in ("c3dict", "w20"):
        s = os.path.join(SRC, f"prof_{'r2' if name == 'w20' else 'r2b'}_{name}_source.csv.gz")
        if os.path.exists(s):
            shutil.copy(s, os.path.join(DST, f"r2_ncu_{name}_source.csv.gz"))
    for a, b in (("launches_r2.csv", "r2_launches_bench.csv"), ("r2_sweep_configs.json", "r2_sweep_configs.json"), ("sweep_final.json", "r2_sweep_configs.json"),
                 ("bench_final.json", "r2_bench_1gpu.json"), ("bench_ref_final.json", "r2_bench_ref.json")):
        if os.path.exists(os.path.join(SRC, a)):
            shutil.copy(os.path.join(SRC, a), os.path.join(DST, b))


if __name__ == "__main__":
    main()
