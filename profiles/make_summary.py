#!/usr/bin/env python3
"""make_summary.py — renders a sweep JSON (profiles/sweep_configs.py) as the markdown table kept under profiles/.

usage: python profiles/make_summary.py gpurun_out/sweep.json [more.json ...] > profiles/r1_sweep_summary.md
Later files override earlier ones case by case (re-measured cases)."""
import json
import sys


def main():
    peak, rows, order, health = 6450.0, {}, [], []
    for path in sys.argv[1:]:
        d = json.load(open(path))
        peak = d.get("peak_GBps", peak)
        health += d.get("box_health", {}).get("canary_frac_of_peak", [])
        for r in d["results"]:
            if r["case"] not in rows:
                order.append(r["case"])
            rows[r["case"]] = r
    print("# Round-1 device-time sweep over BASELINE.json configs 1-4 (B200, `profiles/sweep_configs.py`)\n")
    print(f"Peak = measured copy bandwidth {peak:.0f} GB/s (`MEASURED_PEAKS.json`); times are CUDA-event durations of the scan kernels "
          "(pre-pass kernels included; median of 5 launches, inputs >> L2).  Every case first checks its result against the oracle: bitsets "
          "and counts bit for bit, integer aggregates bit for bit, float64 sums against the exact sum.  \"GB/s\" = algorithmic bytes "
          "(SURVEY §8d) per second; values above 100 % mean the launch touched fewer bytes than the algorithmic count (read-only "
          "streams beat the copy peak; value columns are only read where rows match).")
    if health:
        print(f"\nBox health: the canary scan (raw uint64 range → count, 8 B/row) ran at {', '.join(f'{100 * h:.0f} %' for h in health)} of the peak "
              "before / after the sweep.")
    print("\n| case | kernel ms | G rows/s | alg. GB/s | % of peak | selectivity |\n|---|---:|---:|---:|---:|---:|")
    c4 = []
    for name in order:
        r = rows[name]
        if "error" in r:
            print(f"| {name} | ERROR {r['error']} | | | | |")
        elif "rows_per_s" in r:
            print(f"| {name} | {r['kernel_ms']:.3f} | {r['rows_per_s'] / 1e9:.0f} | {r['algorithmic_GBps']:.0f} | {100 * r['frac_of_measured_peak']:.1f} | {r.get('selectivity', 0):.4f} |")
        else:
            c4.append(r)
    if c4:
        print("\n## Config 4 — pruning over the resident statistics index (1 B rows = 15 259 packs x 65 536)\n")
        for r in c4:
            if "values_per_s" in r:
                print(f"* {r['case']}: {r['values']:,} values in {r['seconds_incl_h2d_of_values']:.2f} s incl. H2D = {r['values_per_s'] / 1e6:.0f} M values/s; {r['parity']}")
            elif "packs_per_s" in r:
                print(f"* {r['case']}: kernel {r['kernel_ms'] * 1e3:.1f} us, call {r['total_ms'] * 1e3:.1f} us, {r['packs_per_s'] / 1e9:.2f} G packs/s, "
                      f"{r['bit_probes_per_s'] / 1e9:.1f} G bit probes/s, {r['survivors']} survivors; {r['parity']}")
            else:
                print(f"* {json.dumps(r)}")


if __name__ == "__main__":
    main()
