#!/usr/bin/env python3
"""bench.py — scan-filter-aggregate throughput of the KnoxDB pack scan on B200.

Headline workload (BASELINE.json configs[1]): bit-packed uint64 column (BitpackContainer, min-FOR, width w = 20 bits),
packs of 4 Mi rows, predicate `Less(median)` → match count per pack, fused decode + compare + popcount in one kernel
launch over all resident packs of the rank.  One "step" = one pass over all resident packs.

One JSON line (rank 0):
  value         rows/s with the packs resident in HBM (wall clock around K steps: launch, descriptor upload, counts back;
                N > 1: plus the library's own NCCL all-gather + device combine of the per-rank totals, kx_scan_sharded)
  e2e           rows/s through kx_scan_host: encoded blocks in pinned HOST memory, H2D of the blocks and D2H of
                bitsets + counts inside the timed region (PCIe-bound)
  roofline      algorithmic bytes of the scan kernel ÷ its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the CPU port of the reference's fused bitpack compare on the host cores (N = 1)
  configs       (N = 1) sub-records for the other BASELINE configs: c1 raw uint64 Between → bitset + popcount,
                c3 the NORTH-STAR path — ts RANGE AND acct IN{64} + sum/min/max over int64 / float64 at three
                selectivities, with a roofline on the bytes the launch must touch — and c4 zone-map + bloom pruning
  strong        config 5: a FIXED table of 8 B rows (1907 packs x 4 Mi) sharded by pack over the N ranks

`--impl reference` times the reference's CPU algorithm for the headline workload on the host cores (the Go reference
cannot be built here: no Go toolchain; oracle/ holds its C restatement, plus an AVX-512 kernel that computes the same
words) on a bounded sample of >= 2 GiB of distinct packs per step.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W_BITS = 20
PACK_ROWS = 4 * 1024 * 1024
FOR_BASE = 1_000_000
FIELD = 1
METRIC = "scan_filter_count_rows_per_s"
UNIT = "rows/s"
DISTINCT_PACKS = 208            # 208 x 10.5 MB = 2.18 GB of distinct packed payload (> 2 GiB, >> any last-level cache)
STRONG_PACKS = 1907             # config 5: 8 000 000 000 rows / 4 194 304


def workload_config():
    """identical in both arms (the driver compares it)"""
    return {"workload": f"bitpacked uint64 w={W_BITS} min-FOR, {PACK_ROWS}-row packs, Less(median) -> match count per pack",
            "pack_rows": PACK_ROWS, "width_bits": W_BITS, "predicate": "Less(median)", "output": "match count per pack",
            "l2_policy": "inputs larger than the last-level cache: every step streams its packs from memory (GPU arm: >= 10.7 GB of "
                         "resident packs per GPU vs 126 MB L2; CPU arm: 2.18 GB of distinct packs vs the host LLC); no flush needed",
            "sharding": "packs; no data-path collective, one NCCL all-gather of a 208 B record per rank and query inside the library"}


# ------------------------------------------------------------------------------------------------ block encoders
def uv(x):   # pkg/num/varint.go PutUvarint
    if x <= 240:
        return bytes([x])
    if x <= 2287:
        y = x - 240
        return bytes([241 + (y >> 8), y & 0xFF])
    if x <= 67823:
        y = x - 2288
        return bytes([249, y >> 8, y & 0xFF])
    nb = max(3, (x.bit_length() + 7) // 8)
    return bytes([247 + nb]) + x.to_bytes(nb, "big")


def random_payload(rng, nbytes):
    """any bit string is a valid stream of uniform w-bit fields; + 64 readable bytes for the CPU arm's vector loads"""
    words = rng.integers(0, 2**64, nbytes // 8 + 8, dtype=np.uint64)
    words[nbytes // 8:] = 0
    return words.view(np.uint8)


def bitpack_header(base, w, n):
    """[IntBitpacked=4][uvarint For][uvarint Log2][uvarint N] — int_bitpack.go:92-98"""
    return bytes([4]) + uv(base) + uv(w) + uv(n)


def pack_bits(vals, w):
    """LSB-first bit stream of w-bit fields in 64-bit LE words (bitpack/encode.go:216-246)"""
    n = vals.size
    bits = ((vals.astype(np.uint64)[:, None] >> np.arange(w, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)
    pad = (-bits.size) % 64
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, dtype=np.uint8)])
    return np.packbits(bits, bitorder="little")


def enc_bitpack(vals_u64_pattern, signed=False):
    v = vals_u64_pattern.astype(np.int64) if signed else vals_u64_pattern.astype(np.uint64)
    mn, mx = int(v.min()), int(v.max())
    w = (mx - mn).bit_length()
    fields = (v - v.dtype.type(mn)).astype(np.uint64)
    hdr = bitpack_header(mn & (2**64 - 1), w, v.size)
    return np.concatenate([np.frombuffer(hdr, dtype=np.uint8), pack_bits(fields, w) if w else np.zeros(0, np.uint8)]), w


def enc_raw(vals, is_float=False):
    return np.concatenate([np.frombuffer(bytes([15 if is_float else 7]) + uv(vals.size), dtype=np.uint8), np.ascontiguousarray(vals).view(np.uint8)])


def enc_dict(vals_u64):
    """[IntDictionary=5][Dict container: sorted unique values][Codes container (uint16)] — int_dict.go:76-99"""
    uniq, codes = np.unique(vals_u64, return_inverse=True)
    d, _ = enc_bitpack(uniq)
    c, cw = enc_bitpack(codes.astype(np.uint64))
    return np.concatenate([np.frombuffer(bytes([5]), dtype=np.uint8), d, c]), cw, uniq


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        inwin = [s for (ts, s) in self.samples if self.t0 is None or (self.t0 <= ts <= (self.t1 or ts) + 0.15)]
        if not inwin:   # timed region shorter than one sampling period: take the nearest samples
            inwin = [s for (_, s) in self.samples[-3:]]
        for s in inwin:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_info():
    model, mhz = "unknown", []
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and model == "unknown":
                model = line.split(":", 1)[1].strip()
            if line.startswith("cpu MHz"):
                mhz.append(float(line.split(":", 1)[1]))
    except Exception:
        pass
    return {"model": model, "mhz_median": float(np.median(mhz)) if mhz else None, "logical_cpus": os.cpu_count()}


def ncu_traffic(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_scan(payloads, threshold, nthreads, min_seconds=0.0, scalar=False):
    """the CPU port of bitpack.Less over packed words + popcount (oracle/), packs statically partitioned over threads;
    nthreads < 0 inside the library call = scalar port only"""
    import ctypes as C
    import oracle as ko
    L = ko.lib()
    n = len(payloads)
    bits = [np.zeros(PACK_ROWS // 8 + 8, dtype=np.uint8) for _ in range(n)]
    pp = (C.c_void_p * n)(*[a.ctypes.data for a in payloads])
    bp = (C.c_void_p * n)(*[b.ctypes.data for b in bits])
    nr = (C.c_size_t * n)(*([PACK_ROWS] * n))
    nt = -nthreads if scalar else nthreads
    t0 = time.perf_counter(); reps = 0; total = 0
    while True:
        total = L.ko_baseline_bitpack_scan(pp, nr, n, W_BITS, ko.LT, threshold, 0, bp, nt)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= 4096:
            break
    return reps * n * PACK_ROWS / dt, dt, reps, int(total), bits


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the headline workload, all host threads"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as ko
    ncores = os.cpu_count() or 1
    rng = np.random.default_rng(1)
    nbytes = PACK_ROWS * W_BITS // 8
    payloads = [random_payload(rng, nbytes) for _ in range(DISTINCT_PACKS)]
    thr = 1 << (W_BITS - 1)
    simd = bool(ko.lib().ko_simd_available())
    for _ in range(max(args.warmup, 1)):
        cpu_scan(payloads, thr, ncores)
    t_all = time.perf_counter()
    rows = 0; secs = 0.0
    for _ in range(args.steps):
        r, dt, reps, _, _ = cpu_scan(payloads, thr, ncores)
        rows += DISTINCT_PACKS * PACK_ROWS * reps; secs += dt
    v = rows / secs
    sample = (f"{DISTINCT_PACKS} distinct packs x {PACK_ROWS} rows ({DISTINCT_PACKS * nbytes / 2**30:.2f} GiB packed) per step, "
              f"{'AVX-512 VBMI kernel (oracle/ko_simd.c) computing the words of' if simd else 'scalar C port of'} bitpack.Less + popcount, "
              f"{ncores} pthreads; the Go reference itself cannot run here (no Go toolchain)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(),
            "step": {"packs_per_step": DISTINCT_PACKS, "rows_per_step": DISTINCT_PACKS * PACK_ROWS},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample, "simd": simd, "cpu": cpu_info()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ sub-records (N = 1)
def median_kernel_ms(ctx, fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ks = []
    for _ in range(reps):
        fn()
        ks.append(ctx.last_scan_stats()["kernel_ms"])
    return float(np.median(ks))


def config1(ctx, kb, peak):
    """BASELINE config 1: raw uint64, 1 Mi-row packs, Between(5, 127) (internal/cmp/tests/bench.go:48-60) and a ~50 %
    range → bitset + popcount; 256 packs per launch (2 GiB >> L2)"""
    M1, npacks, nd = 1 << 20, 256, 4
    rng = np.random.default_rng(11)
    vals = [rng.integers(0, 2**60 - 1, M1, dtype=np.uint64) for _ in range(nd)]
    for p in range(npacks):
        ctx.block_put(100000 + p, 1, 1, kb.UINT64, enc_raw(vals[p % nd]))
    refs = ctx.pack_refs([(100000 + p, 1) for p in range(npacks)])
    nrows = [M1] * npacks
    buf = ctx.host_array(ctx.bitset_layout(nrows)[1])
    out = []
    for name, a, b in (("Between(5,127)", 5, 127), ("Between(2^58,3*2^58)", 1 << 58, 3 << 58)):
        prog = kb.Program(ctx, [kb.Leaf(1, kb.UINT64, kb.RANGE, a, b)])
        r = ctx.scan(prog, refs, nrows=nrows, want_bitsets=True, bitset_buf=buf)
        for d in range(nd):   # parity: numpy truth of U(v - a) <= U(b - a)
            want = np.packbits(((vals[d] - np.uint64(a)) <= np.uint64(b - a)).astype(np.uint8), bitorder="little")
            assert (r["bitsets"][d] == want).all() and int(r["counts"][d]) == int(np.unpackbits(want).sum()), "config 1 parity"
        ms = median_kernel_ms(ctx, lambda: ctx.scan(prog, refs, nrows=nrows, want_bitsets=True, bitset_buf=buf))
        gbs = npacks * M1 * 8.125 / (ms * 1e-3) / 1e9
        out.append({"case": f"raw uint64 1Mi-row packs x {npacks}, {name} -> bitset + popcount", "kernel_ms": ms, "rows_per_s": npacks * M1 / (ms * 1e-3),
                    "bytes_per_row": 8.125, "achieved_GBps": gbs, "frac_of_peak": gbs / peak, "parity": "bit-exact vs numpy on every distinct pack"})
        prog.close()
    for p in range(npacks):
        ctx.block_drop(100000 + p, 1, 1)
    return out


def config2_widths(ctx, kb, peak):
    """BASELINE config 2 at other widths than the headline's: bit-packed uint64, 1 Mi-row packs x 512 (>= 0.5 GiB at w = 8,
    inputs >> L2), Less(median) -> match count per pack; parity: numpy truth on every distinct pack."""
    M1, npacks, nd = 1 << 20, 512, 4
    out = []
    for w in (8, 12, 32, 48):
        rng = np.random.default_rng(100 + w)
        vals = [rng.integers(0, 2**w, M1, dtype=np.uint64) for _ in range(nd)]
        for v in vals:
            v[0], v[1] = 0, 2**w - 1          # every pack spans the full width
        blocks = [enc_bitpack(v)[0] for v in vals]
        for p in range(npacks):
            ctx.block_put(200000 + p, 1, 1, kb.UINT64, blocks[p % nd])
        refs = ctx.pack_refs([(200000 + p, 1) for p in range(npacks)])
        nrows = [M1] * npacks
        thr = 2**(w - 1)
        prog = kb.Program(ctx, [kb.Leaf(1, kb.UINT64, kb.LT, thr)])
        r = ctx.scan(prog, refs, nrows=nrows)
        for d in range(nd):
            assert int(r["counts"][d]) == int((vals[d] < np.uint64(thr)).sum()), "config 2 parity"
        ms = median_kernel_ms(ctx, lambda: ctx.scan(prog, refs, nrows=nrows))
        gbs = npacks * M1 * (w / 8.0) / (ms * 1e-3) / 1e9
        out.append({"case": f"bit-packed uint64 w={w}, 1Mi-row packs x {npacks}, Less(median) -> count", "kernel_ms": ms, "rows_per_s": npacks * M1 / (ms * 1e-3),
                    "bytes_per_row": w / 8.0, "achieved_GBps": gbs, "frac_of_peak": gbs / peak, "parity": "counts equal numpy on every distinct pack"})
        prog.close()
        for p in range(npacks):
            ctx.block_drop(200000 + p, 1, 1)
    return out


def config3(ctx, kb, peak):
    """BASELINE config 3 — the north-star path: ts BETWEEN [t0, t1] AND acct IN {64 values} → count / sum / min / max over an
    int64 and a float64 amount column, 256 packs x 1 Mi rows (ts: sorted, bit-packed; acct: dictionary with 15-bit
    codes; amounts: raw 64-bit), at ts selectivities 0.1 % / 10 % / 90 %, and the same without the acct leaf."""
    M1, npacks, nd = 1 << 20, 256, 2
    rng = np.random.default_rng(3)
    ts = [(1_700_000_000 + np.cumsum(rng.integers(0, 3, M1))).astype(np.int64) for _ in range(nd)]
    uniq = np.unique(rng.integers(0, 2**40, 40000, dtype=np.uint64))[:32768]
    acct = [uniq[rng.integers(0, uniq.size, M1)] for _ in range(nd)]
    amt_i = [rng.integers(-10**9, 10**9, M1).astype(np.int64) for _ in range(nd)]
    amt_k = [rng.integers(0, 2**40, M1).astype(np.int64) for _ in range(nd)]
    amt_f = [k.astype(np.float64) / 64.0 for k in amt_k]           # multiples of 1/64: the exact sum is an integer sum
    b_ts = [enc_bitpack(t.view(np.uint64), signed=True) for t in ts]
    b_ac = [enc_dict(a) for a in acct]
    base = 200000
    for p in range(npacks):
        d = p % nd
        ctx.block_put(base + p, 1, 1, kb.INT64, b_ts[d][0]); ctx.block_put(base + p, 1, 2, kb.UINT64, b_ac[d][0])
        ctx.block_put(base + p, 1, 3, kb.INT64, enc_raw(amt_i[d])); ctx.block_put(base + p, 1, 4, kb.FLOAT64, enc_raw(amt_f[d], True))
    refs = ctx.pack_refs([(base + p, 1) for p in range(npacks)])
    nrows = [M1] * npacks
    e_ts, e_ac = b_ts[0][0].size / M1, b_ac[0][0].size / M1
    tmin, tmax = int(min(t[0] for t in ts)), int(max(t[-1] for t in ts))
    in64 = uniq[:: uniq.size // 64][:64]
    mult = [len(range(d, npacks, nd)) for d in range(nd)]
    out = []
    for with_acct in (True, False):
        for frac in (0.001, 0.1, 0.9):
            t1 = tmin + int((tmax - tmin) * frac)
            leaves = [kb.Leaf(1, kb.INT64, kb.RANGE, tmin, t1)] + ([kb.Leaf(2, kb.UINT64, kb.IN, values=in64)] if with_acct else [])
            prog = kb.Program(ctx, leaves)
            masks = [((ts[d] >= tmin) & (ts[d] <= t1)) & (np.isin(acct[d], in64) if with_acct else True) for d in range(nd)]
            rec = {"case": f"ts RANGE({frac * 100:g}%)" + (" AND acct(dict, 15-bit codes) IN{64}" if with_acct else "") + " -> count/sum/min/max",
                   "rows_per_launch": npacks * M1, "selectivity": float(sum(int(m.sum()) * k for m, k in zip(masks, mult))) / (npacks * M1)}
            # bytes the launch must touch: both filter columns for every row + the 32 B sectors of the value column that
            # hold at least one matching row (dense tiles are streamed whole: the model is a lower bound there)
            sect = sum(int(np.count_nonzero(m.reshape(-1, 4).any(axis=1))) * k for m, k in zip(masks, mult))
            filt_bytes = npacks * M1 * (e_ts + (e_ac if with_acct else 0.0))
            for kind, field, kbt, vals in (("int64", 3, kb.INT64, amt_i), ("float64", 4, kb.FLOAT64, amt_f)):
                r = ctx.scan(prog, refs, nrows=nrows, aggs=[(field, kbt)])
                g = r["aggs"][0]
                cnt = sum(int(m.sum()) * k for m, k in zip(masks, mult))
                assert g.count == cnt == int(r["counts"].sum()), "config 3 count parity"
                if kind == "int64":
                    s = sum(int(v[m].sum()) * k for v, m, k in zip(vals, masks, mult)) & (2**64 - 1)
                    assert g.sum_bits == s, "config 3 int64 sum parity"
                    if cnt:
                        assert g.value("min", kbt) == min(int(v[m].min()) for v, m in zip(vals, masks) if m.any())
                        assert g.value("max", kbt) == max(int(v[m].max()) for v, m in zip(vals, masks) if m.any())
                elif cnt:
                    got = g.value("sum", kbt)
                    exact_k = sum(int(amt_k[d][masks[d]].sum()) * mult[d] for d in range(nd))       # exact: integers / 64
                    exact = float(exact_k) / 64.0 if exact_k < 2**1000 else float("inf")
                    seq = 0.0                                                                   # the reference's order: one running float64 sum, pack by pack
                    for p in range(npacks):
                        seq = float(np.cumsum(np.concatenate([[seq], vals[p % nd][masks[p % nd]]]))[-1])
                    rec["f64_sum_rel_dist_to_sequential_reference_order"] = abs(got - seq) / abs(seq)
                    rec["f64_sum_rel_dist_to_exact"] = abs(got - exact) / abs(exact)
                    rec["sequential_reference_order_rel_dist_to_exact"] = abs(seq - exact) / abs(exact)
                    assert rec["f64_sum_rel_dist_to_exact"] <= 1e-14, "config 3 float64 sum vs the exact sum"
                    assert rec["f64_sum_rel_dist_to_sequential_reference_order"] <= 1e-10
                ms = median_kernel_ms(ctx, lambda: ctx.scan(prog, refs, nrows=nrows, aggs=[(field, kbt)]))
                touched = filt_bytes + 32.0 * sect
                algo = npacks * M1 * (e_ts + (e_ac if with_acct else 0.0) + 8.0)
                rec[kind] = {"kernel_ms": ms, "rows_per_s": npacks * M1 / (ms * 1e-3),
                             "roofline_touched": {"bytes": touched, "achieved_GBps": touched / (ms * 1e-3) / 1e9, "frac_of_peak": touched / (ms * 1e-3) / 1e9 / peak},
                             "roofline_survey_8d": {"bytes": algo, "achieved_GBps": algo / (ms * 1e-3) / 1e9, "frac_of_peak": algo / (ms * 1e-3) / 1e9 / peak}}
            rec["parity"] = "count, int64 sum/min/max bit-exact vs numpy; float64 sum <= 1e-14 of the exact sum"
            out.append(rec)
            prog.close()
    for p in range(npacks):
        for f in (1, 2, 3, 4):
            ctx.block_drop(base + p, 1, f)
    return out


def config4(ctx, kb):
    """BASELINE config 4: zone-map + bloom pruning over a 1 B-row block table (15 259 packs x 65 536 rows) on the resident
    statistics index; the filters are built on the device from 20-byte address strings."""
    npacks, per_pack, nd = 15259, 65536, 8
    rng = np.random.default_rng(4)
    sets = [rng.integers(0, 256, (per_pack, 20), dtype=np.uint8) for _ in range(nd)]
    offs = np.arange(per_pack + 1, dtype=np.uint32) * 20
    heights = np.arange(npacks, dtype=np.int64) * per_pack
    mins = np.stack([heights.view(np.uint64), np.zeros(npacks, dtype=np.uint64)])
    maxs = np.stack([(heights + per_pack - 1).view(np.uint64), np.zeros(npacks, dtype=np.uint64)])
    st = kb.Stats(ctx, [(1, kb.INT64), (2, kb.BYTES)], mins, maxs)
    flat = [x.reshape(-1) for x in sets]
    t0 = time.time()
    for p in range(npacks):
        st.build_bloom(1, p, kb.BYTES, flat[p % nd], per_pack, 2, offsets=offs)
    t_build = time.time() - t0
    x = sets[5][1234]
    hx = kb.lib().kx_hash_bytes(x.ctypes.data, 20)
    lo, hi = 1000 * per_pack + 17, 2500 * per_pack
    out = [{"case": "bloom build on device: XXH3-64 of 20-byte strings + 4 bit sets per value", "values": npacks * per_pack, "seconds_incl_h2d": t_build,
            "values_per_s": npacks * per_pack / t_build}]
    for name, leaves, hashes, zone in (("height BETWEEN (10% of packs) AND address = X", [kb.Leaf(1, kb.INT64, kb.RANGE, lo, hi), kb.Leaf(2, kb.BYTES, kb.EQ)], [[], [hx]], True),
                                       ("address = X (all packs probed)", [kb.Leaf(2, kb.BYTES, kb.EQ)], [[hx]], False)):
        prog = kb.Program(ctx, leaves)
        bits, n = st.prune(prog, hashes)
        alive = np.unpackbits(bits, bitorder="little")[:npacks].astype(bool)
        must = (np.arange(npacks) % nd) == 5
        if zone:
            must &= (heights <= hi) & (heights + per_pack - 1 >= lo)
        assert (alive | ~must).all(), "config 4: a pack that holds X was pruned (bloom filters have no false negatives)"
        ks, tt = [], []
        for _ in range(20):
            st.prune(prog, hashes)
            s_ = ctx.last_scan_stats()
            ks.append(s_["kernel_ms"]); tt.append(s_["total_ms"])
        km = float(np.median(ks))
        out.append({"case": name, "packs": npacks, "rows_represented": npacks * per_pack, "kernel_us": 1e3 * km, "call_us": 1e3 * float(np.median(tt)),
                    "packs_per_s": npacks / (km * 1e-3), "survivors": int(n), "packs_holding_X": int(must.sum()),
                    "bound": "launch latency / random 32 B sectors (not HBM bandwidth)"})
        prog.close()
    st.close()
    return out


# ------------------------------------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--packs", type=int, default=1024, help="resident packs per GPU (4 Mi rows each; 1024 packs = 10.7 GB packed)")
    ap.add_argument("--e2e-packs", type=int, default=64, help="packs per e2e step (host-resident blocks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 1 / 3 / 4 sub-records")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-size (config 5) strong-scaling sub-record")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import knoxdb_b200 as kb
    from knoxdb_b200 import shard
    from knoxdb_b200.lib import COMM_ID_BYTES

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = kb.Context(local)
    if world > 1:
        # the library owns the query's collective: its communicator id travels once over the launcher's process group
        cid = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            cid.copy_(torch.from_numpy(kb.Context.comm_unique_id().copy()))
        dist.broadcast(cid, 0)
        ctx.comm_init(world, rank, cid.cpu().numpy())
    peak, peak_kind = measured_peak()

    # ---- packs: DISTINCT_PACKS distinct payloads in pinned host memory, registered under npacks pack ids
    strong_lo, strong_hi = shard.shard_range(STRONG_PACKS, rank, world)
    npacks = max(args.packs, 0 if args.no_strong else strong_hi - strong_lo)
    rng = np.random.default_rng(1 + rank)
    nbytes = PACK_ROWS * W_BITS // 8
    hdr = np.frombuffer(bitpack_header(FOR_BASE, W_BITS, PACK_ROWS), dtype=np.uint8)
    pinned, payloads = [], []
    for _ in range(DISTINCT_PACKS):
        buf = ctx.host_array(hdr.size + nbytes + 64)
        buf[: hdr.size] = hdr
        buf[hdr.size:] = random_payload(rng, nbytes)
        pinned.append(buf[: hdr.size + nbytes])
        payloads.append(buf[hdr.size:])
    for p in range(npacks):
        assert ctx.block_put(p, 1, FIELD, kb.UINT64, pinned[p % DISTINCT_PACKS]) == PACK_ROWS
    thr_field = 1 << (W_BITS - 1)             # median of the uniform w-bit fields
    prog = kb.Program(ctx, [kb.Leaf(FIELD, kb.UINT64, kb.LT, FOR_BASE + thr_field)])
    step_packs = args.packs
    packs = ctx.pack_refs([(p, 1) for p in range(step_packs)])   # kx_packref[] built once, like a Go caller would
    nrows = [PACK_ROWS] * step_packs
    offs, total_bits = ctx.bitset_layout(nrows)
    bitbuf = ctx.host_array(total_bits)       # pinned result buffer

    # ---- parity spot check against numpy truth on one pack (the full checks live in tests/)
    r = ctx.scan(prog, [(0, 1), (1, 1)], nrows=nrows[:2], want_bitsets=True)
    words = payloads[0][: nbytes].view(np.uint64)
    sample_rows = 100_000
    bitoff = np.arange(sample_rows, dtype=np.uint64) * np.uint64(W_BITS)
    idx = (bitoff >> np.uint64(6)).astype(np.int64); sh = bitoff & np.uint64(63)
    lo = words[idx] >> sh
    hi = np.where(sh > 0, words[np.minimum(idx + 1, words.size - 1)] << ((np.uint64(64) - sh) & np.uint64(63)), np.uint64(0))
    fields = (lo | hi) & np.uint64((1 << W_BITS) - 1)
    truth = np.packbits((fields < np.uint64(thr_field)).astype(np.uint8), bitorder="little")
    assert (r["bitsets"][0][: sample_rows // 8] == truth[: sample_rows // 8]).all(), "GPU scan disagrees with numpy truth"

    def step_resident():
        if world > 1:
            # one call: scan kernel, ONE NCCL all-gather of the 208 B per-rank record and the rank-order combine, all
            # enqueued on the library's scan stream, one sync
            res = ctx.scan_sharded(prog, packs)
            return ctx.last_scan_stats(), res["total_count"]
        res = ctx.scan(prog, packs, nrows=nrows, want_bitsets=False)     # counts only: bitsets stay in HBM
        return ctx.last_scan_stats(), int(res["counts"].sum())

    def step_resident_bits():
        res = ctx.scan(prog, packs, nrows=nrows, want_bitsets=True, bitset_buf=bitbuf)
        return ctx.last_scan_stats(), int(res["counts"].sum())

    sampler = ClockSampler(local); sampler.start()
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.mark_start()
    t0 = time.perf_counter()
    kernel_ms = 0.0; total_ms = 0.0; launches = 0; matches = 0
    for _ in range(args.steps):
        st, matches = step_resident()
        kernel_ms += st["kernel_ms"]; total_ms += st["total_ms"]; launches += st["launches"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sampler.mark_stop()
    if world > 1:
        dist.barrier()
        tmax = torch.tensor([wall, total_ms, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        wall, total_ms, kernel_ms = [float(x) for x in tmax.tolist()]
    clocks = sampler.stop()
    rows_step = step_packs * PACK_ROWS * world
    value = rows_step * args.steps / wall

    # ---- roofline of the dominant kernel (scan_kernel): algorithmic bytes / CUDA-event time.
    # count-only launch: reads w/8 B per row, writes nothing per row
    alg_bytes = step_packs * PACK_ROWS * W_BITS / 8
    k_ms = kernel_ms / args.steps
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    # bitset-materialising variant (w/8 + 1/8 B per row) timed separately for the record
    for _ in range(2):
        step_resident_bits()
    kb_ms = 0.0
    nb = max(3, min(args.steps, 5))
    for _ in range(nb):
        st, _ = step_resident_bits()
        kb_ms += st["kernel_ms"]
    kb_ms /= nb
    achieved_bits = step_packs * PACK_ROWS * (W_BITS + 1) / 8 / (kb_ms * 1e-3) / 1e9

    # ---- e2e through the C ABI with HOST blocks (pinned): H2D of blocks + D2H of bitsets/counts per step
    e2e_packs = min(args.e2e_packs, step_packs)
    hb = [[pinned[p % DISTINCT_PACKS]] for p in range(e2e_packs)]
    e_nrows = [PACK_ROWS] * e2e_packs
    _, e_total_bits = ctx.bitset_layout(e_nrows)
    fields_spec = [(FIELD, kb.UINT64)]
    for _ in range(2):
        res_e = ctx.scan_host(prog, fields_spec, hb, nrows=e_nrows, want_bitsets=True, bitset_buf=bitbuf)
    # the host-block path (pipelined uploads, several batches per call) must give what the resident packs give
    res_r = ctx.scan(prog, ctx.pack_refs([(p, 1) for p in range(e2e_packs)]), nrows=e_nrows, want_bitsets=True)
    assert res_e["counts"].tolist() == res_r["counts"].tolist(), "kx_scan_host disagrees with the resident scan"
    assert all((x == y).all() for x, y in zip(res_e["bitsets"], res_r["bitsets"])), "kx_scan_host bitsets disagree with the resident scan"
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_steps = max(3, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        ctx.scan_host(prog, fields_spec, hb, nrows=e_nrows, want_bitsets=True, bitset_buf=bitbuf)
    torch.cuda.synchronize()
    e_wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
        tm = torch.tensor([e_wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e_wall = float(tm.item())
    e2e_value = e2e_packs * PACK_ROWS * world * e_steps / e_wall
    h2d = sum(int(b[0].size) for b in hb)
    d2h = int(e_total_bits) + 8 * e2e_packs

    # ---- config 5: a FIXED table of 8 B rows sharded by pack (strong scaling); every rank scans its contiguous range
    strong = None
    if not args.no_strong:
        s_refs = ctx.pack_refs([(p, 1) for p in range(strong_hi - strong_lo)])
        s_nrows = [PACK_ROWS] * (strong_hi - strong_lo)

        def strong_query():
            if world > 1:
                return ctx.scan_sharded(prog, s_refs, want_counts=False)["total_count"]
            return int(ctx.scan(prog, s_refs, nrows=s_nrows)["counts"].sum())
        for _ in range(3):
            strong_query()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s_steps = max(10, min(args.steps, 200))
        t0 = time.perf_counter()
        for _ in range(s_steps):
            s_total = strong_query()
        torch.cuda.synchronize()
        s_wall = time.perf_counter() - t0
        s_kernel = ctx.last_scan_stats()["kernel_ms"]
        if world > 1:
            dist.barrier()
            tm = torch.tensor([s_wall, s_kernel], device="cuda", dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            s_wall, s_kernel = [float(x) for x in tm.tolist()]
        strong = {"workload": "config 5: fixed table of 1907 packs x 4 Mi rows (8.0 B rows), sharded by contiguous pack range, count + combine over ranks",
                  "scaling": "strong", "rows_total": STRONG_PACKS * PACK_ROWS, "n_gpus": world, "queries": s_steps, "ms_per_query": 1e3 * s_wall / s_steps,
                  "rows_per_s": STRONG_PACKS * PACK_ROWS * s_steps / s_wall, "kernel_ms_last_query_max_over_ranks": s_kernel, "matches": int(s_total)}

    traffic, traffic_note = None, None
    tj = ncu_traffic("r2_ncu_w20_traffic.json") or ncu_traffic("r1_ncu_w20_traffic.json")
    if tj:
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * (alg_bytes / tj["algorithmic_bytes"])
        traffic_note = f"ncu dram__bytes_read+write of a {tj['launch']} scaled by packs ({tj['source']})"
    cfg = workload_config()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": cfg,
        "step": {"packs_per_gpu": step_packs, "rows_per_step": rows_step, "distinct_packs_per_gpu": DISTINCT_PACKS,
                 "collective": "kx_scan_sharded: ncclAllGather of one 208 B record per rank + device combine on the scan stream" if world > 1 else None},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "packs_per_step": e2e_packs,
                "note": "kx_scan_host: encoded blocks in pinned host memory, bitsets + counts returned to pinned host memory; PCIe-bound "
                        "(and host-memory / root-complex bound when several GPUs of one box pull at once)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                     "kernel": "kx::scan_kernel", "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes, "peak_kind": peak_kind,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "roofline_bitset": {"achieved": achieved_bits, "unit": "GB/s", "frac": achieved_bits / peak, "kernel_ms": kb_ms,
                            "algorithmic_bytes": step_packs * PACK_ROWS * (W_BITS + 1) / 8},
        "clocks": clocks, "matches_per_step": matches, "host_overhead_ms_per_step": 1e3 * wall / args.steps - k_ms,
    }
    if strong:
        line["strong"] = strong
    if rank == 0 and world == 1 and not args.no_configs:
        t0 = time.time()
        c3 = config3(ctx, kb, peak)
        line["configs"] = {"c1": config1(ctx, kb, peak), "c2_widths": config2_widths(ctx, kb, peak), "c3_north_star": c3, "c4": config4(ctx, kb)}
        # the north-star kernel's own roofline: the sparse two-leaf + reduce case (what VERDICT r1 asked to lift)
        ns = c3[0]["int64"]
        t3 = ncu_traffic("r2_ncu_c3dict_traffic.json")
        line["roofline_c3"] = {"bound": "hbm", "kernel": "kx::scan_warp_kernel<1, 512>", "case": c3[0]["case"] + " (int64)", "kernel_ms": ns["kernel_ms"],
                               "achieved": ns["roofline_touched"]["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": ns["roofline_touched"]["frac_of_peak"],
                               "bytes_model": "both filter columns for every row + the 32 B sectors of the value column that hold a match",
                               "traffic": None if not t3 else (t3["dram_bytes_read"] + t3["dram_bytes_write"]) * (c3[0]["rows_per_launch"] / t3["rows"]),
                               "traffic_note": None if not t3 else f"ncu dram__bytes of a {t3['launch']} scaled by rows ({t3['source']})",
                               "frac_survey_8d_bytes": ns["roofline_survey_8d"]["frac_of_peak"]}
        line["configs_wall_s"] = time.time() - t0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as ko
        ncores = os.cpu_count() or 1
        simd = bool(ko.lib().ko_simd_available())
        cpu_payloads = [p[: nbytes + 64] for p in payloads]
        v, dt, reps, _, cbits = cpu_scan(cpu_payloads, thr_field, ncores, 8.0)
        # parity of the CPU port and the GPU on the same packed bytes
        assert (cbits[0][: PACK_ROWS // 8] == r["bitsets"][0]).all(), "oracle and GPU bitsets differ"
        v1, _, _, _, _ = cpu_scan(cpu_payloads[:16], thr_field, 1, 3.0)
        vs, _, _, _, sbits = cpu_scan(cpu_payloads, thr_field, ncores, 5.0, scalar=True)
        vs1, _, _, _, _ = cpu_scan(cpu_payloads[:4], thr_field, 1, 3.0, scalar=True)
        assert (sbits[0] == cbits[0]).all(), "scalar and vector CPU ports differ"
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": ncores, "kind": "port", "simd": simd, "cpu": cpu_info(),
                                "sample": f"{DISTINCT_PACKS} distinct packs x {PACK_ROWS} rows ({DISTINCT_PACKS * nbytes / 2**30:.2f} GiB) x {reps} reps in {dt:.1f} s, "
                                          f"{'AVX-512 VBMI kernel (oracle/ko_simd.c: same bitset words as' if simd else 'scalar C port ('} bitpack.Less) + popcount, {ncores} pthreads",
                                "single_thread_value": v1, "scalar_port_value": vs, "scalar_port_single_thread_value": vs1,
                                "reference_published": "2.5-2.8 values/ns/core, generated scalar Go, 64 K-row inputs, i9-12900K (internal/encode/bitpack/bench.md:72-74)"}
    if rank == 0:
        print(json.dumps(line))
    prog.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
