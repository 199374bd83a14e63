#!/usr/bin/env python3
"""bench.py — scan-filter-aggregate throughput of the KnoxDB pack scan on B200.

Workload (BASELINE.json configs[1]): bit-packed uint64 column (BitpackContainer, min-FOR,
width w = 20 bits), packs of 4 Mi rows, predicate `Less(median)` → LSB-first bitset per pack +
match count, fused decode+compare+popcount in one kernel launch over all packs.

One "step" = one pass over all resident packs of this rank.  Lines printed (one JSON object):
  value      rows/s with the packs resident in HBM (kernel + launch + result copies of counts)
  e2e        rows/s through kx_scan_host: encoded blocks in pinned HOST memory, H2D copies of
             the blocks and D2H of bitsets + counts inside the timed region
  roofline   algorithmic bytes of the scan kernel ÷ its CUDA-event duration vs measured HBM peak
  cpu_baseline  the oracle's C port of the reference's fused bitpack compare on the host cores

`--impl reference` times that CPU port (the reference's Go path cannot be built here: no Go
toolchain) on the same workload shape, bounded sample per step.
"""
import argparse
import json
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


def build_pack_payloads(n_distinct, seed=1):
    """Random packed payloads: any bit string is a valid stream of uniform w-bit fields."""
    rng = np.random.default_rng(seed)
    nbytes = PACK_ROWS * W_BITS // 8
    return [rng.integers(0, 256, nbytes, dtype=np.uint8) for _ in range(n_distinct)]


def encode_block(payload):
    """[IntBitpacked=4][uvarint For][uvarint Log2][uvarint N][packed] — int_bitpack.go:92-98"""
    def uv(x):  # pkg/num/varint.go PutUvarint
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
    hdr = bytes([4]) + uv(FOR_BASE) + uv(W_BITS) + uv(PACK_ROWS)
    return np.concatenate([np.frombuffer(hdr, dtype=np.uint8), payload])


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


def cpu_baseline(payloads, threshold, nthreads, target_s, npacks_sample):
    """oracle port of bitpack.Less over packed words + popcount on the host cores"""
    import oracle as ko
    import ctypes as C
    L = ko.lib()
    arrs = [np.ascontiguousarray(payloads[i % len(payloads)]).view(np.uint64) for i in range(npacks_sample)]
    bits = [np.zeros(PACK_ROWS // 8 + 8, dtype=np.uint8) for _ in range(npacks_sample)]
    pp = (C.c_void_p * npacks_sample)(*[a.ctypes.data for a in arrs])
    bp = (C.c_void_p * npacks_sample)(*[b.ctypes.data for b in bits])
    nr = (C.c_size_t * npacks_sample)(*([PACK_ROWS] * npacks_sample))
    L.ko_baseline_bitpack_scan(pp, nr, npacks_sample, W_BITS, ko.LT, threshold, 0, bp, nthreads)   # warm-up
    t0 = time.perf_counter(); reps = 0; total = 0
    while True:
        total = L.ko_baseline_bitpack_scan(pp, nr, npacks_sample, W_BITS, ko.LT, threshold, 0, bp, nthreads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or reps >= 4096:
            break
    rows = reps * npacks_sample * PACK_ROWS
    return rows / dt, dt, reps, int(total), bits


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (C port in oracle/, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    payloads = build_pack_payloads(4)
    thr = 1 << (W_BITS - 1)
    sample = max(ncores, 16)
    rates = []
    for _ in range(args.warmup):
        cpu_baseline(payloads, thr, ncores, 0.0, sample)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, dt, reps, _, _ = cpu_baseline(payloads, thr, ncores, 0.0, sample)
        rates.append((sample * PACK_ROWS * reps, dt))
    rows = sum(r for r, _ in rates); secs = sum(d for _, d in rates)
    v = rows / secs
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"bitpacked uint64 w={W_BITS} min-FOR, {PACK_ROWS}-row packs, Less(median) -> bitset + count", "pack_rows": PACK_ROWS,
                       "width_bits": W_BITS, "sample_packs_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": f"{sample} packs x {PACK_ROWS} rows per step, C port of bitpack.Less+popcount (oracle/), {ncores} pthreads; Go toolchain absent so the reference itself cannot run"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--packs", type=int, default=1024, help="resident packs per GPU (4 Mi rows each; 1024 packs = 10.7 GB packed)")
    ap.add_argument("--e2e-packs", type=int, default=64, help="packs per e2e step (host-resident blocks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import knoxdb_b200 as kb

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
    npacks = args.packs
    payloads = build_pack_payloads(8, seed=1 + rank)
    pinned = []
    for pl in payloads:                       # encoded blocks live in pinned host memory
        enc = encode_block(pl)
        buf = ctx.host_array(enc.size)
        buf[:] = enc
        pinned.append(buf)
    for p in range(npacks):
        assert ctx.block_put(p, 1, FIELD, kb.UINT64, pinned[p % len(pinned)]) == PACK_ROWS
    thr_field = 1 << (W_BITS - 1)             # median of the uniform w-bit fields
    prog = kb.Program(ctx, [kb.Leaf(FIELD, kb.UINT64, kb.LT, FOR_BASE + thr_field)])
    packs = ctx.pack_refs([(p, 1) for p in range(npacks)])   # kx_packref[] built once, like a Go caller would
    nrows = [PACK_ROWS] * npacks
    offs, total_bits = ctx.bitset_layout(nrows)
    bitbuf = ctx.host_array(total_bits)       # pinned result buffer

    # ---- parity spot check against numpy truth on one pack (full check lives in tests/)
    r = ctx.scan(prog, [(0, 1), (1, 1)], nrows=nrows[:2], want_bitsets=True)
    words = payloads[0].view(np.uint64)
    sample_rows = 100_000
    bitoff = np.arange(sample_rows, dtype=np.uint64) * np.uint64(W_BITS)
    idx = (bitoff >> np.uint64(6)).astype(np.int64); sh = bitoff & np.uint64(63)
    lo = words[idx] >> sh
    hi = np.where(sh > 0, words[np.minimum(idx + 1, words.size - 1)] << ((np.uint64(64) - sh) & np.uint64(63)), np.uint64(0))
    fields = (lo | hi) & np.uint64((1 << W_BITS) - 1)
    truth = np.packbits((fields < np.uint64(thr_field)).astype(np.uint8), bitorder="little")
    assert (r["bitsets"][0][: sample_rows // 8] == truth[: sample_rows // 8]).all(), "GPU scan disagrees with numpy truth"

    from knoxdb_b200 import shard
    from knoxdb_b200.lib import AggOut
    dev = torch.device("cuda", local)
    exch = shard.PartialExchange(1, dist, dev) if world > 1 else None

    def step_resident():
        res = ctx.scan(prog, packs, nrows=nrows, want_bitsets=False)     # counts only: bitsets stay in HBM
        st = ctx.last_scan_stats()
        total = int(res["counts"].sum())
        if world > 1:
            # ONE small NCCL collective per query: all-gather of the 64 B per-rank partial, combined in
            # rank order through kx_agg_combine (the same path sum/min/max partials take)
            mine = AggOut(); mine.count = total; mine.sum_bits = total; mine.min_bits = total; mine.max_bits = total; mine.valid = 1
            total = int(exch.exchange([mine], [kb.UINT64])[0].sum_bits)
        return st, total

    # the headline kernel writes bitsets too; kx_scan(bitsets=…) would also copy them to the host, so
    # for the HBM-resident number the bitsets are produced into the device buffer and only counts return.
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
    rows_step = npacks * PACK_ROWS * world
    value = rows_step * args.steps / wall

    # ---- roofline of the dominant kernel (scan_kernel): algorithmic bytes / CUDA-event time.
    # count-only launch: reads w/8 B per row, writes nothing per row
    alg_bytes = npacks * PACK_ROWS * W_BITS / 8
    k_ms = kernel_ms / args.steps
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    peak, peak_kind = measured_peak()

    # bitset-materialising variant (w/8 + 1/8 B per row) timed separately for the record
    for _ in range(2):
        step_resident_bits()
    kb_ms = 0.0
    nb = max(3, min(args.steps, 5))
    for _ in range(nb):
        st, _ = step_resident_bits()
        kb_ms += st["kernel_ms"]
    kb_ms /= nb
    achieved_bits = npacks * PACK_ROWS * (W_BITS + 1) / 8 / (kb_ms * 1e-3) / 1e9

    # ---- e2e through the C ABI with HOST blocks (pinned): H2D of blocks + D2H of bitsets/counts per step
    e2e_packs = min(args.e2e_packs, npacks)
    hb = [[pinned[p % len(pinned)]] for p in range(e2e_packs)]
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

    # DRAM traffic of the same kernel from the committed ncu capture (a 256-pack launch), scaled to this launch's packs
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_w20_traffic.json")))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * (alg_bytes / tj["algorithmic_bytes"])
        traffic_note = f"ncu dram__bytes_read+write of a {tj['launch']} scaled by packs ({tj['source']})"
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"bitpacked uint64 w={W_BITS} min-FOR, {PACK_ROWS}-row packs, Less(median) -> match count (bitset variant in 'roofline_bitset')",
                   "pack_rows": PACK_ROWS, "width_bits": W_BITS, "packs_per_gpu": npacks, "rows_per_step": rows_step,
                   "l2_policy": f"inputs larger than L2: {alg_bytes / 1e6:.0f} MB packed per step vs 126 MB L2", "sharding": "packs, no data-path collective; one NCCL all_gather of a 64 B partial per query"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "packs_per_step": e2e_packs,
                "note": "kx_scan_host: encoded blocks in pinned host memory, bitsets + counts returned to pinned host memory"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                     "kernel": "kx::scan_kernel", "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes, "peak_kind": peak_kind,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "roofline_bitset": {"achieved": achieved_bits, "unit": "GB/s", "frac": achieved_bits / peak, "kernel_ms": kb_ms,
                            "algorithmic_bytes": npacks * PACK_ROWS * (W_BITS + 1) / 8},
        "clocks": clocks, "matches_per_step": matches, "host_overhead_ms_per_step": 1e3 * wall / args.steps - k_ms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncores = os.cpu_count() or 1
        v, dt, reps, _, cbits = cpu_baseline(payloads, thr_field, ncores, 10.0, max(ncores, 16))
        # parity of the CPU port and the GPU on the same packed bytes
        assert (cbits[0][: PACK_ROWS // 8] == r["bitsets"][0]).all(), "oracle and GPU bitsets differ"
        v1, dt1, reps1, _, _ = cpu_baseline(payloads, thr_field, 1, 5.0, 4)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
                                "sample": f"{max(ncores, 16)} packs x {PACK_ROWS} rows x {reps} reps in {dt:.1f} s, C port of bitpack.Less + popcount (oracle/), {ncores} pthreads",
                                "single_thread_value": v1}
    if rank == 0:
        print(json.dumps(line))
    prog.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
