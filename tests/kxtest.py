"""Shared helpers for the test suite: synthetic column shapes, encoders (via the oracle's
Store functions), numpy truth predicates, and the CPU-only host-translation harness."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS_DIR = os.path.join(ROOT, "tests", "harness")
HARNESS_SO = os.path.join(HARNESS_DIR, "libkx_host_harness.so")

OPS = {ko.EQ: lambda v, a, b: v == a, ko.NE: lambda v, a, b: v != a, ko.LT: lambda v, a, b: v < a,
       ko.LE: lambda v, a, b: v <= a, ko.GT: lambda v, a, b: v > a, ko.GE: lambda v, a, b: v >= a,
       ko.RG: lambda v, a, b: (v >= a) & (v <= b)}
INT_TYPES = [ko.I64, ko.U64, ko.I32, ko.U32, ko.I16, ko.U16, ko.I8, ko.U8]


def pack_bits(mask):
    return np.packbits(np.asarray(mask).astype(np.uint8), bitorder="little")


def unpack_bits(bits, n):
    return np.unpackbits(np.asarray(bits, dtype=np.uint8), bitorder="little")[:n].astype(bool)


def rnd_bits(rng, n, w, dtype=np.uint64):
    if w == 0:
        return np.zeros(n, dtype=dtype)
    v = (rng.integers(0, 2**32, n, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 2**32, n, dtype=np.uint64)
    if w < 64:
        v &= np.uint64((1 << w) - 1)
    return v.astype(dtype)


def typed_rand(rng, t, n, lo=None, hi=None):
    info = np.iinfo(ko.NP[t])
    lo = info.min if lo is None else max(lo, info.min)
    hi = info.max if hi is None else min(hi, info.max)
    return rng.integers(lo, hi, n, dtype=ko.NP[t], endpoint=True)


def shapes(rng, t, n):
    """named shapes of the reference's container tests (internal/encode/tests/tests.go:26-41)"""
    info = np.iinfo(ko.NP[t])
    span = min(info.max, 2**20)
    base = int(typed_rand(rng, t, 1, 0, min(info.max // 2, 1000))[0])
    out = {
        "rnd": typed_rand(rng, t, n, -span, span),
        "small": typed_rand(rng, t, n, 0, 100),
        "dups": rng.choice(typed_rand(rng, t, 7, -span, span), n),
        "runs": np.repeat(typed_rand(rng, t, (n + 4) // 5, -span, span), 5)[:n],
        "const": np.full(n, base, dtype=ko.NP[t]),
    }
    if n * 3 + base < info.max:
        out["delta+"] = (base + 3 * np.arange(n)).astype(ko.NP[t])
    if info.min < 0 and n * 2 < info.max:
        out["delta-"] = (base - 2 * np.arange(n)).astype(ko.NP[t])
        out["neg"] = typed_rand(rng, t, n, -span, -1)
    out["edge"] = np.resize(np.array([info.min, info.max, 0, 1, info.max - 1, info.min + 1], dtype=ko.NP[t]), n)
    return out


def container_kinds(t, vals):
    """container schemes that may legally hold `vals` (see tests/test_oracle_props.py)"""
    n = vals.size
    kinds = ["raw", "bitpack", "best"]
    if n >= 2:
        kinds += ["dict", "runend"]
    bits_t = np.iinfo(ko.NP[t]).bits
    rng_ = int(vals.max()) - int(vals.min())
    if rng_ >= 2 ** (bits_t - 1) and bits_t < 64 and t <= ko.I8:
        kinds.remove("bitpack")   # ineligible at full width (context.go:266-269)
    if bits_t == 64 and rng_ < 2**60:
        kinds.append("s8b")
    return kinds


def operands(t, vals):
    info = np.iinfo(ko.NP[t])
    n = vals.size
    picks = {int(vals[0]), int(vals[n // 2]), int(vals.min()), int(vals.max()), 0}
    picks |= {max(info.min, int(vals.min()) - 1), min(info.max, int(vals.max()) + 1)}
    return sorted(picks)


# ------------------------------------------------------------------ CPU harness over kx_host.cpp
_h = None


def harness():
    global _h
    if _h is None:
        srcs = [os.path.join(HARNESS_DIR, "host_harness.cpp"), os.path.join(ROOT, "knoxdb_b200", "csrc", "kx_host.cpp")]
        deps = srcs + [os.path.join(ROOT, "knoxdb_b200", "csrc", f) for f in ("kx_host.h", "kx_types.h")]
        if not os.path.exists(HARNESS_SO) or any(os.path.getmtime(d) > os.path.getmtime(HARNESS_SO) for d in deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", HARNESS_SO] + srcs)
        L = C.CDLL(HARNESS_SO)
        vp = C.c_void_p
        L.kxh_match.restype = C.c_long
        L.kxh_match.argtypes = [C.c_int, vp, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64, vp, C.c_uint32, vp, C.POINTER(C.c_int)]
        L.kxh_decode.restype = C.c_long
        L.kxh_decode.argtypes = [C.c_int, vp, C.c_size_t, vp, C.c_size_t]
        L.kxh_view_kind.restype = C.c_int
        L.kxh_view_kind.argtypes = [C.c_int, vp, C.c_size_t]
        L.kxh_xxh3_fixed.restype = C.c_uint64
        L.kxh_xxh3_fixed.argtypes = [C.c_int, C.c_uint64]
        L.kxh_xxh3_bytes.restype = C.c_uint64
        L.kxh_xxh3_bytes.argtypes = [vp, C.c_size_t]
        L.kxh_str_match.restype = C.c_long
        L.kxh_str_match.argtypes = [vp, C.c_size_t, C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp]
        _h = L
    return _h


def host_str_match(blob, n, op, a, b=b""):
    """the product's host normalisation of a string block + its scalar string predicate, row by row (CPU)"""
    enc = np.frombuffer(blob, dtype=np.uint8).copy()
    bits = np.zeros((n + 7) // 8 + 8, dtype=np.uint8)
    aa, bb = np.frombuffer(a + b"\0", dtype=np.uint8).copy(), np.frombuffer(b + b"\0", dtype=np.uint8).copy()
    rows = harness().kxh_str_match(enc.ctypes.data, enc.size, op, aa.ctypes.data, len(a), bb.ctypes.data, len(b), bits.ctypes.data)
    assert rows == n, rows
    return bits[: (n + 7) // 8]


def host_match(t, blob, n, op, a=0, b=0, values=None):
    enc = np.frombuffer(blob, dtype=np.uint8).copy()
    bits = np.zeros((n + 7) // 8 + 8, dtype=np.uint8)
    s = None if values is None else np.ascontiguousarray(values, dtype=np.uint64)
    mode = C.c_int()
    rows = harness().kxh_match(t, enc.ctypes.data, enc.size, op, a & (2**64 - 1), b & (2**64 - 1),
                               None if s is None else s.ctypes.data, 0 if s is None else s.size, bits.ctypes.data, C.byref(mode))
    assert rows == n, rows
    return bits[:(n + 7) // 8], mode.value
