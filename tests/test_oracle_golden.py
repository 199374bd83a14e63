"""Pin the CPU oracle against the reference's own golden vectors (CPU only).

Fixtures were transcribed (numbers only) from the reference's tests by the scripts in
tests/golden/ — see SURVEY.md §8(c) / Appendix C.
"""
import json
import os

import numpy as np
import pytest

import oracle as ko

G = os.path.join(os.path.dirname(__file__), "golden")
CMP = json.load(open(os.path.join(G, "cmp_vectors.json")))["types"]
BITSET = json.load(open(os.path.join(G, "bitset_vectors.json")))
XXH = json.load(open(os.path.join(G, "xxh3_vectors.json")))


def typed_src(tname, patterns):
    t = ko.TYPE_BY_NAME[tname]
    if t == ko.F64:
        return np.array(patterns, dtype=np.uint64).view(np.float64)
    if t == ko.F32:
        return np.array(patterns, dtype=np.uint32).view(np.float32)
    return np.array(patterns, dtype=np.int64 if t <= ko.I8 else np.uint64).astype(ko.NP[t])


def operand(tname, v):
    t = ko.TYPE_BY_NAME[tname]
    return v & (2**64 - 1) if t not in (ko.F64, ko.F32) else v


@pytest.mark.parametrize("tname", sorted(CMP))
def test_cmp_golden(tname):
    """internal/cmp/tests/<type>.go: expected bitset bytes + count for 7 ops."""
    t = ko.TYPE_BY_NAME[tname]
    ncases = 0
    for opname, cases in CMP[tname].items():
        for c in cases:
            src = typed_src(tname, c["src"])
            bits, cnt = ko.cmp(t, ko.OP_BY_NAME[opname], src, operand(tname, c["a"]), operand(tname, c["b"]))
            assert bits.tobytes().hex() == c["bits"], (tname, opname, c["name"])
            assert cnt == c["count"], (tname, opname, c["name"])
            ncases += 1
    assert ncases >= 100


def test_bitset_popcount_golden():
    """internal/bitset/tests/pop.go:19-47: dirty tail bits must not be counted."""
    for c in BITSET["pop"]:
        buf = np.frombuffer(bytes.fromhex(c["source"]), dtype=np.uint8).copy()
        assert ko.lib().ko_bitset_popcount(ko._p(buf), c["size"]) == c["count"], c["name"]


def test_bitset_indexes_golden():
    """internal/bitset/tests/run.go:32-520: bitset → ascending row ids."""
    for c in BITSET["index"]:
        buf = np.frombuffer(bytes.fromhex(c["buf"]), dtype=np.uint8).copy()
        out = np.zeros(c["size"] + 8, dtype=np.uint32)
        n = ko.lib().ko_bitset_indexes(ko._p(buf) if buf.size else None, c["size"], ko._p(out))
        assert out[:n].tolist() == c["idx"], c["name"]


def test_xxh3_golden():
    """internal/hash/xxh3_test.go:14-31."""
    for inp, r32, r64 in zip(XXH["input_bytes"], XXH["u32"], XXH["u64"]):
        b = bytes(inp)
        assert ko.lib().ko_xxh3_u32(int.from_bytes(b[:4], "little")) == r32
        assert ko.lib().ko_xxh3_u64(int.from_bytes(b, "little")) == r64
        # the closed forms must agree with the generic byte-string hash
        assert ko.lib().ko_xxh3_bytes(ko._p(np.frombuffer(b[:4], dtype=np.uint8).copy()), 4) == r32
        assert ko.lib().ko_xxh3_bytes(ko._p(np.frombuffer(b, dtype=np.uint8).copy()), 8) == r64


def test_xxh3_bytes_vs_python_xxhash():
    """zeebo/xxh3 is canonical XXH3_64bits(seed 0); python-xxhash is an independent witness."""
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(7)
    for n in list(range(0, 260)) + [511, 1024, 1025, 4096, 10000]:
        b = rng.integers(0, 256, n, dtype=np.uint8)
        assert ko.lib().ko_xxh3_bytes(ko._p(b) if n else None, n) == xxhash.xxh3_64_intdigest(b.tobytes()), n
    for v in (0, 1, 255):
        assert ko.lib().ko_xxh3_u8(v) == xxhash.xxh3_64_intdigest(bytes([v]))
    for v in (0, 1, 0xbeef, 0xffff):
        assert ko.lib().ko_xxh3_u16(v) == xxhash.xxh3_64_intdigest(v.to_bytes(2, "little"))


def test_varint_roundtrip():
    """pkg/num/varint.go: boundaries of every length class."""
    edges = [0, 1, 240, 241, 2287, 2288, 67823, 67824, 2**24 - 1, 2**24, 2**32 - 1, 2**32, 2**40 - 1, 2**40,
             2**48 - 1, 2**48, 2**56 - 1, 2**56, 2**64 - 1]
    lens = [1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9]
    import ctypes as C
    for x, l in zip(edges, lens):
        buf = np.zeros(16, dtype=np.uint8)
        n = ko.lib().ko_put_uvarint(ko._p(buf), x)
        assert n == l, (x, n)
        v = C.c_uint64()
        assert ko.lib().ko_uvarint(ko._p(buf), C.byref(v)) == l and v.value == x
    # known encodings (SQLite4 varint): 241 → f1 01 ; 2288 → f9 00 00
    buf = np.zeros(16, dtype=np.uint8)
    ko.lib().ko_put_uvarint(ko._p(buf), 241); assert buf[:2].tolist() == [241, 1]
    ko.lib().ko_put_uvarint(ko._p(buf), 2288); assert buf[:3].tolist() == [249, 0, 0]


def test_window_edges_match_the_reference_timeunit_next_table():
    """ko_window_edges restates TimeUnit.Next for fixed-duration units; the reference pins Next with a table
    (pkg/util/timeunit_test.go:162-228, transcribed by tests/golden/extract_timeunit_vectors.py): the second edge of a
    walk that starts at `in` is Next(in, 1)."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "timeunit_vectors.json")))
    assert len(d["cases"]) >= 15
    for c in d["cases"]:
        e = ko.window_edges(c["in"], c["in"] + 1, c["step_s"])
        assert int(e[0]) == c["in"] and int(e[1]) == c["next"], c
        # and the walk continues in whole steps from there (TestSteps: consecutive steps are one Duration apart)
        e = ko.window_edges(c["in"], c["in"] + 5 * c["step_s"], c["step_s"])
        assert all(int(b) - int(a) == c["step_s"] for a, b in zip(e[1:-1], e[2:]))
