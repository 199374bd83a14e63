"""CPU tier: the product's HOST logic (knoxdb_b200/csrc/kx_host.cpp — container parsing,
normalisation into device views, leaf → per-pack predicate translation) against the oracle.
The device semantics are emulated by tests/harness/host_harness.cpp (test infrastructure);
the CUDA kernels are covered by the -m gpu tests."""
import numpy as np
import pytest

import kxtest as kt
import oracle as ko

RNG = np.random.default_rng(42)


@pytest.mark.parametrize("t", kt.INT_TYPES)
def test_leaf_translation_matches_oracle(t):
    modes_seen = set()
    for n in (1, 3, 64, 67, 640, 1025):
        for name, vals in kt.shapes(RNG, t, n).items():
            for kind in kt.container_kinds(t, vals):
                blob = ko.store(kind, t, vals)
                oc = ko.Container(t, blob)
                # decode parity
                out = np.zeros(n, dtype=np.uint64)
                assert kt.harness().kxh_decode(t, np.frombuffer(blob, np.uint8).copy().ctypes.data, len(blob), out.ctypes.data, n) == n
                assert (out == oc.decode()).all(), (name, kind)
                for a in kt.operands(t, vals):
                    b = min(np.iinfo(ko.NP[t]).max, a + 5)
                    for op in kt.OPS:
                        want = oc.match(op, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                        got, mode = kt.host_match(t, blob, n, op, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                        modes_seen.add(mode)
                        # (run-end blocks whose Values child is affine inherit DeltaContainer.MatchBetween's rounding quirk
                        # in the reference: the product matches such blocks with the same index arithmetic over the runs)
                        if not (got == want).all():
                            raise AssertionError((ko.NP[t].__name__, n, name, kind, op, a, b, mode))
                setv = np.unique(np.concatenate([vals[: min(3, n)], kt.typed_rand(RNG, t, 3)]))
                su = ko.as_u64(t, setv)
                for neg, op in ((False, ko.IN), (True, ko.NI)):
                    got, _ = kt.host_match(t, blob, n, op, values=su)
                    assert (got == oc.match_set(su, negate=neg)).all(), (name, kind, "set", neg)
    assert len(modes_seen) >= 4


def test_delta_quirk_is_reproduced():
    """The reference's DeltaContainer.MatchBetween rounding below For is reproduced bit for bit."""
    blob = ko.store("delta", ko.I64, base=100, delta=10, n=8)
    got, _ = kt.host_match(ko.I64, blob, 8, ko.RG, 95, 135)
    assert got.tolist() == ko.Container(ko.I64, blob).match(ko.RG, 95, 135).tolist() == [0b00001110]


def test_delta_quirk_is_reproduced_under_run_end_blocks():
    """RunEndContainer.Match* hands the predicate to its Values child (int_runend.go:224-283); an affine child answers with
    DeltaContainer's index arithmetic over the RUNS, quirk included: run values 100, 110, … with ranges that start below For
    and between grid points, for every operator (the encoder picks a delta child for positive steps only)."""
    for t, base, delta in ((ko.I64, 100, 10), (ko.U64, 1000, 3), (ko.I32, -50, 4), (ko.I64, -10**12, 977)):
        nruns = 23
        runs = (base + delta * np.arange(nruns)).astype(ko.NP[t])
        lens = RNG.integers(1, 9, nruns)
        vals = np.repeat(runs, lens)
        blob = ko.store("runend", t, vals)
        oc = ko.Container(t, blob)
        assert oc.ctype == ko.TRUNEND and oc.value_delta_sequences(), "the encoder must pick an affine Values child here"
        lo_all, hi_all = int(runs.min()), int(runs.max())
        quirk_hits = 0
        for a in range(lo_all - 2 * abs(delta) - 1, hi_all + 2 * abs(delta) + 2, 3):
            if t == ko.U64 and a < 0:
                continue
            for b in (a, a + 5, a + 37, hi_all + 50):
                for op in kt.OPS:
                    want = oc.match(op, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                    got, _ = kt.host_match(t, blob, len(vals), op, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                    assert (got == want).all(), (ko.NP[t].__name__, delta, op, a, b)
                    if op == ko.RG and not (want == kt.pack_bits(kt.OPS[op](vals, ko.NP[t](a), ko.NP[t](min(b, np.iinfo(ko.NP[t]).max))))).all():
                        quirk_hits += 1
        assert quirk_hits > 0   # the sweep really entered the quirk domain


@pytest.mark.parametrize("w", [0, 1, 7, 8, 13, 20, 31, 32, 33, 47, 63, 64])
def test_bitpack_range_translation_edges(w):
    """operands at / beyond the field mask, inverted ranges, 32- vs 64-bit path selection"""
    t = ko.U64
    n = 333
    vals = kt.rnd_bits(RNG, n, w)
    vals[:2] = [0, (1 << w) - 1 if w < 64 else 2**64 - 1]
    base = 1000 if w < 60 else 0
    col = vals + np.uint64(base)
    blob = ko.store("bitpack", t, col)
    oc = ko.Container(t, blob)
    top = (1 << w) - 1 if w < 64 else 2**64 - 1
    cands = [0, 1, base, base + 1, base + top, min(base + top + 1, 2**64 - 1), 2**32 - 1, 2**32, 2**63, 2**64 - 1, int(col[5])]
    for a in cands:
        for op in kt.OPS:
            for b in (a, min(a + 1000, 2**64 - 1), 2**64 - 1, 0):
                want = oc.match(op, a, b)
                got, _ = kt.host_match(t, blob, n, op, a, b)
                assert (got == want).all(), (w, op, a, b)
                if op != ko.RG:
                    break


def test_float_translation():
    for t, dt in ((ko.F64, np.float64), (ko.F32, np.float32)):
        vals = (RNG.integers(0, 2**30, 500) / 100.0).astype(dt)
        vals[::31] = np.nan
        vals[3] = np.inf
        blob = ko.store("raw", t, vals)
        oc = ko.Container(t, blob)
        for a, b in ((vals[10], vals[10] * 2), (np.nan, 1.0), (-np.inf, np.inf)):
            for op in kt.OPS:
                ua, ub = ko.scalar_u64(t, dt(a)), ko.scalar_u64(t, dt(b))
                got, _ = kt.host_match(t, blob, 500, op, ua, ub)
                assert (got == oc.match(op, ua, ub)).all(), (t, op, a, b)


def test_xxh3_host_matches_oracle_and_golden():
    import json
    import os
    x = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "xxh3_vectors.json")))
    for inp, r32, r64 in zip(x["input_bytes"], x["u32"], x["u64"]):
        b = np.array(inp, dtype=np.uint8)
        assert kt.harness().kxh_xxh3_bytes(b.ctypes.data, 4) == r32
        assert kt.harness().kxh_xxh3_bytes(b.ctypes.data, 8) == r64
    for n in list(range(0, 300, 7)) + [1023, 1024, 1025, 5000]:
        b = RNG.integers(0, 256, max(n, 1), dtype=np.uint8)
        assert kt.harness().kxh_xxh3_bytes(b.ctypes.data, n) == ko.lib().ko_xxh3_bytes(b.ctypes.data, n), n


def test_xxh3_fixed_width_specialisations():
    """hash.Uint64/Uint32/Uint16/Uint8 (internal/hash/xxh3.go:22-58) = XXH3-64 of the value's LE bytes"""
    import xxhash
    L = ko.lib()
    for v in [0, 1, 0x7F, 0xFF, 0x1234, 0xFFFF, 0xDEADBEEF, 0xFFFFFFFF, 0x0123456789ABCDEF, 2**64 - 1] + [int(x) for x in RNG.integers(0, 2**63, 50)]:
        for nb, ofn in ((8, L.ko_xxh3_u64), (4, L.ko_xxh3_u32), (2, L.ko_xxh3_u16), (1, L.ko_xxh3_u8)):
            x = v & ((1 << (8 * nb)) - 1)
            want = xxhash.xxh3_64_intdigest(x.to_bytes(nb, "little"))
            assert kt.harness().kxh_xxh3_fixed(nb, x) == want == ofn(x), (nb, hex(x))
    for n in list(range(0, 260)) + [300, 511, 512, 513, 1024, 4097]:
        b = RNG.integers(0, 256, max(n, 1), dtype=np.uint8)
        assert kt.harness().kxh_xxh3_bytes(b.ctypes.data, n) == xxhash.xxh3_64_intdigest(b[:n].tobytes()), n


def alp_columns(rng, n):
    """float64 shapes for the ALP container: decimals with a few non-decimal exceptions, specials, all-exception"""
    dec2 = np.round(rng.uniform(-1000, 1000, n), 2)
    mixed = dec2.copy()
    mixed[::37] = rng.uniform(-1, 1, mixed[::37].size)            # not representable with 2 decimals → patches
    special = mixed.copy()
    special[1 % n] = np.nan; special[5 % n] = np.inf; special[7 % n] = -np.inf; special[9 % n] = -0.0
    ints = rng.integers(-50, 50, n).astype(np.float64)
    const = np.full(n, 12.5)
    return {"dec2": dec2, "mixed": mixed, "special": special, "ints": ints, "const": const,
            "allpatch": rng.uniform(0, 1, n)}


def alp_operands(vals):
    fin = vals[np.isfinite(vals)]
    ops = [float(fin[0]), float(fin[fin.size // 2]), float(fin.min()), float(fin.max()), 0.0, 0.005, -3.3333, 1e300, -1e300,
           float(fin[0]) + 1e-9, np.nan, np.inf, -np.inf]
    return ops


def test_alp_host_translation_matches_oracle():
    """FloatAlpContainer.Match* (float_alp.go:238-495): the product's translation into the encoded integer domain +
    patch correction equals the oracle's restatement, which equals the scalar predicate on the original floats
    wherever the reference itself is exact."""
    t = ko.F64
    for n in (1, 33, 640, 5000):
        for name, vals in alp_columns(RNG, n).items():
            for e, f in ((-1, -1), (2, 0), (14, 12)):
                blob = ko.store("alp", t, vals, e=e, f=f)
                oc = ko.Container(t, blob)
                out = np.zeros(n, dtype=np.uint64)
                assert kt.harness().kxh_decode(t, np.frombuffer(blob, np.uint8).copy().ctypes.data, len(blob), out.ctypes.data, n) == n
                assert (out == oc.decode()).all(), (name, e, f)
                # decode is lossless except for the sign of zero (the reference does not special-case -0.0)
                assert (out.view(np.float64)[~np.isnan(vals)] == vals[~np.isnan(vals)]).all()
                for a in alp_operands(vals):
                    for b in (a + 2.5, a, a - 1.0):
                        for op in kt.OPS:
                            ua, ub = ko.scalar_u64(t, a), ko.scalar_u64(t, b)
                            want = oc.match(op, ua, ub)
                            got, _ = kt.host_match(t, blob, n, op, ua, ub)
                            assert (got == want).all(), (n, name, e, f, op, a, b)


def test_string_blocks_host_normalisation_matches_the_oracle():
    """normalize_string_block (containers 16..19 → byte buffer + flat index array) and string_pred, interpreted row by
    row on the CPU, against the oracle's string containers for all seven modes; truncated buffers are rejected."""
    rng = np.random.default_rng(8)
    vocab = [b"", b"a", b"ab", b"abc", b"abd", b"tz1VSUr8wwNhLAzempoch5d6hLRiTh8Cjcjb", b"zz", bytes(range(200))]
    ragged = [vocab[i] for i in rng.integers(0, len(vocab), 3000)]
    fixed = [bytes(r) for r in rng.integers(0, 3, (2000, 4), dtype=np.uint8)]
    cases = [(ko.STR_COMPACT, ragged), (ko.STR_DICT, ragged), (ko.STR_FIXED, fixed), (ko.STR_COMPACT, fixed), (ko.STR_DICT, fixed),
             (ko.STR_CONST, [b"same"] * 100), (ko.STR_COMPACT, [b""]), (ko.STR_DICT, [b"x", b"x", b"y"])]
    modes = ((1, ko.EQ), (2, ko.NE), (3, ko.GT), (4, ko.GE), (5, ko.LT), (6, ko.LE))
    for kind, rows in cases:
        blob = ko.store_str(kind, rows)
        oc = ko.StrContainer(blob)
        for a in (rows[0], rows[len(rows) // 2], b"", b"ab", b"\x01\x01", b"\xff"):
            for m, kom in modes:
                assert (kt.host_str_match(blob, len(rows), m, a) == oc.match(kom, a)).all(), (kind, m, a)
            assert (kt.host_str_match(blob, len(rows), 9, a, a + b"\x01") == oc.match(ko.RG, a, a + b"\x01")).all(), (kind, "range", a)
        enc = np.frombuffer(blob, dtype=np.uint8).copy()
        bits = np.zeros(len(rows) // 8 + 16, dtype=np.uint8)
        aa = np.zeros(4, dtype=np.uint8)
        if len(blob) > 8:
            assert kt.harness().kxh_str_match(enc.ctypes.data, len(blob) - 3, 1, aa.ctypes.data, 1, aa.ctypes.data, 0, bits.ctypes.data) < 0


@pytest.mark.parametrize("t", [ko.U64, ko.I64, ko.I16])
def test_in_set_structures_have_no_false_negatives_or_positives(t):
    """IN / NOT IN on packed and raw blocks goes through two host-built structures: a one-hash prefilter bitmap and the
    bucketised exact table (kx_host.cpp: build_set_prefilter, build_set_table).  The harness applies them in the order
    the device does; set sizes from 1 to 20 000 keys, members and non-members, sequential and clustered keys."""
    rng = np.random.default_rng(13)
    info = np.iinfo(ko.NP[t])
    n = 20_000
    span = min(int(info.max) - 10, 1 << 36)
    vals = rng.integers(0, span, n, dtype=np.int64).astype(ko.NP[t])
    for kind in ("bitpack", "raw"):
        blob = ko.store(kind, t, vals)
        oc = ko.Container(t, blob)
        for nset in (1, 7, 64, 1000, 20_000):
            members = rng.choice(vals, nset // 2 + 1)
            seq = (np.arange(nset // 2 + 1) + int(vals[0])).astype(ko.NP[t])          # sequential keys: worst case for weak hashes
            su = ko.as_u64(t, np.unique(np.concatenate([members, seq])))
            for neg, op in ((False, ko.IN), (True, ko.NI)):
                got, mode = kt.host_match(t, blob, n, op, values=su)
                assert (got == oc.match_set(su, negate=neg)).all(), (kind, nset, neg, mode)


def test_string_block_parser_survives_corrupt_buffers():
    """kx_block_put parses bytes that come from storage: truncations and bit flips must end in an error code or in a block
    whose every row lies inside the byte buffer (kxh_str_match checks that) — never in an out-of-bounds read."""
    rng = np.random.default_rng(17)
    rows = [bytes(rng.integers(97, 123, int(k), dtype=np.uint8)) for k in rng.integers(0, 12, 400)]
    fixed = [bytes(r) for r in rng.integers(0, 256, (300, 6), dtype=np.uint8)]
    H = kt.harness()
    a = np.frombuffer(b"abc\0", dtype=np.uint8).copy()
    for kind, data in ((ko.STR_COMPACT, rows), (ko.STR_DICT, rows), (ko.STR_FIXED, fixed), (ko.STR_CONST, [b"constant"] * 50)):
        blob = np.frombuffer(ko.store_str(kind, data), dtype=np.uint8)
        bits = np.zeros((1 << 26) // 8 + 64, dtype=np.uint8)   # the parser refuses blocks that claim more than 2^26 rows
        for trial in range(300):
            enc = blob.copy()
            if trial % 3 == 0:
                enc = enc[: int(rng.integers(1, enc.size))]                      # truncation
            else:
                for _ in range(int(rng.integers(1, 4))):                         # bit flips, mostly in the headers
                    pos = int(rng.integers(0, min(enc.size, 64 if trial % 3 == 1 else enc.size)))
                    enc[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
            enc = np.ascontiguousarray(enc)
            rc = H.kxh_str_match(enc.ctypes.data, enc.size, 1, a.ctypes.data, 3, a.ctypes.data, 0, bits.ctypes.data) if enc.size else -1
            assert rc < 0 or rc <= (1 << 26), (kind, trial, rc)


def test_integer_block_parser_survives_corrupt_buffers():
    """the same for the integer containers: truncated and bit-flipped blocks (headers and payload) decode to an error
    or to at most `cap` rows, never to an out-of-bounds access"""
    rng = np.random.default_rng(23)
    H = kt.harness()
    n, cap = 3000, 1 << 22
    vals = rng.integers(-1000, 1000, n).astype(np.int64)
    dst = np.zeros(cap, dtype=np.uint64)
    blobs = [ko.store(kind, ko.I64, np.repeat(vals[: n // 5], 5) if kind == "runend" else vals) for kind in ("raw", "bitpack", "dict", "runend", "s8b", "best")]
    blobs += [ko.store("delta", ko.I64, base=5, delta=3, n=1000), ko.store("const", ko.I64, val=7, n=1000), ko.store("alp", ko.F64, np.round(rng.uniform(0, 99, n), 2))]
    for bi, blob in enumerate(blobs):
        blob = np.frombuffer(blob, dtype=np.uint8)
        t = ko.F64 if bi == len(blobs) - 1 else ko.I64
        for trial in range(200):
            enc = blob.copy()
            if trial % 3 == 0 and enc.size > 2:
                enc = enc[: int(rng.integers(1, enc.size))]
            else:
                for _ in range(int(rng.integers(1, 4))):
                    pos = int(rng.integers(0, min(enc.size, 48 if trial % 3 == 1 else enc.size)))
                    enc[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
            enc = np.ascontiguousarray(enc)
            rc = H.kxh_decode(t, enc.ctypes.data, enc.size, dst.ctypes.data, cap)
            assert rc <= cap, (bi, trial, rc)


def _uv(x):
    """pkg/num/varint.go PutUvarint"""
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


def test_crafted_block_headers_are_refused_not_fatal():
    """headers a bit flip cannot reach but a corrupt store can hold: a few bytes that claim billions of rows in a nested
    child (materialised on the host), a chain of nested dictionary ids (recursion), an affine block with a zero delta
    (the matchers divide by it).  All must end in an error code (rc < 0): no allocation of gigabytes, no stack overflow,
    no SIGFPE."""
    H = kt.harness()
    dst = np.zeros(1 << 16, dtype=np.uint64)
    bits = np.zeros((1 << 16) // 8, dtype=np.uint8)
    huge = (1 << 32) - 1
    const_huge = bytes([1]) + _uv(5) + _uv(huge)                     # IntConstant, 4 G rows in 7 bytes
    delta_huge = bytes([2]) + _uv(0) + _uv(1) + _uv(huge)            # IntDelta
    codes_ok = bytes([1]) + _uv(0) + _uv(100)                        # 100 codes, all 0
    crafted = {
        "dict over a 4 G-row constant dictionary": bytes([5]) + const_huge + codes_ok,
        "dict with 4 G constant codes": bytes([5]) + bytes([1]) + _uv(7) + _uv(1) + const_huge,
        "run-end with 4 G affine ends": bytes([3]) + const_huge + delta_huge,
        "nested dictionary ids": bytes([5]) * 100000,
        "nested run-end ids": bytes([3]) * 100000,
        "dict of dict of dict of dict": bytes([5, 5, 5, 5, 5]) + codes_ok * 8,
        "zero delta": bytes([2]) + _uv(10) + _uv(0) + _uv(1000),
        "s8b with 4 G rows and no words": bytes([6]) + _uv(0) + _uv(huge) + _uv(0),
    }
    for name, blob in crafted.items():
        enc = np.frombuffer(blob, dtype=np.uint8).copy()
        rc = H.kxh_decode(ko.I64, enc.ctypes.data, enc.size, dst.ctypes.data, dst.size)
        assert rc < 0, (name, rc)
        rc = H.kxh_match(ko.I64, enc.ctypes.data, enc.size, ko.LT, 5, 0, None, 0, bits.ctypes.data, None)
        assert rc < 0, (name, rc)
    # ALP: 4 G patch positions behind a tiny value stream
    alp = bytes([13]) + _uv(2) + _uv(0) + bytes([1]) + bytes([1]) + _uv(3) + _uv(10) + bytes([10]) + _uv(0) + _uv(huge) + delta_huge
    enc = np.frombuffer(alp, dtype=np.uint8).copy()
    assert H.kxh_decode(ko.F64, enc.ctypes.data, enc.size, dst.ctypes.data, dst.size) < 0
    # a zero-delta block must also be refused on the matcher path the reference would divide in
    enc = np.frombuffer(crafted["zero delta"], dtype=np.uint8).copy()
    for op in (ko.EQ, ko.LT, ko.RG):
        assert H.kxh_match(ko.I64, enc.ctypes.data, enc.size, op, 10, 20, None, 0, bits.ctypes.data, None) < 0
