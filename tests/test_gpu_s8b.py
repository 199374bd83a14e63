"""Simple8b blocks (container 6, internal/encode/int_s8b.go:26-235, s8b/generic/decode.go:15-81): the product transcodes
the codewords ON THE DEVICE at kx_block_put (selector counts → exclusive scan → fixed-width stream); decode and every
predicate must equal the oracle's Simple8Container on streams that use every selector."""
import numpy as np
import pytest

import kxtest as kt
import oracle as ko

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(68)


@pytest.fixture(scope="module")
def ctx():
    import knoxdb_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def _mixed_selectors(n):
    """runs of zeros / ones (selectors 0 and 1: 128 values per word) and values of every selector width"""
    parts = [np.zeros(300, np.int64), np.ones(260, np.int64)]
    for bits in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60):
        parts.append(RNG.integers(0, 1 << bits, 97, dtype=np.int64))
    v = np.concatenate(parts)
    reps = -(-n // v.size)
    return np.tile(v, reps)[:n]


@pytest.mark.parametrize("t", [ko.U64, ko.I64, ko.U32])
def test_simple8b_device_transcode_decode_and_match(ctx, t):
    for n in (1, 127, 128, 129, 5000, 70_001):
        vals = _mixed_selectors(n)
        if t == ko.U32:
            vals = vals & 0x3fffffff
        vals = (vals + (1000 if t != ko.I64 else -1000)).astype(ko.NP[t])   # a non-zero For, negative for signed
        blob = ko.store("s8b", t, vals)
        oc = ko.Container(t, blob)
        assert oc.ctype == 6
        assert (ctx.container_decode(t, blob, n) == vals).all(), n
        for a in kt.operands(t, vals)[:4]:
            b = min(np.iinfo(ko.NP[t]).max, a + 40)
            for op in kt.OPS:
                want = oc.match(op, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                got, cnt = ctx.container_match(t, blob, op, a, b, nrows=n)
                assert (got == want).all(), (n, op, a, b)
                assert cnt == int(np.unpackbits(got).sum())


def _uv(x):
    """num.PutUvarint (pkg/num/varint.go:85-192) through the oracle"""
    import ctypes as C
    buf = (C.c_uint8 * 16)()
    k = ko.lib().ko_put_uvarint(buf, C.c_uint64(x))
    return bytes(buf[:k])


def _read_uv(buf, pos):
    import ctypes as C
    v = C.c_uint64()
    raw = (C.c_uint8 * 16)(*buf[pos:pos + 16].ljust(16, b"\0"))
    k = ko.lib().ko_uvarint(raw, C.byref(v))
    assert k > 0
    return v.value, pos + k


def test_simple8b_short_stream_is_refused(ctx):
    """a header that promises more rows than the codewords hold (int_s8b.go:96-115 would read past the stream): the
    device-side count of the selectors finds it"""
    import knoxdb_b200 as kb
    vals = np.arange(1000, dtype=np.uint64) % 7
    good = bytes(ko.store("s8b", ko.U64, vals))
    assert (ctx.container_decode(ko.U64, good, 1000) == vals).all()
    assert good[0] == 6
    minv, pos = _read_uv(good, 1)
    n, pos = _read_uv(good, pos)
    ln, pos = _read_uv(good, pos)
    assert n == 1000 and ln == len(good) - pos and ln % 8 == 0
    short = bytes([6]) + _uv(minv) + _uv(n) + _uv(ln - 8) + good[pos:-8]   # consistent header, one codeword less
    with pytest.raises(kb.KnoxError):
        ctx.container_decode(ko.U64, short, 1000)
    with pytest.raises(kb.KnoxError):
        ctx.block_put(7, 1, 1, kb.UINT64, np.frombuffer(short, dtype=np.uint8))
    # the store is unchanged by the failed put, and a good block still registers afterwards
    assert ctx.block_put(7, 1, 1, kb.UINT64, np.frombuffer(good, dtype=np.uint8)) == 1000
