"""Error behaviour of the C ABI on bad inputs that used to reach the device unchecked (round-1 review): selection ids
outside a pack, and block replacement that fails half way."""
import numpy as np
import pytest

import oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import knoxdb_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def test_gather_refuses_row_ids_outside_the_pack(ctx):
    import knoxdb_b200 as kb
    vals = np.arange(1000, dtype=np.int64) * 3
    assert ctx.block_put(901, 1, 1, kb.INT64, ko.store("best", ko.I64, vals)) == 1000
    good = ctx.gather([(901, 1)], 1, kb.INT64, np.array([0, 7, 999], dtype=np.uint32), np.array([0, 3], dtype=np.uint64))
    assert good.tolist() == [0, 21, 2997]
    with pytest.raises(kb.KnoxError):
        ctx.gather([(901, 1)], 1, kb.INT64, np.array([0, 1000], dtype=np.uint32), np.array([0, 2], dtype=np.uint64))
    with pytest.raises(kb.KnoxError):
        ctx.gather([(901, 1)], 1, kb.INT64, np.array([0xffffffff], dtype=np.uint32), np.array([0, 1], dtype=np.uint64))


def test_failed_put_keeps_the_resident_block(ctx):
    import knoxdb_b200 as kb
    vals = np.arange(500, dtype=np.uint64)
    assert ctx.block_put(902, 1, 1, kb.UINT64, ko.store("bitpack", ko.U64, vals)) == 500
    before = ctx.store_stats()
    with pytest.raises(kb.KnoxError):
        ctx.block_put(902, 1, 1, kb.UINT64, np.frombuffer(b"\x04\x00", dtype=np.uint8))   # truncated bit-packed header
    assert ctx.store_stats()["blocks"] == before["blocks"]
    prog = kb.Program(ctx, [kb.Leaf(1, kb.UINT64, kb.LT, 100)])
    r = ctx.scan(prog, [(902, 1)], nrows=[500])
    assert int(r["counts"][0]) == 100
    prog.close()
