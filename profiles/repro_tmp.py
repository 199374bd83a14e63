import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "profiles")); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import knoxdb_b200 as kb
import oracle as ko
from sweep_configs import bitpack_block
rng = np.random.default_rng(1)
M4 = 1 << 22
which, npk, rows = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = kb.Context(0)
if which == "dict":
    uniq = np.unique(rng.integers(0, 2**40, 40000, dtype=np.uint64))[:32768]
    d = uniq[rng.integers(0, uniq.size, rows)]
    blk = np.frombuffer(ko.store("dict", ko.U64, d), dtype=np.uint8)
    leaf = kb.Leaf(1, kb.UINT64, kb.EQ, int(d[777]))
else:
    blk = bitpack_block(rng, rows, 15)
    leaf = kb.Leaf(1, kb.UINT64, kb.EQ, 1000 + 20000)
pin = ctx.host_array(blk.size); pin[:] = blk
for p in range(npk):
    assert ctx.block_put(p, 1, 1, kb.UINT64, pin) == rows
prog = kb.Program(ctx, [leaf])
try:
    r = ctx.scan(prog, ctx.pack_refs([(p, 1) for p in range(npk)]), nrows=[rows] * npk)
    print(which, npk, rows, os.environ.get("KX_SCAN_GEOMETRY"), "ok", r["counts"][:3])
except Exception as e:
    print(which, npk, rows, os.environ.get("KX_SCAN_GEOMETRY"), "FAIL", str(e)[:80])
