"""quick_dict_check.py — a 5-second spot check of the single-leaf kernel's dictionary IN / NOT IN path (per-tile dispatch) on
blocks of 1 … 70 001 rows (tail tiles, short passes) against the oracle; run under gpurun from the repo root.  Last run of the
round: 56 cases, 0 mismatches."""
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle as ko, kxtest as kt, knoxdb_b200 as kb
ctx = kb.Context(0)
rng = np.random.default_rng(5)
bad = 0; n_cases = 0
for t in (ko.U64, ko.I64, ko.I32, ko.U16):
    for n in (1, 31, 33, 67, 1025, 9000, 70001):
        vals = kt.typed_rand(rng, t, 40)[rng.integers(0, 40, n)]
        blob = ko.store("dict", t, vals)
        if blob is None: continue
        oc = ko.Container(t, blob)
        setv = np.unique(np.concatenate([vals[: min(3, n)], kt.typed_rand(rng, t, 3)]))
        su = ko.as_u64(t, setv)
        for neg, op in ((False, ko.IN), (True, ko.NI)):
            got, cnt = ctx.container_match(t, blob, op, values=su, nrows=n)
            ok = (got == oc.match_set(su, negate=neg)).all() and cnt == int(np.unpackbits(got).sum())
            bad += (not ok); n_cases += 1
print("dict IN/NIN tail cases:", n_cases, "bad:", bad)
