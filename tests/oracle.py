"""ctypes binding of the CPU oracle (oracle/libknox_oracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "libknox_oracle.so")

# element types (types.BlockType) and ops (types.FilterMode)
I64, I32, I16, I8, U64, U32, U16, U8, F64, F32 = range(1, 11)
EQ, NE, GT, GE, LT, LE, IN, NI, RG = range(1, 10)
TCONST, TDELTA, TRUNEND, TBITPACK, TDICT, TS8B, TRAW, TFLOATRAW = 1, 2, 3, 4, 5, 6, 7, 15
TFLOATALP = 13

NP = {I64: np.int64, I32: np.int32, I16: np.int16, I8: np.int8, U64: np.uint64, U32: np.uint32,
      U16: np.uint16, U8: np.uint8, F64: np.float64, F32: np.float32}
TYPE_BY_NAME = {"int64": I64, "int32": I32, "int16": I16, "int8": I8, "uint64": U64, "uint32": U32,
                "uint16": U16, "uint8": U8, "float64": F64, "float32": F32}
OP_BY_NAME = {"eq": EQ, "ne": NE, "gt": GT, "ge": GE, "lt": LT, "le": LE, "bw": RG, "in": IN, "ni": NI}


def build(force=False):
    srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["make", "-C", ODIR, "-s"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        u8p, u64p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_void_p
        sig = {
            "ko_put_uvarint": (C.c_int, [vp, C.c_uint64]),
            "ko_uvarint": (C.c_int, [vp, u64p]),
            "ko_cmp": (C.c_int64, [C.c_int, C.c_int, vp, C.c_size_t, C.c_uint64, C.c_uint64, vp]),
            "ko_bitset_and": (None, [vp, vp, C.c_size_t]),
            "ko_bitset_and_flag": (None, [vp, vp, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
            "ko_bitset_andnot": (None, [vp, vp, C.c_size_t]),
            "ko_bitset_or": (None, [vp, vp, C.c_size_t]),
            "ko_bitset_or_flag": (None, [vp, vp, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
            "ko_bitset_xor": (None, [vp, vp, C.c_size_t]),
            "ko_bitset_neg": (None, [vp, C.c_size_t]),
            "ko_bitset_one": (None, [vp, C.c_size_t]),
            "ko_bitset_set_range": (None, [vp, C.c_size_t, C.c_int64, C.c_int64]),
            "ko_bitset_popcount": (C.c_int64, [vp, C.c_size_t]),
            "ko_bitset_indexes": (C.c_size_t, [vp, C.c_size_t, vp]),
            "ko_bitpack_size": (C.c_size_t, [C.c_int, C.c_size_t]),
            "ko_log2range": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64]),
            "ko_bitpack_encode": (C.c_size_t, [vp, vp, C.c_size_t, C.c_int, C.c_uint64]),
            "ko_bitpack_decode": (None, [vp, vp, C.c_size_t, C.c_int, C.c_uint64]),
            "ko_bitpack_cmp": (None, [C.c_int, vp, C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, vp]),
            "ko_container_load": (C.c_long, [C.c_int, vp, C.c_size_t, C.POINTER(vp)]),
            "ko_container_free": (None, [vp]),
            "ko_container_get": (C.c_uint64, [vp, C.c_size_t]),
            "ko_container_decode": (None, [vp, vp]),
            "ko_container_match": (None, [vp, C.c_int, C.c_uint64, C.c_uint64, vp]),
            "ko_container_match_set": (None, [vp, C.c_int, vp, C.c_size_t, vp]),
            "ko_store_const": (C.c_size_t, [vp, C.c_uint64, C.c_size_t]),
            "ko_store_delta": (C.c_size_t, [vp, C.c_uint64, C.c_uint64, C.c_size_t]),
            "ko_store_raw": (C.c_size_t, [vp, C.c_int, vp, C.c_size_t]),
            "ko_store_bitpack": (C.c_size_t, [vp, C.c_int, vp, C.c_size_t]),
            "ko_store_best": (C.c_size_t, [vp, C.c_int, vp, C.c_size_t, C.c_int]),
            "ko_store_dict": (C.c_size_t, [vp, C.c_int, vp, C.c_size_t]),
            "ko_store_runend": (C.c_size_t, [vp, C.c_int, vp, C.c_size_t]),
            "ko_store_s8b": (C.c_size_t, [vp, C.c_int, vp, C.c_size_t]),
            "ko_store_bound": (C.c_size_t, [C.c_int, C.c_size_t]),
            "ko_store_alp": (C.c_size_t, [vp, vp, C.c_size_t, C.c_int, C.c_int]),
            "ko_alp_encode_single": (C.c_int64, [C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int)]),
            "ko_alp_decode": (C.c_double, [C.c_int64, C.c_int, C.c_int]),
            "ko_s8b_encode": (C.c_size_t, [vp, vp, C.c_size_t, C.c_uint64]),
            "ko_s8b_decode": (C.c_size_t, [vp, C.c_size_t, vp, C.c_size_t, C.c_uint64]),
            "ko_xxh3_u64": (C.c_uint64, [C.c_uint64]),
            "ko_xxh3_u32": (C.c_uint64, [C.c_uint32]),
            "ko_xxh3_u16": (C.c_uint64, [C.c_uint16]),
            "ko_xxh3_u8": (C.c_uint64, [C.c_uint8]),
            "ko_xxh3_bytes": (C.c_uint64, [vp, C.c_size_t]),
            "ko_bloom_bytes": (C.c_size_t, [C.c_size_t]),
            "ko_bloom_init": (None, [vp, C.c_size_t]),
            "ko_bloom_add": (None, [vp, C.c_size_t, C.c_uint64]),
            "ko_bloom_contains": (C.c_int, [vp, C.c_size_t, C.c_uint64]),
            "ko_bloom_build": (C.c_size_t, [vp, C.c_size_t, C.c_int, vp, vp, C.c_size_t, C.c_int, C.c_int]),
            "ko_reduce": (None, [C.c_int, vp, C.c_size_t, vp, vp]),
            "ko_bucket_reduce": (None, [C.c_int, vp, C.c_int, vp, C.c_size_t, vp, vp, C.c_int, vp]),
            "ko_window_edges": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, vp, C.c_int]),
            "ko_tree_eval": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_size_t, vp]),
            "ko_store_str": (C.c_size_t, [C.c_int, vp, vp, C.c_size_t, vp]),
            "ko_str_load": (C.c_long, [vp, C.c_size_t, C.POINTER(vp)]),
            "ko_str_free": (None, [vp]),
            "ko_str_len": (C.c_size_t, [vp]),
            "ko_str_get": (vp, [vp, C.c_size_t, C.POINTER(C.c_size_t)]),
            "ko_str_match": (None, [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp]),
            "ko_match_range": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
            "ko_baseline_bitpack_scan": (C.c_int64, [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_uint64, C.c_uint64, vp, C.c_int]),
            "ko_store_alprd": (C.c_size_t, [vp, C.c_int, vp, C.c_size_t, C.c_int]),
            "ko_simd_available": (C.c_int, []),
            "ko_bitpack_cmp_simd": (C.c_int, [C.c_int, vp, C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def as_u64(type_, values):
    """Sign-/zero-extended u64 carrier of a typed numpy array (floats: IEEE bits)."""
    a = np.ascontiguousarray(values, dtype=NP[type_])
    if type_ == F64:
        return a.view(np.uint64).copy()
    if type_ == F32:
        return a.view(np.uint32).astype(np.uint64)
    return a.astype(np.int64).view(np.uint64) if type_ <= I8 else a.astype(np.uint64)


def scalar_u64(type_, v):
    return int(as_u64(type_, np.array([v], dtype=NP[type_]))[0])


def nbytes(n):
    return (n + 7) // 8


def cmp(type_, op, src, a, b=0):
    """oracle compare kernel → (bitset bytes ndarray, count). a/b as raw u64 patterns."""
    src = np.ascontiguousarray(src, dtype=NP[type_])
    bits = np.zeros(nbytes(src.size) + 32, dtype=np.uint8)
    bits[nbytes(src.size):] = 0xFA  # poison guard like internal/cmp/tests/gen.go:13-44
    cnt = lib().ko_cmp(type_, op, _p(src), src.size, a & (2**64 - 1), b & (2**64 - 1), _p(bits))
    assert (bits[nbytes(src.size):] == 0xFA).all(), "oracle wrote past the bitset"
    return bits[:nbytes(src.size)].copy(), int(cnt)


class Container:
    """A loaded oracle container over encoded bytes (keeps the buffer alive)."""

    def __init__(self, type_, buf):
        self.type = type_
        self.buf = np.frombuffer(bytes(buf), dtype=np.uint8).copy()
        h = C.c_void_p()
        used = lib().ko_container_load(type_, _p(self.buf), self.buf.size, C.byref(h))
        if used < 0:
            raise ValueError("oracle: cannot load container")
        self.used, self.h = used, h
        self.n = int(C.cast(h, C.POINTER(_KoContainer)).contents.n)
        self.ctype = int(C.cast(h, C.POINTER(_KoContainer)).contents.ctype)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ko_container_free(self.h)
            self.h = None

    def decode(self):
        out = np.zeros(max(self.n, 1), dtype=np.uint64)
        lib().ko_container_decode(self.h, _p(out))
        return out[:self.n]

    def value_delta_sequences(self):
        """Decoded sequences of delta containers that see the ORIGINAL predicate operands
        (the container itself, or the Values child of nested run-end containers)."""
        out = []
        h = self.h
        while h:
            st = C.cast(h, C.POINTER(_KoContainer)).contents
            if st.ctype == TDELTA:
                seq = np.zeros(max(st.n, 1), dtype=np.uint64)
                lib().ko_container_decode(h, _p(seq))
                out.append(seq[:st.n].view(np.int64) if self.type <= I8 else seq[:st.n])
            h = C.c_void_p(st.child[0]) if st.ctype == TRUNEND and st.child[0] else None
        return out

    def match(self, op, a, b=0):
        bits = np.zeros(nbytes(self.n) + 8, dtype=np.uint8)
        lib().ko_container_match(self.h, op, a & (2**64 - 1), b & (2**64 - 1), _p(bits))
        return bits[:nbytes(self.n)].copy()

    def match_set(self, values_u64, negate=False):
        s = np.unique(np.asarray(values_u64, dtype=np.uint64))
        bits = np.zeros(nbytes(self.n) + 8, dtype=np.uint8)
        lib().ko_container_match_set(self.h, int(negate), _p(s), s.size, _p(bits))
        return bits[:nbytes(self.n)].copy()


class _KoContainer(C.Structure):
    _fields_ = [("ctype", C.c_int), ("type", C.c_int), ("n", C.c_size_t), ("val", C.c_uint64),
                ("delta", C.c_uint64), ("log2", C.c_int), ("payload", C.c_void_p),
                ("payload_len", C.c_size_t), ("child", C.c_void_p * 3), ("alp_e", C.c_int), ("alp_f", C.c_int),
                ("alp_flags", C.c_int)]


def store(kind, type_, values=None, **kw):
    """Encode values (typed numpy array) with the named scheme → bytes (container only)."""
    L = lib()
    if kind == "const":
        buf = np.zeros(64, dtype=np.uint8)
        n = L.ko_store_const(_p(buf), scalar_u64(type_, kw["val"]), kw["n"])
        return buf[:n].tobytes()
    if kind == "delta":
        buf = np.zeros(64, dtype=np.uint8)
        n = L.ko_store_delta(_p(buf), scalar_u64(type_, kw["base"]), scalar_u64(type_, kw["delta"]), kw["n"])
        return buf[:n].tobytes()
    v = as_u64(type_, values)
    if kind == "alprd":   # FloatAlpRdContainer[float64,uint64] / [float32,uint32]; shift: the cut (default: analysed)
        assert type_ in (F64, F32)
        buf = np.zeros(3 * L.ko_store_bound(U64, v.size) + 64, dtype=np.uint8)
        n = L.ko_store_alprd(_p(buf), type_, _p(v), v.size, kw.get("shift", -1))
        return buf[:n].tobytes()
    if kind == "alp":   # FloatAlpContainer[float64,int64]; e/f: exponents (default: chosen by sampling)
        assert type_ == F64
        buf = np.zeros(3 * L.ko_store_bound(I64, v.size) + 64, dtype=np.uint8)
        n = L.ko_store_alp(_p(buf), _p(v), v.size, kw.get("e", -1), kw.get("f", -1))
        return buf[:n].tobytes()
    buf = np.zeros(L.ko_store_bound(type_, v.size), dtype=np.uint8)
    fn = {"raw": L.ko_store_raw, "bitpack": L.ko_store_bitpack, "dict": L.ko_store_dict,
          "runend": L.ko_store_runend, "s8b": L.ko_store_s8b}.get(kind)
    if kind == "best":
        n = L.ko_store_best(_p(buf), type_, _p(v), v.size, kw.get("lvl", 3))
    else:
        n = fn(_p(buf), type_, _p(v), v.size)
    return buf[:n].tobytes()


class Agg(C.Structure):
    _fields_ = [("count", C.c_int64), ("sum_bits", C.c_uint64), ("min_bits", C.c_uint64),
                ("max_bits", C.c_uint64), ("valid", C.c_int)]


def reduce(type_, values, bits=None, state=None):
    v = as_u64(type_, values)
    st = state or Agg()
    b = np.ascontiguousarray(bits, dtype=np.uint8) if bits is not None else None
    lib().ko_reduce(type_, _p(v), v.size, _p(b) if b is not None else None, C.byref(st))
    return st


def bucket_reduce(type_, values, ts_type, ts, bits, edges, states=None):
    """ko_bucket_reduce: per-window reducer states (list of Agg) of the matching rows; edges = nbuckets + 1 window starts"""
    ts64 = as_u64(ts_type, ts)
    e = as_u64(ts_type, np.asarray(edges, dtype=NP[ts_type]))
    nb = e.size - 1
    st = states if states is not None else (Agg * nb)()
    v = as_u64(type_, values) if values is not None else None
    b = np.ascontiguousarray(bits, dtype=np.uint8) if bits is not None else None
    lib().ko_bucket_reduce(type_, _p(v) if v is not None else None, ts_type, _p(ts64), ts64.size, _p(b) if b is not None else None, _p(e), nb, st)
    return st


def window_edges(t_from, t_to, step):
    """ko_window_edges: window starts of a fixed-duration unit (From, then aligned multiples of step) up to >= To"""
    out = np.zeros(int((t_to - t_from) // step) + 4, dtype=np.int64)
    n = lib().ko_window_edges(int(t_from), int(t_to), int(step), _p(out), out.size)
    return out[:n].copy()


vp_t = C.c_void_p
STR_CONST, STR_FIXED, STR_COMPACT, STR_DICT = 16, 17, 18, 19


def store_str(kind, rows):
    """Encode a list of byte strings with the named string container (ids 16..19) → bytes, or None when the rows do
    not fit the scheme (constant: all rows equal; fixed: equal lengths)"""
    flat = np.frombuffer(b"".join(rows), dtype=np.uint8) if rows else np.zeros(0, dtype=np.uint8)
    offs = np.zeros(len(rows) + 1, dtype=np.uint32)
    offs[1:] = np.cumsum([len(r) for r in rows])
    flat = np.ascontiguousarray(np.concatenate([flat, np.zeros(8, dtype=np.uint8)]))
    buf = np.zeros(flat.size + 3 * lib().ko_store_bound(U32, len(rows) + 1) + 64, dtype=np.uint8)
    n = lib().ko_store_str(kind, _p(flat), _p(offs), len(rows), _p(buf))
    return buf[:n].tobytes() if n else None


class StrContainer:
    """ko_str_load + Get / Match<Op> of the string containers"""

    def __init__(self, blob):
        self.blob = np.frombuffer(blob, dtype=np.uint8).copy()
        h = vp_t()
        used = lib().ko_str_load(_p(self.blob), self.blob.size, C.byref(h))
        assert used == self.blob.size, (used, self.blob.size)
        self.h, self.n = h, lib().ko_str_len(h)

    def get(self, i):
        ln = C.c_size_t()
        p = lib().ko_str_get(self.h, i, C.byref(ln))
        return C.string_at(p, ln.value)

    def match(self, op, a, b=b""):
        bits = np.zeros(nbytes(self.n) + 8, dtype=np.uint8)
        aa, bb = np.frombuffer(a + b"\0", dtype=np.uint8), np.frombuffer(b + b"\0", dtype=np.uint8)
        lib().ko_str_match(self.h, op, _p(aa), len(a), _p(bb), len(b), _p(bits))
        return bits[:nbytes(self.n)].copy()

    def __del__(self):
        try:
            lib().ko_str_free(self.h)
        except Exception:
            pass


def tree_eval(postfix, leaf_bits, n):
    pf = np.asarray(postfix, dtype=np.uint8)
    arrs = [np.ascontiguousarray(b, dtype=np.uint8) for b in leaf_bits]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    out = np.zeros(nbytes(n) + 8, dtype=np.uint8)
    rc = lib().ko_tree_eval(_p(pf), pf.size, ptrs, len(arrs), n, _p(out))
    assert rc == 0
    return out[:nbytes(n)].copy()


def bloom_build(values, cardinality, factor, elem_bytes=None, offsets=None):
    """stats.BuildBloomFilter restated: values = typed numpy array, or uint8 bytes + offsets for strings"""
    if offsets is not None:
        v = np.ascontiguousarray(values, dtype=np.uint8); off = np.ascontiguousarray(offsets, dtype=np.uint32)
        n, eb = off.size - 1, 0
    else:
        v = np.ascontiguousarray(values); off = None
        n, eb = v.size, elem_bytes or v.dtype.itemsize
        v = v.view(np.uint8)
    m = 8
    while m < cardinality * factor * 8:
        m <<= 1
    buf = np.zeros(1 + m // 8, dtype=np.uint8)
    ln = lib().ko_bloom_build(_p(buf), buf.size, eb, _p(v), None if off is None else _p(off), n, cardinality, factor)
    return buf[:ln] if ln else None
