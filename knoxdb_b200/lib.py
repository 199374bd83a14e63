"""ctypes binding of libknoxgpu.so — mirrors include/knoxgpu.h one to one."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

INT64, INT32, INT16, INT8, UINT64, UINT32, UINT16, UINT8, FLOAT64, FLOAT32 = range(1, 11)
BYTES = 12
EQ, NE, GT, GE, LT, LE, IN, NIN, RANGE = range(1, 10)
OP_AND, OP_OR = 0xFE, 0xFF
NP = {INT64: np.int64, INT32: np.int32, INT16: np.int16, INT8: np.int8, UINT64: np.uint64, UINT32: np.uint32,
      UINT16: np.uint16, UINT8: np.uint8, FLOAT64: np.float64, FLOAT32: np.float32}

ABI_SYMBOLS = [
    "kx_abi_version", "kx_device_count", "kx_ctx_create", "kx_ctx_destroy", "kx_last_error", "kx_host_alloc", "kx_host_free",
    "kx_block_put", "kx_block_drop", "kx_store_stats", "kx_prog_compile", "kx_prog_free", "kx_scan", "kx_scan_host",
    "kx_agg_combine", "kx_last_scan_stats", "kx_cmp", "kx_bitpack_cmp", "kx_bitpack_decode", "kx_container_match",
    "kx_container_decode", "kx_bitset_op", "kx_bitset_neg", "kx_bitset_popcount", "kx_bitset_indexes", "kx_prune",
    "kx_hash_value", "kx_hash_bytes",
    "kx_scan_select", "kx_gather", "kx_gather_bytes", "kx_scan_buckets",
    "kx_stats_create", "kx_stats_free", "kx_stats_put_bloom", "kx_stats_build_bloom", "kx_stats_get_bloom", "kx_prune_stats",
    "kx_scan_ex", "kx_debug_check_guards", "kx_last_query_stats", "kx_comm_unique_id", "kx_comm_init", "kx_comm_info", "kx_scan_sharded", "kx_comm_allgather",
]
ABI_VERSION = 2
COMM_ID_BYTES = 128
GUARD = 32          # bytes of 0xFA behind every output buffer the binding allocates (internal/cmp/tests/gen.go:13-44)


class KnoxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libknoxgpu error {code}: {msg}")
        self.code = code


class _Leaf(C.Structure):
    _fields_ = [("field", C.c_uint16), ("block_type", C.c_uint8), ("mode", C.c_uint8), ("nset", C.c_uint32),
                ("a", C.c_uint64), ("b", C.c_uint64), ("set", C.POINTER(C.c_uint64))]


class _PackRef(C.Structure):
    _fields_ = [("pack", C.c_uint32), ("version", C.c_uint32)]


class _AggReq(C.Structure):
    _fields_ = [("field", C.c_uint16), ("block_type", C.c_uint8), ("reserved", C.c_uint8)]


class ScanArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("packs", C.POINTER(_PackRef)), ("npacks", C.c_int32), ("naggs", C.c_int32),
                ("row_masks", C.POINTER(C.c_void_p)), ("bitsets", C.c_void_p), ("bitset_off", C.c_void_p), ("counts", C.c_void_p),
                ("sel", C.c_void_p), ("sel_cap", C.c_size_t), ("sel_off", C.c_void_p), ("aggs", C.POINTER(_AggReq)), ("agg_out", C.c_void_p),
                ("total_count", C.POINTER(C.c_int64))]


SCAN_SHARDED = 1


class QueryStats(C.Structure):
    _fields_ = [("rows_scanned", C.c_uint64), ("packs_scanned", C.c_uint64), ("rows_matched", C.c_uint64), ("scan_time_ns", C.c_uint64),
                ("total_time_ns", C.c_uint64), ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32)]


def guarded(nbytes_or_count, dtype=np.uint8):
    """zeroed output array followed by GUARD bytes of 0xFA (the reference poisons the slack of its test outputs the same way);
    check_guard() proves that a kernel / copy did not write past the end"""
    item = np.dtype(dtype).itemsize
    raw = np.zeros(int(nbytes_or_count) * item + GUARD, dtype=np.uint8)
    raw[int(nbytes_or_count) * item:] = 0xFA
    return raw[:int(nbytes_or_count) * item].view(dtype), raw


def check_guard(raw):
    assert (raw[-GUARD:] == 0xFA).all(), "libknoxgpu wrote past the end of an output buffer"


class AggOut(C.Structure):
    _fields_ = [("count", C.c_int64), ("sum_bits", C.c_uint64), ("sum_err", C.c_double), ("min_bits", C.c_uint64),
                ("max_bits", C.c_uint64), ("valid", C.c_int32), ("reserved", C.c_int32)]

    def value(self, which, block_type):
        bits = {"sum": self.sum_bits, "min": self.min_bits, "max": self.max_bits}[which]
        if block_type == FLOAT64:
            return float(np.uint64(bits).view(np.float64))
        if block_type <= INT8:
            return int(np.uint64(bits).view(np.int64))
        return int(bits)


def library_path():
    return os.path.join(HERE, "libknoxgpu.so")


_lib = None


def lib():
    """Load libknoxgpu.so (building it in-tree first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if not os.path.exists(library_path()) or _build.stale():
        _build.build()
    L = C.CDLL(library_path())
    vp, u8p, u64p, sz = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_size_t
    sig = {
        "kx_abi_version": (C.c_int, []),
        "kx_device_count": (C.c_int, []),
        "kx_ctx_create": (C.c_int, [C.c_int, sz, C.POINTER(vp)]),
        "kx_ctx_destroy": (None, [vp]),
        "kx_last_error": (C.c_char_p, [vp]),
        "kx_host_alloc": (vp, [vp, sz]),
        "kx_host_free": (None, [vp, vp]),
        "kx_block_put": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint16, C.c_uint8, vp, sz, C.POINTER(C.c_uint32)]),
        "kx_block_drop": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint16]),
        "kx_store_stats": (C.c_int, [vp, u64p, u64p, u64p, u64p]),
        "kx_prog_compile": (C.c_int, [vp, C.POINTER(_Leaf), C.c_int, vp, C.c_int, C.POINTER(vp)]),
        "kx_prog_free": (None, [vp]),
        "kx_scan": (C.c_int, [vp, vp, C.POINTER(_PackRef), C.c_int, vp, vp, vp, C.POINTER(_AggReq), C.c_int, C.POINTER(AggOut)]),
        "kx_scan_select": (C.c_int, [vp, vp, C.POINTER(_PackRef), C.c_int, vp, sz, vp, vp, C.POINTER(_AggReq), C.c_int, C.POINTER(AggOut)]),
        "kx_scan_buckets": (C.c_int, [vp, vp, C.POINTER(_PackRef), C.c_int, C.c_uint16, C.c_uint8, vp, C.c_int, C.POINTER(_AggReq), C.c_int, vp,
                                      C.POINTER(AggOut), vp, vp]),
        "kx_gather": (C.c_int, [vp, C.POINTER(_PackRef), C.c_int, C.c_uint16, C.c_uint8, vp, vp, vp]),
        "kx_gather_bytes": (C.c_int, [vp, C.POINTER(_PackRef), C.c_int, C.c_uint16, vp, vp, vp, vp, C.c_size_t]),
        "kx_scan_host": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, vp, C.POINTER(_AggReq), C.c_int, C.POINTER(AggOut)]),
        "kx_agg_combine": (C.c_int, [C.c_uint8, C.POINTER(AggOut), C.c_int, C.POINTER(AggOut)]),
        "kx_last_scan_stats": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
        "kx_cmp": (C.c_int64, [vp, C.c_uint8, C.c_uint8, vp, sz, C.c_uint64, C.c_uint64, vp]),
        "kx_bitpack_cmp": (C.c_int64, [vp, C.c_uint8, vp, C.c_int, C.c_uint64, C.c_uint64, sz, vp]),
        "kx_bitpack_decode": (C.c_int, [vp, C.c_uint8, vp, C.c_int, C.c_uint64, sz, vp]),
        "kx_container_match": (C.c_int64, [vp, C.c_uint8, vp, sz, C.c_uint8, C.c_uint64, C.c_uint64, vp, C.c_uint32, vp]),
        "kx_container_decode": (C.c_int, [vp, C.c_uint8, vp, sz, vp, sz]),
        "kx_bitset_op": (C.c_int, [vp, C.c_int, vp, vp, sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "kx_bitset_neg": (C.c_int, [vp, vp, sz]),
        "kx_bitset_popcount": (C.c_int64, [vp, vp, sz]),
        "kx_bitset_indexes": (C.c_int64, [vp, vp, sz, vp]),
        "kx_prune": (C.c_int64, [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
        "kx_stats_create": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, vp, C.POINTER(vp)]),
        "kx_stats_free": (None, [vp]),
        "kx_stats_put_bloom": (C.c_int, [vp, C.c_int, C.c_int, vp, sz]),
        "kx_stats_build_bloom": (C.c_int, [vp, C.c_int, C.c_int, C.c_uint8, vp, vp, sz, C.c_int, C.c_int]),
        "kx_stats_get_bloom": (C.c_int, [vp, C.c_int, C.c_int, vp, sz, C.POINTER(sz)]),
        "kx_prune_stats": (C.c_int64, [vp, vp, vp, vp, vp, vp]),
        "kx_hash_value": (C.c_uint64, [C.c_uint8, C.c_uint64]),
        "kx_hash_bytes": (C.c_uint64, [vp, sz]),
        "kx_scan_ex": (C.c_int, [vp, vp, C.POINTER(ScanArgs)]),
        "kx_debug_check_guards": (C.c_int, [vp]),
        "kx_last_query_stats": (C.c_int, [vp, C.POINTER(QueryStats)]),
        "kx_comm_unique_id": (C.c_int, [vp, sz]),
        "kx_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp, sz]),
        "kx_comm_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "kx_scan_sharded": (C.c_int, [vp, vp, C.POINTER(_PackRef), C.c_int, vp, C.POINTER(_AggReq), C.c_int, C.POINTER(AggOut), C.POINTER(C.c_int64)]),
        "kx_comm_allgather": (C.c_int, [vp, vp, vp, sz]),
    }
    assert sorted(sig) == sorted(ABI_SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p) if a.size else None
    return a


def _u64(x):
    return int(x) & (2**64 - 1)


def pattern(block_type, v):
    """64-bit operand pattern of a python / numpy scalar of the block's type."""
    if block_type == FLOAT64:
        return int(np.array([v], np.float64).view(np.uint64)[0])
    if block_type == FLOAT32:
        return int(np.array([v], np.float32).view(np.uint32)[0])
    return _u64(int(v))


class Leaf:
    """One filter leaf: (field, block type, mode, operands) — filter.Filter + Matcher."""

    def __init__(self, field, block_type, mode, a=0, b=0, values=None):
        self.field, self.block_type, self.mode = field, block_type, mode
        self.a, self.b = (0, 0) if block_type == BYTES else (pattern(block_type, a), pattern(block_type, b))
        self.set = None
        self.bytes = None
        if block_type == BYTES and isinstance(a, (bytes, bytearray)):
            # row-level string predicate: operand bytes (RANGE: lower bound followed by the upper bound), lengths in a / b
            b = bytes(b) if isinstance(b, (bytes, bytearray)) else b""
            self.a, self.b = len(a), len(b) if mode == RANGE else 0
            self.bytes = np.frombuffer(bytes(a) + (b if mode == RANGE else b"") + b"\0" * 8, dtype=np.uint8).copy()
        if block_type == BYTES and values is not None and mode in (IN, NIN):
            # set of byte strings: uint32 lengths, then the bytes (include/knoxgpu.h); a = size of the buffer
            items = [bytes(v) for v in values]
            buf = np.asarray([len(v) for v in items], dtype="<u4").tobytes() + b"".join(items)
            self.a, self.b = len(buf), 0
            self.bytes = np.frombuffer(buf + b"\0" * 8, dtype=np.uint8).copy()
            self.nbytes_set = len(items)
            values = None
        if values is not None:
            arr = np.asarray(values)
            if arr.dtype.kind == "i":
                arr = arr.astype(np.int64).view(np.uint64)
            self.set = np.ascontiguousarray(arr, dtype=np.uint64)


class Program:
    def __init__(self, ctx, leaves, postfix=None):
        self.ctx = ctx
        self.leaves = list(leaves)
        if postfix is None:  # AND of all leaves
            postfix = [0] + [x for i in range(1, len(self.leaves)) for x in (i, OP_AND)]
        self.postfix = np.asarray(postfix, dtype=np.uint8)
        arr = (_Leaf * len(self.leaves))()
        for i, lf in enumerate(self.leaves):
            arr[i].field, arr[i].block_type, arr[i].mode = lf.field, lf.block_type, lf.mode
            arr[i].a, arr[i].b = lf.a, lf.b
            if lf.set is not None and lf.set.size:
                arr[i].nset = lf.set.size
                arr[i].set = lf.set.ctypes.data_as(C.POINTER(C.c_uint64))
            if getattr(lf, "bytes", None) is not None:
                arr[i].nset = getattr(lf, "nbytes_set", 1)
                arr[i].set = C.cast(lf.bytes.ctypes.data, C.POINTER(C.c_uint64))
        h = C.c_void_p()
        ctx._check(lib().kx_prog_compile(ctx.h, arr, len(self.leaves), _ptr(self.postfix), self.postfix.size, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().kx_prog_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """kx_ctx wrapper.  Raises KnoxError(KX_ENODEV) when there is no CUDA device."""

    def __init__(self, device=0, hbm_budget=0):
        h = C.c_void_p()
        rc = lib().kx_ctx_create(device, hbm_budget, C.byref(h))
        if rc != 0:
            raise KnoxError(rc, (lib().kx_last_error(None) or b"").decode())
        self.h = h

    def _check(self, rc):
        if rc < 0:
            raise KnoxError(rc, (lib().kx_last_error(self.h) or b"").decode())
        return rc

    def close(self):
        if self.h:
            self.free_host_arrays()
            lib().kx_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pinned host memory
    def host_array(self, nbytes):
        """uint8 numpy array over pinned host memory (freed with the context or free_host_arrays)."""
        p = lib().kx_host_alloc(self.h, max(int(nbytes), 1))
        if not p:
            raise KnoxError(-3, "kx_host_alloc failed")
        buf = (C.c_uint8 * max(int(nbytes), 1)).from_address(p)
        arr = np.frombuffer(buf, dtype=np.uint8, count=int(nbytes))
        self.__dict__.setdefault("_pinned", []).append(p)
        return arr

    def free_host_arrays(self):
        """release every pinned array handed out so far (the numpy views must not be used afterwards)"""
        for p in self.__dict__.get("_pinned", []):
            lib().kx_host_free(self.h, p)
        self.__dict__["_pinned"] = []

    # ---- device pack store
    def block_put(self, pack, version, field, block_type, enc):
        enc = np.frombuffer(enc, dtype=np.uint8) if not isinstance(enc, np.ndarray) else enc
        n = C.c_uint32()
        self._check(lib().kx_block_put(self.h, pack, version, field, block_type, _ptr(enc), enc.size, C.byref(n)))
        return n.value

    def block_drop(self, pack, version, field):
        self._check(lib().kx_block_drop(self.h, pack, version, field))

    def store_stats(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(lib().kx_store_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"blocks": a.value, "encoded_bytes": b.value, "device_bytes": c.value, "slab_bytes": d.value}

    # ---- multi-GPU: pack-sharded scans (one context per GPU / rank)
    @staticmethod
    def comm_unique_id():
        """rank 0: the communicator id (ncclGetUniqueId) to hand to the other ranks"""
        buf = np.zeros(COMM_ID_BYTES, dtype=np.uint8)
        rc = lib().kx_comm_unique_id(_ptr(buf), buf.size)
        if rc != 0:
            raise KnoxError(rc, (lib().kx_last_error(None) or b"").decode())
        return buf

    def comm_init(self, nranks, rank, comm_id=None):
        cid = None if comm_id is None else np.ascontiguousarray(comm_id, dtype=np.uint8)
        self._check(lib().kx_comm_init(self.h, nranks, rank, _ptr(cid), 0 if cid is None else cid.size))

    def comm_info(self):
        n, r, v = C.c_int(), C.c_int(), C.c_int()
        self._check(lib().kx_comm_info(self.h, C.byref(n), C.byref(r), C.byref(v)))
        return {"nranks": n.value, "rank": r.value, "nccl_version": v.value}

    def scan_sharded(self, prog, packs, aggs=(), want_counts=True):
        """kx_scan_sharded: this rank's packs in, totals combined over all ranks out (collective).
        Returns dict(counts (local, per pack), total_count (global), aggs (global))."""
        refs = packs if isinstance(packs, C.Array) else self.pack_refs(packs)
        counts = np.zeros(max(len(refs), 1), dtype=np.int64) if want_counts else None
        areq = (_AggReq * max(len(aggs), 1))(*[_AggReq(f, t, 0) for f, t in aggs])
        aout = (AggOut * max(len(aggs), 1))()
        total = C.c_int64()
        self._check(lib().kx_scan_sharded(self.h, prog.h, refs, len(refs), _ptr(counts), areq, len(aggs), aout, C.byref(total)))
        return {"counts": None if counts is None else counts[:len(refs)], "total_count": total.value, "aggs": list(aout)[:len(aggs)]}

    def comm_allgather(self, send):
        """every rank contributes the same number of bytes; returns [nranks, nbytes] uint8 in rank order"""
        send = np.ascontiguousarray(send).view(np.uint8).reshape(-1)
        n = self.comm_info()["nranks"]
        recv = np.zeros((n, send.size), dtype=np.uint8)
        self._check(lib().kx_comm_allgather(self.h, _ptr(send), _ptr(recv), send.size))
        return recv

    def check_guards(self):
        """guard mode (KX_GUARD=1): number of device buffers whose trailing guard zone was overwritten"""
        return self._check(lib().kx_debug_check_guards(self.h))

    def last_query_stats(self):
        q = QueryStats()
        self._check(lib().kx_last_query_stats(self.h, C.byref(q)))
        return {k: getattr(q, k) for k, _ in QueryStats._fields_ if k != "reserved"}

    # ---- scans
    @staticmethod
    def bitset_layout(nrows):
        """8-byte aligned concatenation of per-pack bitsets → (offsets, total bytes)."""
        offs, total = [], 0
        for n in nrows:
            offs.append(total)
            total += ((int(n) + 7) // 8 + 7) // 8 * 8
        return np.asarray(offs, dtype=np.uint64), total

    @staticmethod
    def pack_refs(packs):
        """kx_packref[] of a list of (pack, version) — build once, reuse across scans"""
        return (_PackRef * len(packs))(*[_PackRef(p, v) for p, v in packs])

    def scan(self, prog, packs, nrows=None, want_bitsets=False, want_counts=True, aggs=(), bitset_buf=None):
        """packs: list of (pack, version) or a prebuilt pack_refs() array. Returns dict(counts, bitsets(list of arrays), aggs)."""
        refs = packs if isinstance(packs, C.Array) else self.pack_refs(packs)
        counts = np.zeros(len(packs), dtype=np.int64) if want_counts else None
        offs = bits = None
        if want_bitsets:
            offs, total = self.bitset_layout(nrows)
            bits, raw = (bitset_buf, None) if bitset_buf is not None else guarded(total)
        areq = (_AggReq * max(len(aggs), 1))(*[_AggReq(f, t, 0) for f, t in aggs])
        aout = (AggOut * max(len(aggs), 1))()
        self._check(lib().kx_scan(self.h, prog.h, refs, len(packs), _ptr(bits), _ptr(offs), _ptr(counts), areq, len(aggs), aout))
        if want_bitsets and raw is not None:
            check_guard(raw)
        out = {"counts": counts, "aggs": list(aout)[:len(aggs)]}
        if want_bitsets:
            out["bitsets"] = [bits[int(o):int(o) + (int(n) + 7) // 8] for o, n in zip(offs, nrows)]
        return out

    def scan_select(self, prog, packs, cap=None, aggs=()):
        """kx_scan_select → dict(sel (uint32 ids), sel_off (npacks + 1), counts, aggs).  cap: capacity in ids
        (default: grows to the required size when the first call reports an overflow)."""
        refs = packs if isinstance(packs, C.Array) else self.pack_refs(packs)
        n = len(refs)
        sel_off = np.zeros(n + 1, dtype=np.uint64)
        counts = np.zeros(n, dtype=np.int64)
        areq = (_AggReq * max(len(aggs), 1))(*[_AggReq(f, t, 0) for f, t in aggs])
        aout = (AggOut * max(len(aggs), 1))()
        cap = 1 << 16 if cap is None else cap
        while True:
            sel = np.zeros(max(cap, 1), dtype=np.uint32)
            rc = lib().kx_scan_select(self.h, prog.h, refs, n, _ptr(sel), cap, _ptr(sel_off), _ptr(counts), areq, len(aggs), aout)
            if rc == -3 and int(sel_off[n]) > cap:
                cap = int(sel_off[n])
                continue
            self._check(rc)
            break
        return {"sel": sel[:int(sel_off[n])], "sel_off": sel_off, "counts": counts, "aggs": list(aout)[:len(aggs)]}

    def scan_ex(self, prog, packs, nrows=None, masks=None, want_bitsets=False, want_sel=False, sel_cap=None, aggs=(), sharded=False):
        """kx_scan_ex: masks = per pack None or a bitset (np.uint8, bit set = row stays eligible).
        Returns dict(counts, aggs[, bitsets][, sel, sel_off][, total_count])."""
        refs = packs if isinstance(packs, C.Array) else self.pack_refs(packs)
        n = len(refs)
        a = ScanArgs()
        a.struct_size = C.sizeof(ScanArgs); a.flags = SCAN_SHARDED if sharded else 0
        a.packs = refs; a.npacks = n; a.naggs = len(aggs)
        keep = []
        if masks is not None:
            keep = [None if m is None else np.ascontiguousarray(m, dtype=np.uint8) for m in masks]
            mp = (C.c_void_p * max(n, 1))(*[None if m is None else m.ctypes.data for m in keep])
            a.row_masks = mp
        counts = np.zeros(max(n, 1), dtype=np.int64)
        a.counts = counts.ctypes.data
        offs = bits = raw = None
        if want_bitsets:
            offs, total = self.bitset_layout(nrows)
            bits, raw = guarded(total)
            a.bitsets = raw.ctypes.data; a.bitset_off = offs.ctypes.data
        areq = (_AggReq * max(len(aggs), 1))(*[_AggReq(f, t, 0) for f, t in aggs])
        aout = (AggOut * max(len(aggs), 1))()
        a.aggs = areq; a.agg_out = C.cast(aout, C.c_void_p)
        total_count = C.c_int64()
        a.total_count = C.pointer(total_count)
        sel = sel_off = sraw = None
        if want_sel:
            sel_off = np.zeros(n + 1, dtype=np.uint64)
            cap = 1 << 16 if sel_cap is None else sel_cap
            while True:
                sel, sraw = guarded(cap, np.uint32)
                a.sel = sraw.ctypes.data; a.sel_cap = cap; a.sel_off = sel_off.ctypes.data
                rc = lib().kx_scan_ex(self.h, prog.h, C.byref(a))
                if rc == -3 and int(sel_off[n]) > cap:
                    cap = int(sel_off[n])
                    continue
                self._check(rc)
                check_guard(sraw)
                break
        else:
            self._check(lib().kx_scan_ex(self.h, prog.h, C.byref(a)))
        out = {"counts": counts[:n], "aggs": list(aout)[:len(aggs)]}
        if want_bitsets:
            check_guard(raw)
            out["bitsets"] = [bits[int(o):int(o) + (int(k) + 7) // 8] for o, k in zip(offs, nrows)]
        if want_sel:
            out["sel"], out["sel_off"] = sel[:int(sel_off[n])], sel_off
        if sharded:
            out["total_count"] = total_count.value
        return out

    def scan_buckets(self, prog, packs, ts_field, ts_type, edges, aggs=(), masks=None):
        """kx_scan_buckets → dict(bucket_counts (nbuckets), aggs: per value column a list of nbuckets AggOut, counts (npacks)).
        edges: nbuckets + 1 ascending window starts (values of the timestamp column's type)."""
        refs = packs if isinstance(packs, C.Array) else self.pack_refs(packs)
        n = len(refs)
        e = np.ascontiguousarray(np.asarray(edges, dtype=NP[ts_type]).astype(np.int64 if ts_type <= INT8 else np.uint64)).view(np.uint64)
        nb = e.size - 1
        counts = np.zeros(n, dtype=np.int64)
        bcounts = np.zeros(max(nb, 1), dtype=np.int64)
        areq = (_AggReq * max(len(aggs), 1))(*[_AggReq(f, t, 0) for f, t in aggs])
        aout = (AggOut * max(len(aggs) * max(nb, 1), 1))()
        mp = keep = None
        if masks is not None:
            keep = [None if m is None else np.ascontiguousarray(m, dtype=np.uint8) for m in masks]
            mp = (C.c_void_p * max(n, 1))(*[None if m is None else m.ctypes.data for m in keep])
        self._check(lib().kx_scan_buckets(self.h, prog.h, refs, n, ts_field, ts_type, _ptr(e), nb, areq, len(aggs), _ptr(bcounts), aout, _ptr(counts), mp))
        return {"bucket_counts": bcounts[:nb], "aggs": [[aout[j * nb + k] for k in range(nb)] for j in range(len(aggs))], "counts": counts}

    def gather(self, packs, field, block_type, sel, sel_off):
        """kx_gather: values of `field` at the selected rows, concatenated in pack order"""
        refs = packs if isinstance(packs, C.Array) else self.pack_refs(packs)
        sel = np.ascontiguousarray(sel, dtype=np.uint32)
        sel_off = np.ascontiguousarray(sel_off, dtype=np.uint64)
        out, raw = guarded(int(sel_off[-1]), NP[block_type])
        self._check(lib().kx_gather(self.h, refs, len(refs), field, block_type, _ptr(sel), _ptr(sel_off), _ptr(raw)))
        check_guard(raw)
        return out

    def gather_bytes(self, packs, field, sel, sel_off, capacity=None):
        """kx_gather_bytes: the byte strings of `field` at the selected rows, in pack order (a list of bytes objects).
        capacity=None asks the library for the size first (the KX_ENOMEM convention of kx_scan_select)."""
        refs = packs if isinstance(packs, C.Array) else self.pack_refs(packs)
        sel = np.ascontiguousarray(sel, dtype=np.uint32)
        sel_off = np.ascontiguousarray(sel_off, dtype=np.uint64)
        total = int(sel_off[-1])
        offs = np.zeros(total + 1, dtype=np.uint64)
        if capacity is None:
            rc = lib().kx_gather_bytes(self.h, refs, len(refs), field, _ptr(sel), _ptr(sel_off), _ptr(offs), None, 0)
            if rc not in (0, -3):   # KX_ENOMEM: offs[total] holds the required capacity
                self._check(rc)
            capacity = int(offs[total])
        out, raw = guarded(capacity)
        self._check(lib().kx_gather_bytes(self.h, refs, len(refs), field, _ptr(sel), _ptr(sel_off), _ptr(offs), _ptr(raw), capacity))
        check_guard(raw)
        buf = out.tobytes()
        return [buf[int(offs[i]):int(offs[i + 1])] for i in range(total)]

    def scan_host(self, prog, fields, blocks, nrows=None, want_bitsets=False, want_counts=True, aggs=(), bitset_buf=None):
        """fields: [(field id, block type)]; blocks: per pack a list of encoded blocks (np.uint8 arrays) per field."""
        npacks, nf = len(blocks), len(fields)
        keep = [b if isinstance(b, np.ndarray) else np.frombuffer(b, dtype=np.uint8) for row in blocks for b in row]
        ptrs = (C.c_void_p * max(len(keep), 1))(*[b.ctypes.data for b in keep])
        lens = (C.c_size_t * max(len(keep), 1))(*[b.size for b in keep])
        fid = np.asarray([f for f, _ in fields], dtype=np.uint16)
        fty = np.asarray([t for _, t in fields], dtype=np.uint8)
        counts = np.zeros(npacks, dtype=np.int64) if want_counts else None
        offs = bits = None
        if want_bitsets:
            offs, total = self.bitset_layout(nrows)
            bits, raw = (bitset_buf, None) if bitset_buf is not None else guarded(total)
        areq = (_AggReq * max(len(aggs), 1))(*[_AggReq(f, t, 0) for f, t in aggs])
        aout = (AggOut * max(len(aggs), 1))()
        self._check(lib().kx_scan_host(self.h, prog.h, npacks, _ptr(fid), _ptr(fty), nf, ptrs, lens, _ptr(bits), _ptr(offs),
                                       _ptr(counts), areq, len(aggs), aout))
        if want_bitsets and raw is not None:
            check_guard(raw)
        out = {"counts": counts, "aggs": list(aout)[:len(aggs)]}
        if want_bitsets:
            out["bitsets"] = [bits[int(o):int(o) + (int(n) + 7) // 8] for o, n in zip(offs, nrows)]
        return out

    def last_scan_stats(self):
        k, t, n = C.c_double(), C.c_double(), C.c_int()
        lib().kx_last_scan_stats(self.h, C.byref(k), C.byref(t), C.byref(n))
        return {"kernel_ms": k.value, "total_ms": t.value, "launches": n.value}

    # ---- narrow drop-ins
    def cmp(self, block_type, mode, src, a, b=0):
        src = np.ascontiguousarray(src, dtype=NP[block_type])
        bits, raw = guarded((src.size + 7) // 8)
        cnt = self._check(lib().kx_cmp(self.h, block_type, mode, _ptr(src), src.size, pattern(block_type, a), pattern(block_type, b), _ptr(raw)))
        check_guard(raw)
        return bits, cnt

    def bitpack_cmp(self, mode, packed, log2, a, b, n):
        packed = np.ascontiguousarray(packed)
        bits, raw = guarded((n + 7) // 8)
        cnt = self._check(lib().kx_bitpack_cmp(self.h, mode, _ptr(packed), log2, _u64(a), _u64(b), n, _ptr(raw)))
        check_guard(raw)
        return bits, cnt

    def bitpack_decode(self, block_type, packed, log2, minv, n):
        packed = np.ascontiguousarray(packed)
        out, raw = guarded(n, NP[block_type])
        self._check(lib().kx_bitpack_decode(self.h, block_type, _ptr(packed), log2, _u64(minv), n, _ptr(raw)))
        check_guard(raw)
        return out

    def container_match(self, block_type, enc, mode, a=0, b=0, values=None, nrows=None):
        enc = np.frombuffer(enc, dtype=np.uint8) if not isinstance(enc, np.ndarray) else enc
        bits, raw = guarded((nrows + 7) // 8)
        s = None
        if values is not None:
            s = np.asarray(values)
            s = s.astype(np.int64).view(np.uint64) if s.dtype.kind == "i" else s.astype(np.uint64)
            s = np.ascontiguousarray(s)
        cnt = self._check(lib().kx_container_match(self.h, block_type, _ptr(enc), enc.size, mode, pattern(block_type, a),
                                                   pattern(block_type, b), _ptr(s), 0 if s is None else s.size, _ptr(raw)))
        check_guard(raw)
        return bits, cnt

    def container_decode(self, block_type, enc, nrows):
        enc = np.frombuffer(enc, dtype=np.uint8) if not isinstance(enc, np.ndarray) else enc
        out, raw = guarded(nrows, NP[block_type])
        self._check(lib().kx_container_decode(self.h, block_type, _ptr(enc), enc.size, _ptr(raw), nrows))
        check_guard(raw)
        return out

    def bitset_op(self, op, dst, src, nbits):
        dst = np.ascontiguousarray(dst, dtype=np.uint8).copy()
        src = np.ascontiguousarray(src, dtype=np.uint8)
        any_, all_ = C.c_int(), C.c_int()
        self._check(lib().kx_bitset_op(self.h, op, _ptr(dst), _ptr(src), nbits, C.byref(any_), C.byref(all_)))
        return dst, bool(any_.value), bool(all_.value)

    def bitset_neg(self, buf, nbits):
        buf = np.ascontiguousarray(buf, dtype=np.uint8).copy()
        self._check(lib().kx_bitset_neg(self.h, _ptr(buf), nbits))
        return buf

    def bitset_popcount(self, buf, nbits):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        return self._check(lib().kx_bitset_popcount(self.h, _ptr(buf), nbits))

    def bitset_indexes(self, buf, nbits):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        cnt = int(np.unpackbits(buf[:(nbits + 7) // 8], bitorder="little")[:nbits].sum())
        out, raw = guarded(cnt, np.uint32)   # exactly as many ids as bits are set: any extra write hits the guard
        n = self._check(lib().kx_bitset_indexes(self.h, _ptr(buf), nbits, _ptr(raw)))
        check_guard(raw)
        return out[:n]

    def prune(self, prog, mins, maxs, blooms=None, hashes=None):
        """mins/maxs: [npacks, nleaves] uint64 patterns; blooms: [npacks][nleaves] of np.uint8 arrays or None;
        hashes: per leaf list of probe hashes."""
        mins = np.ascontiguousarray(mins, dtype=np.uint64)
        maxs = np.ascontiguousarray(maxs, dtype=np.uint64)
        npacks, nleaves = mins.shape
        out = np.zeros((npacks + 7) // 8 + 8, dtype=np.uint8)
        bp = bl = hs = ho = None
        keep = []
        if blooms is not None:
            flat = [b for row in blooms for b in row]
            keep = [None if b is None else np.ascontiguousarray(b, dtype=np.uint8) for b in flat]
            bp = (C.c_void_p * len(keep))(*[None if b is None else b.ctypes.data for b in keep])
            bl = (C.c_size_t * len(keep))(*[0 if b is None else b.size for b in keep])
            hs = np.asarray([h for hl in hashes for h in hl], dtype=np.uint64)
            ho = np.asarray(np.concatenate([[0], np.cumsum([len(hl) for hl in hashes])]), dtype=np.uint32)
        n = self._check(lib().kx_prune(self.h, prog.h, npacks, _ptr(mins), _ptr(maxs), bp, bl, _ptr(hs), _ptr(ho), _ptr(out)))
        return out[:(npacks + 7) // 8], n


class Stats:
    """kx_stats wrapper: device-resident statistics index (zone maps + bloom filters) of `npacks` data packs.
    fields: [(field id, block type)]; mins / maxs: [nfields][npacks] uint64 patterns."""

    def __init__(self, ctx, fields, mins, maxs):
        self.ctx = ctx
        mins = np.ascontiguousarray(mins, dtype=np.uint64)
        maxs = np.ascontiguousarray(maxs, dtype=np.uint64)
        self.nfields, self.npacks = mins.shape
        fid = np.asarray([f for f, _ in fields], dtype=np.uint16)
        fty = np.asarray([t for _, t in fields], dtype=np.uint8)
        h = C.c_void_p()
        ctx._check(lib().kx_stats_create(ctx.h, self.npacks, _ptr(fid), _ptr(fty), self.nfields, _ptr(mins), _ptr(maxs), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().kx_stats_free(self.h)
            self.h = None

    def put_bloom(self, field_index, pack_index, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self.ctx._check(lib().kx_stats_put_bloom(self.h, field_index, pack_index, _ptr(buf), buf.size))

    def build_bloom(self, field_index, pack_index, block_type, values, cardinality, factor, offsets=None):
        """values: numpy array of the block's type, or (BYTES) a uint8 array of concatenated strings + offsets[n+1]"""
        if block_type == BYTES:
            values = np.ascontiguousarray(values, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
            n = offsets.size - 1
        else:
            values = np.ascontiguousarray(values, dtype=NP[block_type])
            n = values.size
        self.ctx._check(lib().kx_stats_build_bloom(self.h, field_index, pack_index, block_type, _ptr(values), _ptr(offsets), n,
                                                   int(cardinality), int(factor)))

    def get_bloom(self, field_index, pack_index):
        n = C.c_size_t()
        self.ctx._check(lib().kx_stats_get_bloom(self.h, field_index, pack_index, None, 0, C.byref(n)))
        if n.value == 0:
            return None
        out = np.zeros(n.value, dtype=np.uint8)
        self.ctx._check(lib().kx_stats_get_bloom(self.h, field_index, pack_index, _ptr(out), out.size, C.byref(n)))
        return out

    def prune(self, prog, hashes=None):
        """hashes: per leaf a list of probe hashes (None: the library hashes numeric EQ / IN operands itself).
        Returns (bitset bytes over the packs, number of surviving packs)."""
        out = np.zeros((self.npacks + 7) // 8 + 8, dtype=np.uint8)
        hs = ho = None
        if hashes is not None:
            hs = np.asarray([h for hl in hashes for h in hl], dtype=np.uint64)
            ho = np.asarray(np.concatenate([[0], np.cumsum([len(hl) for hl in hashes])]), dtype=np.uint32)
            if hs.size == 0:
                hs = np.zeros(1, dtype=np.uint64)
        n = self.ctx._check(lib().kx_prune_stats(self.ctx.h, prog.h, self.h, _ptr(hs), _ptr(ho), _ptr(out)))
        return out[:(self.npacks + 7) // 8], n
