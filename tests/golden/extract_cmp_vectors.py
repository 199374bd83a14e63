#!/usr/bin/env python3
"""Transcribe the NUMERIC golden vectors of the reference's compare-kernel tests
into a JSON fixture (numbers only, no code).

Source (read-only, only available in the build container, never at test time):
  /root/reference/internal/cmp/tests/{uint8..int64,float32,float64}.go
The Go files hold, per type, input slices (`*_s0..s4`), match operands
(`*_mat_*`), expected LSB-first bitset bytes (`*_res_*`) and case tables built
with `mk<T>(name, src, match, match2, result, length)` which tile / truncate the
source + result to `length` and mask the tail bits
(reference: internal/cmp/tests/uint64.go:118-168).

Output: tests/golden/cmp_vectors.json
  {type: {op: [ {name, n, src:[bit patterns as ints], a, b, bits: hex, count} ]}}
Float values are stored as IEEE bit patterns (uint) so NaN/Inf survive JSON.

Usage: python tests/golden/extract_cmp_vectors.py [/root/reference]
"""
import json, math, re, struct, sys, os

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "internal/cmp/tests")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cmp_vectors.json")

TYPES = {  # file stem -> (go type, bits, kind)
    "uint8": ("uint8", 8, "u"), "uint16": ("uint16", 16, "u"), "uint32": ("uint32", 32, "u"),
    "uint64": ("uint64", 64, "u"), "int8": ("int8", 8, "i"), "int16": ("int16", 16, "i"),
    "int32": ("int32", 32, "i"), "int64": ("int64", 64, "i"),
    "float32": ("float32", 32, "f"), "float64": ("float64", 64, "f"),
}
OPS = {"Equal": "eq", "NotEqual": "ne", "Less": "lt", "LessEqual": "le",
       "Greater": "gt", "GreaterEqual": "ge", "Between": "bw"}

class Math:
    MaxUint8, MaxUint16, MaxUint32, MaxUint64 = 2**8 - 1, 2**16 - 1, 2**32 - 1, 2**64 - 1
    MaxInt8, MaxInt16, MaxInt32, MaxInt64 = 2**7 - 1, 2**15 - 1, 2**31 - 1, 2**63 - 1
    MinInt8, MinInt16, MinInt32, MinInt64 = -2**7, -2**15, -2**31, -2**63
    MaxFloat32 = struct.unpack("<f", struct.pack("<I", 0x7f7fffff))[0]
    SmallestNonzeroFloat32 = struct.unpack("<f", struct.pack("<I", 1))[0]
    MaxFloat64 = sys.float_info.max
    SmallestNonzeroFloat64 = 5e-324
    @staticmethod
    def Inf(s): return math.inf if s >= 0 else -math.inf
    @staticmethod
    def NaN(): return math.nan

def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)

def go_to_py(expr):
    e = expr.strip()
    e = re.sub(r"\[\]\w+\{", "[", e)          # []T{ ... }  -> [ ... ]
    e = e.replace("}", "]")
    e = re.sub(r"make\(\[\s*\w+,\s*0\)", "[]", e)
    e = re.sub(r"make\(\[\]\w+,\s*0\)", "[]", e)
    e = re.sub(r"\bnil\b", "[]", e)
    e = re.sub(r"\bfloat32\(", "(", e).replace("float64(", "(")
    # append(a, b...) -> (a + b)
    while "append(" in e:
        e = re.sub(r"append\(([^()]*?),\s*([^()]*?)\.\.\.\)", r"(\1 + \2)", e)
    return e

def split_top(s):
    """split on commas at nesting depth 0"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{": depth += 1
        if ch in ")]}": depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip(): out.append(cur)
    return out

def parse_file(stem):
    gotype, nbits, kind = TYPES[stem]
    text = strip_comments(open(os.path.join(SRC, stem + ".go")).read())
    ns = {"math": Math}
    # ---- var ( ... ) block with slices / scalars
    m = re.search(r"\nvar \(\n(.*?)\n\)\n", text, re.S)
    body = m.group(1)
    # statements: name [type] = value   (value may span lines until braces balance)
    pos = 0
    for mm in re.finditer(r"^\s*(\w+)(?:\s+\w+)?\s*=\s*", body, re.M):
        if mm.start() < pos: continue
        start = mm.end()
        depth, i = 0, start
        while i < len(body):
            ch = body[i]
            if ch in "{(": depth += 1
            elif ch in "})": depth -= 1
            elif ch == "\n" and depth == 0: break
            i += 1
        pos = i
        ns[mm.group(1)] = eval(go_to_py(body[start:i]), {"math": Math}, ns)

    def f32r(x):  # round python float to float32 like Go's typed constant conversion
        return struct.unpack("<f", struct.pack("<f", x))[0]
    def enc(v):
        if kind == "f":
            if nbits == 64: return struct.unpack("<Q", struct.pack("<d", float(v)))[0]
            return struct.unpack("<I", struct.pack("<f", float(v)))[0]
        return int(v)

    def mk(name, src, a, b, res, length):
        src, res = list(src), list(res)
        assert len(src) % 8 == 0 and len(res) == (len(src) + 7) // 8, (stem, name)
        while length > len(src): src = src + src
        src = src[:length]
        l = (length + 7) // 8
        while l > len(res): res = res + res
        res = res[:l]
        if length % 8: res[-1] &= 0xff >> (8 - length % 8)
        return dict(name=name, n=length, src=[enc(v) for v in src], a=enc(a), b=enc(b),
                    bits=bytes(res).hex(), count=sum(bin(x).count("1") for x in res))

    out = {}
    for mm in re.finditer(r"\nvar (\w+?)(Equal|NotEqual|LessEqual|Less|GreaterEqual|Greater|Between)Cases = \[\]MatchTest\[\w+\]\{\n(.*?)\n\}\n", text, re.S):
        op = OPS[mm.group(2)]
        cases = []
        for line in mm.group(3).split("\n"):
            line = line.strip().rstrip(",")
            if not line: continue
            if line.startswith("{"):
                parts = split_top(line[1:-1])
                name = eval(parts[0]); src = eval(go_to_py(parts[1]), {"math": Math}, ns)
                a = eval(go_to_py(parts[2]), {"math": Math}, ns); b = eval(go_to_py(parts[3]), {"math": Math}, ns)
                res = eval(go_to_py(parts[4]), {"math": Math}, ns)
                cases.append(dict(name=name, n=len(src), src=[enc(v) for v in src], a=enc(a), b=enc(b),
                                  bits=bytes(res).hex(), count=int(eval(parts[5]))))
            else:
                inner = line[line.index("(") + 1: line.rindex(")")]
                parts = split_top(inner)
                args = [eval(go_to_py(p), {"math": Math}, ns) for p in parts]
                cases.append(mk(*args))
        out[op] = cases
    assert set(out) == set(OPS.values()), (stem, sorted(out))
    return out

def main():
    allv = {stem: parse_file(stem) for stem in TYPES}
    ncases = sum(len(c) for t in allv.values() for c in t.values())
    with open(OUT, "w") as f:
        json.dump({"source": "blockwatch-cc/knoxdb internal/cmp/tests/*.go (numeric vectors only)",
                   "types": allv}, f, separators=(",", ":"))
    print(f"wrote {OUT}: {len(allv)} types, {ncases} cases, {os.path.getsize(OUT)} bytes")

if __name__ == "__main__":
    main()
