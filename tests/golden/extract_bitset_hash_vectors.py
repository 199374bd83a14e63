#!/usr/bin/env python3
"""Transcribe numeric golden vectors for bitset popcount / index extraction and the
XXH3 integer hashes from the reference's own tests (numbers only, no code).

Sources (container only, never read at test time):
  /root/reference/internal/bitset/tests/pop.go:19-47      popcount with dirty tail bits
  /root/reference/internal/bitset/tests/run.go:32-520     bitset -> ascending index list
  /root/reference/internal/hash/xxh3_test.go:14-31        XXH3-64 of u32 / u64 inputs
Output: tests/golden/bitset_vectors.json, tests/golden/xxh3_vectors.json
"""
import json, os, re, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

def strip_comments(s): return re.sub(r"//[^\n]*", "", s)

def bytemask(size): return (0xff >> (7 - ((size - 1) & 7))) & 0xff
def FillBitset(_nil, size, val):
    n = (size + 7) // 8
    buf = [val] * n
    if n: buf[-1] &= bytemask(size)
    return buf
def fillIndex(start, length): return list(range(start, start + length))
def Repeat(b, n): return list(b) * n
def byte(x): return [x]

def go_to_py(e):
    e = re.sub(r"\[\]\w+\{", "[", e).replace("}", "]")
    e = e.replace("bytes.Repeat", "Repeat").replace("nil", "None")
    while "append(" in e:
        # append(x, y...) or append(x, byte(v))
        e2 = re.sub(r"append\(((?:[^()]|\([^()]*\))*?),\s*((?:[^()]|\([^()]*\))*?)\.\.\.\)", r"(\1 + \2)", e)
        if e2 == e:
            e2 = re.sub(r"append\(((?:[^()]|\([^()]*\))*?),\s*(byte\([^()]*\))\)", r"(\1 + \2)", e)
        if e2 == e: raise ValueError(e)
        e = e2
    return e

ENV = dict(FillBitset=FillBitset, fillIndex=fillIndex, Repeat=Repeat, byte=byte)

def field(block, name):
    m = re.search(r"\b%s:\s*" % name, block)
    if not m: return None
    i = m.end(); depth = 0; j = i
    while j < len(block):
        ch = block[j]
        if ch in "{([": depth += 1
        elif ch in "})]": depth -= 1
        elif ch == "," and depth == 0: break
        j += 1
    return block[i:j]

def blocks(text, var):
    m = re.search(r"var %s = \[\]\w+\{\n(.*)\n\}\n" % var, text, re.S)
    body = m.group(1)
    out, depth, cur = [], 0, ""
    for ch in body:
        if ch == "{":
            depth += 1
            if depth == 1: cur = ""; continue
        if ch == "}":
            depth -= 1
            if depth == 0: out.append(cur); continue
        if depth >= 1: cur += ch
    return out

def main():
    pop = strip_comments(open(os.path.join(REF, "internal/bitset/tests/pop.go")).read())
    run = strip_comments(open(os.path.join(REF, "internal/bitset/tests/run.go")).read())
    pops = []
    for b in blocks(pop, "PopCases"):
        pops.append(dict(name=eval(field(b, "Name")),
                         source=bytes(eval(go_to_py(field(b, "Source")), ENV)).hex(),
                         result=bytes(eval(go_to_py(field(b, "Result")), ENV)).hex(),
                         size=int(eval(field(b, "Size"))), count=int(eval(field(b, "Count")))))
    runs = []
    for b in blocks(run, "RunTestcases"):
        buf = eval(go_to_py(field(b, "Buf")), ENV) or []
        idx = eval(go_to_py(field(b, "Idx")), ENV)
        runs.append(dict(name=eval(field(b, "Name")), buf=bytes(buf).hex(),
                         size=int(eval(field(b, "Size"))), idx=[int(x) for x in idx]))
    with open(os.path.join(HERE, "bitset_vectors.json"), "w") as f:
        json.dump({"source": "knoxdb internal/bitset/tests/{pop,run}.go (numeric vectors only)",
                   "pop": pops, "index": runs}, f, separators=(",", ":"))
    # xxh3
    t = strip_comments(open(os.path.join(REF, "internal/hash/xxh3_test.go")).read())
    inp = re.search(r"xxh3Input = \[\]\[\]byte\{\n(.*?)\n\t\}", t, re.S).group(1)
    inputs = [[int(x) for x in re.findall(r"\d+", line)] for line in inp.strip().split("\n")]
    r32 = [int(x) for x in re.findall(r"\d+", re.search(r"xxh3Uint32Result = \[\]uint64\{(.*?)\}", t, re.S).group(1))]
    r64 = [int(x) for x in re.findall(r"\d+", re.search(r"xxh3Uint64Result = \[\]uint64\{(.*?)\}", t, re.S).group(1))]
    assert len(inputs) == len(r32) == len(r64) == 8
    with open(os.path.join(HERE, "xxh3_vectors.json"), "w") as f:
        json.dump({"source": "knoxdb internal/hash/xxh3_test.go:14-31; input bytes are read little-endian as u32 (first 4) / u64 (all 8)",
                   "input_bytes": inputs, "u32": r32, "u64": r64}, f, separators=(",", ":"))
    print(f"pop={len(pops)} index={len(runs)} xxh3=8")

if __name__ == "__main__":
    main()
