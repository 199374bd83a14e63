"""Build recipe for libknoxgpu.so (sm_100a only, in-tree so it travels with gpurun snapshots).

Every source is compiled to its own object (in parallel; unchanged sources are not recompiled) and the objects are
linked into the shared library.  `ptxas -v` of every CUDA source is parsed per file: registers, stack frame and spill
bytes of each kernel land in build_info.json, and the build FAILS when a single-leaf scan kernel spills at all or the
warp-autonomous scan kernel spills more than a few registers (a spill in these persistent kernels slows every path of
them)."""
import concurrent.futures
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libknoxgpu.so")
SOURCES = ["kx_scan.cu", "kx_warp.cu", "kx_bucket.cu", "kx_string.cu", "kx_stats.cu", "kx_comm.cu", "kx_api.cu", "kx_host.cpp"]
HEADERS = ["kx_types.h", "kx_kernels.h", "kx_host.h", "kx_xxh3.h", "kx_decode.cuh", "kx_leaf.cuh", "kx_comm.h", os.path.join("..", "..", "include", "knoxgpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-Wextra,-Wno-unused-parameter,-ffp-contract=off"]
INFO = os.path.join(HERE, "build_info.json")
GENERAL_SPILL_LIMIT = 128   # bytes of spill stores the general kernel may carry (cold state around the leaf calls)


def _newest_header():
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def resource_usage(ptxas_log):
    """per kernel: registers, stack frame and spill bytes from the `ptxas -v` log of ONE source file"""
    out, cur = {}, None
    for line in ptxas_log.splitlines():
        m = re.search(r"Compiling entry function '(\w+)'", line)
        if m:
            cur = m.group(1); out[cur] = {}
            continue
        m = re.search(r"Function properties for (\w+)", line)
        if m:
            cur = m.group(1) if m.group(1) in out else None   # device functions called through the ABI are not kernels
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            out[cur].update(stack=int(m.group(1)), spill_stores=int(m.group(2)), spill_loads=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            out[cur]["registers"] = int(m.group(1))
    return out


def _compile(nvcc, src, force):
    obj = os.path.join(OBJ, src + ".o")
    log = obj + ".ptxas"
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.exists(log) and os.path.getmtime(obj) > max(os.path.getmtime(path), _newest_header()):
        return obj, open(log).read(), None
    cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        return obj, "", res.stdout + res.stderr
    open(log, "w").write(res.stderr)
    return obj, res.stderr, None


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile(nvcc, s, force), SOURCES))
    usage = {}
    for (obj, log, err), src in zip(results, SOURCES):
        if err is not None:
            sys.stderr.write(err)
            raise RuntimeError("nvcc failed compiling " + src)
        if verbose:
            sys.stderr.write(log)
        usage.update(resource_usage(log))
    res = subprocess.run([nvcc, "-shared", "-cudart", "static", "-o", OUT] + [r[0] for r in results] + ["-ldl"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libknoxgpu.so")
    json.dump(usage, open(INFO, "w"), indent=1, sort_keys=True)
    bad = []
    for k, v in sorted(usage.items()):
        spill = max(v.get("spill_stores", 0), v.get("spill_loads", 0))
        if "scan_kernel" in k and "exclusive" not in k and spill:
            bad.append(k)
        if "scan_warp_kernel" in k and spill > GENERAL_SPILL_LIMIT:
            bad.append(k)
    if bad:
        os.remove(OUT)
        raise RuntimeError("register spills in " + ", ".join(bad) + " (see knoxdb_b200/build_info.json)")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
