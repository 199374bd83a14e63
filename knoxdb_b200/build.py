"""Build recipe for libknoxgpu.so (sm_100a only, in-tree so it travels with gpurun snapshots)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libknoxgpu.so")
SOURCES = ["kx_scan.cu", "kx_general.cu", "kx_bucket.cu", "kx_string.cu", "kx_stats.cu", "kx_comm.cu", "kx_api.cu", "kx_host.cpp"]
HEADERS = ["kx_types.h", "kx_kernels.h", "kx_host.h", "kx_xxh3.h", "kx_decode.cuh", "kx_leaf.cuh", "kx_comm.h", os.path.join("..", "..", "include", "knoxgpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-Wextra,-Wno-unused-parameter,-ffp-contract=off", "-shared", "-cudart", "static"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


INFO = os.path.join(HERE, "build_info.json")


def resource_usage(ptxas_log):
    """per kernel: registers, stack frame and spill bytes from `ptxas -v` (a spill in the persistent scan kernel costs
    every code path of it dearly: the build fails loudly instead of shipping one)"""
    import re
    out, cur = {}, None
    for line in ptxas_log.splitlines():
        m = re.search(r"Compiling entry function '(\w+)'", line)
        if m:
            cur = m.group(1); out[cur] = {}
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


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    import json
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-t", "0", "-Xptxas", "-v"] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libknoxgpu.so")
    if verbose:
        sys.stderr.write(res.stderr)
    usage = resource_usage(res.stderr)
    json.dump(usage, open(INFO, "w"), indent=1, sort_keys=True)
    spilled = sorted(k for k, v in usage.items() if ("scan_kernel" in k or "scan_general_kernel" in k) and (v.get("spill_stores") or v.get("spill_loads")))
    if spilled:
        os.remove(OUT)
        raise RuntimeError("register spills in " + ", ".join(spilled) + " (see knoxdb_b200/build_info.json)")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
