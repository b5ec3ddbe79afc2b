"""Builds multi-purpose-mpc_b200/libmpc_b200.so in-tree with nvcc for sm_100a.

Run directly (`python multi-purpose-mpc_b200/build_ext.py`) or through __graft_entry__.build().
The three translation units get different floating-point flags on purpose:
  geometry.cu  -fmad=false  (grid cells / widths are compared bit-for-bit with numpy arithmetic)
  admm.cu      default      (FMA contraction on: the QP solution is compared within a tolerance)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmpc_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
UNITS = [("admm.cu", ["-DMPC_TUNING_VARIANTS"] if os.environ.get("MPC_TUNING_VARIANTS") else []), ("admm_pair.cu", []), ("admm_quad.cu", []), ("admm_tm.cu", []), ("geometry.cu", ["-fmad=false"]), ("engine.cu", ["-fmad=false"]),
         ("speed_profile.cu", ["-fmad=false"])]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(HERE, "..", "include", "mpc_b200.h"))
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer([s] + hdrs, o):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _newer(objs, OUT):
        tmp = OUT + ".tmp.%d" % os.getpid()   # link next to the target, then rename: the library is replaced atomically
        cmd = [nvcc] + ARCH + ["-shared", "-o", tmp] + objs + ["-ccbin", "/usr/bin/g++", "-ldl"]
        subprocess.check_call(cmd)
        os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
