"""Builds the native libraries in-tree (nvcc for sm_100a, gcc for the host C side).

    python -m dune_sculpt_b200.build            # CUDA library + host library
    python -m dune_sculpt_b200.build --oracle   # also the CPU oracle (test infrastructure)

Outputs: dune_sculpt_b200/lib/libdune_sculpt_cuda.so, dune_sculpt_b200/lib/libdune_sculpt_host.so
(git-ignored; they travel to the GPU box with the tree).
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dune_sculpt_b200")
LIB = os.path.join(PKG, "lib")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-parity with the CPU path: no FMA contraction; IEEE div/sqrt and denormals are nvcc defaults
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_cuda(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(PKG, "csrc", "dsc_api.cu")
    deps = [src, os.path.join(PKG, "csrc", "dsc_kernels.cuh"), os.path.join(ROOT, "include", "dune_sculpt_cuda.h")]
    extra = [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc")) if f.endswith((".cu", ".cuh"))]
    out = os.path.join(LIB, "libdune_sculpt_cuda.so")
    if force or _newer(out, deps + extra):
        srcs = sorted(f for f in extra if f.endswith(".cu"))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs
        subprocess.run(cmd, check=True)
    return out


def build_host(force=False):
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(PKG, "host", "dune_pbvh.c")
    deps = [src, os.path.join(ROOT, "include", "dune_pbvh.h"), os.path.join(ROOT, "include", "dune_sculpt_cuda.h")]
    out = os.path.join(LIB, "libdune_sculpt_host.so")
    if force or _newer(out, deps):
        cmd = [GCC, "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp", "-Wall", "-Wextra", "-o", out, src,
               "-L" + LIB, "-ldune_sculpt_cuda", "-Wl,-rpath,$ORIGIN", "-lm"]
        subprocess.run(cmd, check=True)
    return out


def build_oracle(force=False):
    d = os.path.join(ROOT, "oracle")
    if force:
        subprocess.run(["make", "-C", d, "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", d], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(d, "build", "liboracle.so")


def build_oracle_ref(force=False):
    """oracle/_ref/libref.so: the reference's own hot-path functions, cut out of /root/reference at build time
    (oracle/ref_extract.py) -- test infrastructure that pins the oracle.  Where /root/reference is absent (the GPU box)
    the prebuilt library travels with the tree; returns None when there is neither."""
    d = os.path.join(ROOT, "oracle")
    so = os.path.join(d, "_ref", "libref.so")
    srcs = [os.path.join(d, f) for f in ("ref_extract.py", "ref_shim.h", "ref_api.c", "ref_glue.inc")]
    if os.path.isdir("/root/reference") and (force or _newer(so, srcs)):
        subprocess.run([sys.executable, os.path.join(d, "ref_extract.py")], check=True, stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


def build_all(force=False, oracle=False, verbose=False):
    outs = [build_cuda(force, verbose), build_host(force)]
    if oracle:
        outs.append(build_oracle(force))
        ref = build_oracle_ref(force)
        if ref:
            outs.append(ref)
    return outs


if __name__ == "__main__":
    for o in build_all(force="--force" in sys.argv, oracle="--oracle" in sys.argv, verbose="-v" in sys.argv):
        print(o)
