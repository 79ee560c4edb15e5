"""Builds ``libb200ret.so`` (the C-ABI CUDA library) and, for tests, ``libb200ret_sim.so`` (CPU simulator of the
kernels' tile programs).  In-tree, explicit nvcc: the built ``.so`` travels to the GPU box with the repo snapshot.

    python -m image_retrieval_wavelet_b200.build [--force] [--sim] [--verbose]
"""
import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libb200ret.so")
SIM = os.path.join(HERE, "libb200ret_sim.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas,-ffp-contract=off", "-shared", "-cudart", "static",
    "--threads", "0",          # the translation units compile in parallel
]
NVCC_FLAGS.remove("--use_fast_math=false")   # IEEE division/sqrt stay on: parity with the float32 reference


def _sources(suffix):
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(suffix))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _all_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    deps.append(os.path.abspath(__file__))
    return deps


def nvcc_path():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libb200ret.so cannot be built here")
    return cand


def build_lib(force=False, verbose=False):
    srcs = [s for s in _sources(".cu")]
    if not force and not _stale(LIB, _all_deps()):
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-o", LIB] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    env.pop("CC", None)          # the image exports a gcc wrapper that breaks plain host builds
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


def build_sim(force=False, verbose=False):
    srcs = _sources(".cpp")
    if not force and not _stale(SIM, _all_deps()):
        return SIM
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-I", INCLUDE, "-I", CSRC,
           "-o", SIM] + srcs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return SIM


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--sim", action="store_true", help="also build the CPU simulator used by tests")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build_lib(a.force, a.verbose))
    if a.sim:
        print(build_sim(a.force, a.verbose))


if __name__ == "__main__":
    main()
