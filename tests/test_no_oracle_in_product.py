"""The oracle is test infrastructure: nothing under the product package, the C-ABI sources or the
public header may import, link or mention it, and bench.py may only touch it in its CPU legs."""
import os
import re

from conftest import PRODUCT, ROOT


def _files(top, exts):
    for d, _, fs in os.walk(top):
        if "__pycache__" in d:
            continue
        for f in fs:
            if f.endswith(exts):
                yield os.path.join(d, f)


def test_product_never_references_the_oracle():
    offenders = []
    for path in list(_files(PRODUCT, (".py", ".cu", ".cuh", ".h", "Makefile"))) + [os.path.join(ROOT, "include", "bgs_b200.h")]:
        text = open(path).read()
        if re.search(r"oracle|bgso_|libbgs_oracle", text, flags=re.I):
            offenders.append(os.path.relpath(path, ROOT))
    assert offenders == []


def test_product_library_does_not_link_the_oracle():
    lib = os.path.join(PRODUCT, "csrc", "libbgs_b200.so")
    if os.path.exists(lib):
        assert b"bgso_" not in open(lib, "rb").read()


def test_bench_uses_the_oracle_only_in_cpu_legs():
    path = os.path.join(ROOT, "bench.py")
    if not os.path.exists(path):
        return
    text = open(path).read()
    for m in re.finditer(r"^(\s*)(from oracle|import oracle)", text, flags=re.M):
        assert len(m.group(1)) > 0, "bench.py must import the oracle lazily inside its CPU-baseline functions"
