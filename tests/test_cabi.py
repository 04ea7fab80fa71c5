"""The C-ABI library loads and exports every symbol include/bgs_b200.h declares (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import PRODUCT, ROOT

HEADER = os.path.join(ROOT, "include", "bgs_b200.h")
LIB = os.path.join(PRODUCT, "csrc", "libbgs_b200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgs_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.dirname(LIB)], check=True)
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_surface():
    names = _declared()
    for required in (
        "bgs_version", "bgs_last_error", "bgs_connect_rollout", "bgs_connect_step", "bgs_connect_export",
        "bgs_connect_query", "bgs_connect_rollout_host", "bgs_bounce_rollout", "bgs_bounce_moves", "bgs_bounce_step",
    ):
        assert required in names


def test_every_declared_symbol_is_exported(lib):
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/bgs_b200.h but not exported"


def test_python_binding_covers_header():
    from simulator import _native

    assert sorted(_native.EXPORTED_SYMBOLS) == _declared()


def test_header_is_plain_c():
    # the boundary is a C ABI: the header must compile as C with nothing but <stdint.h>
    out = subprocess.run(
        ["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", "-x", "c", HEADER], capture_output=True, text=True
    )
    assert out.returncode == 0, out.stderr
    assert "torch" not in open(HEADER).read().lower().replace("pytorch", "")


def test_version_and_argument_validation(lib):
    from simulator import _native as N

    L = N.lib()
    assert L.bgs_version() == 100
    assert L.bgs_connect_supported(6, 7, 4) == 1
    assert L.bgs_connect_supported(8, 9, 5) == 1 and L.bgs_connect_supported(10, 12, 6) == 1
    assert L.bgs_connect_supported(16, 7, 4) == 1 and L.bgs_connect_supported(12, 12, 4) == 1  # byte-board fallback
    assert L.bgs_connect_supported(16, 16, 4) == 0 and L.bgs_connect_supported(4, 33, 4) == 0  # > 255 cells, > 32 columns
    assert L.bgs_connect_packed_words(6, 7) == 2 and L.bgs_connect_packed_words(10, 12) == 4
    assert L.bgs_connect_packed_words(12, 12) == 18  # byte boards: the grid itself, padded to 8 bytes
    assert L.bgs_bounce_supported(9, 6, 3) == 1 and L.bgs_bounce_supported(9, 9, 3) == 1  # 81 cells: 128-bit words
    assert L.bgs_bounce_supported(12, 11, 3) == 0 and L.bgs_bounce_supported(5, 17, 3) == 0  # > 128 cells, > 16 columns
    assert L.bgs_bounce_supported(9, 6, 16) == 0
    # unsupported configurations are reported through the status code + bgs_last_error, never a crash
    rc = L.bgs_connect_rollout(20, 20, 4, 10, 0, 0, None, None, None, None, None, None)  # 400 cells
    assert rc == -2 and "unsupported" in N.last_error()


def test_no_silent_cpu_fallback_without_a_gpu():
    import torch

    from simulator import _native as N

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = N.lib()
    rc = L.bgs_connect_rollout(6, 7, 4, 10, 0, 0, None, None, None, None, None, None)
    assert rc == -4 and "no CPU fallback" in N.last_error()
    with pytest.raises(RuntimeError):
        N.require_cuda()
