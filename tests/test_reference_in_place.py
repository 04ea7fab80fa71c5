"""T0: the reference's own test files, unmodified and in place, against the oracle stand-in.

This is what pins the oracle.  /root/reference only exists in the build
container; on the GPU box this test is skipped and tests/test_oracle_golden.py covers the same
positions from the committed fixture.
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF_TESTS = "/root/reference/tests"


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="/root/reference is not present on this machine")
@pytest.mark.parametrize("rules", [0, 1, 2, 4, 5, 6])
def test_reference_tests_pass_against_oracle(rules):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "oracle", "pyapi")
    env["BGS_ORACLE_BOUNCE_RULES"] = str(rules)
    out = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", REF_TESTS],
        env=env, cwd="/tmp", capture_output=True, text=True,
    )
    assert out.returncode == 0, out.stdout + out.stderr
    assert "9 passed" in out.stdout
