"""Oracle-backed stand-in for the reference's ``simulator`` package.  TEST INFRASTRUCTURE ONLY.

Put ``oracle/pyapi`` on ``PYTHONPATH`` to run the reference's own test files, unmodified and in place,
against the CPU oracle (tests/test_reference_in_place.py).  The product package lives in
``board-game-simulator-python_b200/simulator`` and never imports this one.
"""
