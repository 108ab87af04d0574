"""Stand-in for the Python-2 ``StringIO`` module (reference: nxblink/summary.py:30).

TEST INFRASTRUCTURE ONLY -- part of the ref-compat harness (oracle/refcompat).
"""
from io import StringIO  # noqa: F401
