"""Stand-in for the un-vendored ``nxrate`` (only ``nxrate.testing`` asserts are
used by the reference: examples/p53/create_mg94.py:55-57,171-177).
TEST INFRASTRUCTURE ONLY."""
from . import testing  # noqa: F401
