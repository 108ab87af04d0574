"""Stand-in package for the un-vendored dependency ``nxmctree``.

TEST INFRASTRUCTURE ONLY (oracle/refcompat).  Written from the call-site
contracts in the reference and the textbook algorithm; NOT upstream source.
"""
from . import sampling, util  # noqa: F401
