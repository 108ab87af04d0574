"""Stand-in package for the un-vendored dependency ``nxctmctree``.

TEST INFRASTRUCTURE ONLY (oracle/refcompat).  Written from the call-site
contracts in the reference; NOT upstream source.
"""
from . import trajectory, navigation  # noqa: F401
