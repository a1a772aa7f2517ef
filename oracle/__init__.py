"""CPU ORACLE package -- test infrastructure, NOT the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this.  See ``mp2_oracle.h`` for provenance and
pinning status ("parity unpinned" at the commitment boundary; permutations and
constants pinned by published known-answer vectors).
"""
from .oracle import *  # noqa: F401,F403
