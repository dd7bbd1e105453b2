"""CPU oracle for the k-mer set distance hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  PARITY UNPINNED: see oracle/gkd_oracle.h.
"""
from .oracle import *  # noqa: F401,F403
