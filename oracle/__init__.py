"""CPU oracle for the RNA-Bloom k-mer / Bloom-filter hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  PARITY UNPINNED: see rnabloom_oracle.c header and DESIGN.md "Oracle".
"""
from .binding import Oracle, load, build  # noqa: F401
