"""CPU oracle for the SiGMA SpMV + Krylov hot path (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``sigma_b200``) never does.
"""
from .oracle import *  # noqa: F401,F403
