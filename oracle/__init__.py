"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Marin IBDWT engine.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (prmers_b200) never does.
"""
