"""TEST INFRASTRUCTURE — CPU oracle for the FinRL env step path (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product package ``finrl_b200`` never does.
"""
