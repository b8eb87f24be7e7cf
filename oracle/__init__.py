"""
ORACLE package -- test infrastructure, NOT product code.

CPU restatement of the reference's PR-FDD PCG hot path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
