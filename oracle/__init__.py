"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the SlicedNonbondedForce path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker.  The product (openmm-lab_b200/) never does.
"""
