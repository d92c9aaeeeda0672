"""ORACLE package: CPU restatement of the reference hot path (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (sqp_solver_b200) never does.
"""
