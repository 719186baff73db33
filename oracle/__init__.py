"""TEST INFRASTRUCTURE ONLY: CPU oracle for the RDMNet dense-matching hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product package (rdmnet_b200) never does.
"""
