"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the reference's hot path (the parity oracle).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package `esr_b200` never imports this package.
"""
