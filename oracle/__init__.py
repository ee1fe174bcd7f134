"""oracle/ — TEST INFRASTRUCTURE, not product.

CPU restatement of the reference SeisT hot path (models/seist.py, models/loss.py) used as the
parity checker. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import anything from here. The product package `seist_b200` never does.
"""
