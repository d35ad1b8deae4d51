"""TEST INFRASTRUCTURE (CPU oracle). Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this."""
