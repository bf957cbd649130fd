"""CPU oracle for the P2-ViT quantized forward.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or as the timed CPU baseline - never on the product path.
"""
