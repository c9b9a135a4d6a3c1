"""TEST INFRASTRUCTURE ONLY.

Everything under oracle/ exists to check the CUDA engine, never to stand in for it:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it.  The product package (qgate_b200) never imports from here.
"""
