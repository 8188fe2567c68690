"""CPU oracle of the DSVT hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The product
(``dsvt-ai-trt_b200/``) never does.
"""
