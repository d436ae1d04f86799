"""CPU oracle for the bitHTM SP+TM step -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product
(``bithtm_b200``) never does.
"""
