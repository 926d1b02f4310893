"""CPU oracle for the hard-sphere hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (hsmc_b200) never imports this.
"""
