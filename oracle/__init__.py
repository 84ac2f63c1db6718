"""CPU oracles -- TEST INFRASTRUCTURE ONLY (see r1_oracle.py / bins_oracle.c headers).

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs only.
"""
