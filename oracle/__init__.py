"""CPU oracle of the particle-robot update.  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the product."""
