"""CPU oracle of the reference's SpMM path — TEST INFRASTRUCTURE ONLY (see oracle.py)."""
