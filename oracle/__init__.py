"""CPU oracle for the RLZ parse -- TEST INFRASTRUCTURE ONLY (see drlz_oracle.c)."""
