"""CPU checkers for the B200 CDF entropy path.  TEST INFRASTRUCTURE ONLY (see oracle/trc_oracle.c)."""
