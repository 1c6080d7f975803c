"""CPU oracle for the conv hot path -- test infrastructure only (see oracle/w2x_oracle.c)."""
