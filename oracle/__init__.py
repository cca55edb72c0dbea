"""CPU oracle (test infrastructure only -- see oracle/ldpc_oracle.c)."""
