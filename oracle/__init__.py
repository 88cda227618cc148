"""CPU parity oracle — test infrastructure only (see oracle/curvis_oracle.c header)."""
