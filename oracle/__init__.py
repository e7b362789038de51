"""CPU oracle — test infrastructure only (see oracle/qldpc_oracle.h)."""
