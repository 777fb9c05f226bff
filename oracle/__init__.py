"""CPU oracle (test infrastructure). See parseq_oracle.py header."""
